"""B200-native slice-pair interpolation path of itk::MorphologicalContourInterpolator.

Importing the package does not load the CUDA library; constructing a filter does, and fails loudly
without a GPU (there is no CPU fallback).
"""
from .filter import (MorphologicalContourInterpolator, VolumeSeriesInterpolator,  # noqa: F401
                     morphological_contour_interpolator, median_image_filter, interpolate_and_smooth)
from .nrrd import read_nrrd  # noqa: F401
