"""Label smoothing step of the reference's example (examples/mciExample.cxx:29-39: itk::MedianImageFilter, radius 2
by default) -- SURVEY 8(f) rank 3.  CPU: the oracle restatement against scipy.ndimage.median_filter (an independent
implementation of the same published definition).  GPU: the CUDA kernel against the oracle, bit-exact."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle_lib as O


def rand_labels(rng, shape, dtype, n=4):
    lo = -3 if np.dtype(dtype) == np.int16 else 0
    return rng.integers(lo, n, shape).astype(dtype)


def test_oracle_matches_scipy_median_filter():
    from scipy import ndimage
    rng = np.random.default_rng(3)
    for dt in (np.uint8, np.uint16, np.int16):
        for shape in ((7, 9, 11), (1, 5, 6), (16, 5, 16), (3, 3, 3), (2, 1, 40)):
            v = rand_labels(rng, shape, dt, 9)
            for r in (0, 1, 2, 3):
                assert np.array_equal(O.median_image_filter(v, r, threads=3),
                                      ndimage.median_filter(v, size=2 * r + 1, mode="nearest")), (dt, shape, r)


def test_oracle_on_interpolated_fixture_is_a_smoothing(seven_labels):
    out = O.interpolate(seven_labels.astype(np.uint8))
    sm = O.median_image_filter(out, 2)
    assert sm.shape == out.shape and set(np.unique(sm)) <= set(np.unique(out))
    assert np.array_equal(O.median_image_filter(np.full((5, 6, 7), 4, np.uint16), 3), np.full((5, 6, 7), 4, np.uint16))


@pytest.mark.gpu
def test_gpu_median_matches_oracle_on_small_volumes():
    from itkcontourinterpolation_b200 import median_image_filter
    rng = np.random.default_rng(8)
    shapes = ((7, 9, 11), (1, 5, 6), (16, 5, 16), (3, 3, 3), (2, 1, 40), (9, 17, 33), (8, 8, 32), (10, 24, 70), (1, 1, 1))
    for dt in (np.uint8, np.uint16, np.int16):
        for shape in shapes:
            for n in (2, 3, 4, 5, 9, 4000):   # 2..4 values per tile: bit-mask path; more: generic path; noise
                v = rand_labels(rng, shape, dt, min(n, 250) if dt == np.uint8 else n)
                for r in (0, 1, 2, 3):
                    assert np.array_equal(median_image_filter(v, r), O.median_image_filter(v, r, threads=4)), (dt, shape, n, r)
    # label-like blobs: tiles with one to four values, uniform neighbourhoods next to mixed ones, borders
    blobs = np.zeros((37, 70, 150), np.uint16)
    zz, yy, xx = np.mgrid[:37, :70, :150]
    for k, (cz, cy, cx, rr) in enumerate(((10, 20, 30, 9), (14, 26, 41, 8), (25, 50, 100, 12), (30, 60, 140, 10), (5, 5, 5, 7), (18, 35, 75, 15))):
        blobs[(zz - cz) ** 2 + (yy - cy) ** 2 + (xx - cx) ** 2 <= rr * rr] = 3 + 7 * k
    for dt in (np.uint16, np.uint8, np.int16):
        b = blobs.astype(dt)
        if dt == np.int16:
            b[b == 3] = -5
        for r in (1, 2, 3):
            assert np.array_equal(median_image_filter(b, r), O.median_image_filter(b, r, threads=4)), (dt, r)
    big = np.zeros((20, 40, 96), np.int16)
    big[5:15, 10:30, 20:70] = -7
    big[8:12, 15:25, 30:50] = 9
    for r in (1, 2, 3):
        assert np.array_equal(median_image_filter(big, r), O.median_image_filter(big, r, threads=4))


@pytest.mark.gpu
def test_gpu_example_pipeline_interpolate_then_smooth(seven_labels, many_to_many):
    # examples/mciExample.cxx: interpolate (defaults), then MedianImageFilter radius 2; SevenLabels is read as uchar
    from itkcontourinterpolation_b200 import interpolate_and_smooth
    for vol in (seven_labels.astype(np.uint8), many_to_many):
        for r in (1, 2, 3):
            got = interpolate_and_smooth(vol, smoothing_radius=r)
            assert np.array_equal(got, O.median_image_filter(O.interpolate(vol), r, threads=4)), r
    vol, cfg = cases.synthetic_small()["C3"]
    got = interpolate_and_smooth(vol, smoothing_radius=2, axis=2)
    assert np.array_equal(got, O.median_image_filter(O.interpolate(vol, axis=2), 2, threads=os.cpu_count() or 1))


@pytest.mark.gpu
def test_gpu_median_full_size_c3_volume():
    # the smoothing step on the interpolated C3 volume (512x512x256 uint16): the example's pipeline at BASELINE size
    from itkcontourinterpolation_b200 import interpolate_and_smooth
    from itkcontourinterpolation_b200.synthetic import make_config
    vol, _, cfg = make_config("C3")
    got = interpolate_and_smooth(vol, smoothing_radius=2, axis=2)
    ref = O.median_image_filter(O.interpolate(vol, axis=2, threads=os.cpu_count() or 1), 2, threads=os.cpu_count() or 1)
    assert np.array_equal(got, ref)
