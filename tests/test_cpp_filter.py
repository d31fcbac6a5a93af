"""The C++ filter class (include/mci_b200_filter.hpp) through the driver tests/cpp/mci_filter_test.cpp, which is
shaped like the reference's own test driver (test/itkMorphologicalContourInterpolationTest.cxx): CPU part = interface
semantics and error behaviour without a device; GPU part = the reference's regression matrix (algorithms B / C / T,
axis, label; test/CMakeLists.txt:51-61) on the fixtures, compared with the oracle voxel for voxel."""
import os
import subprocess

import numpy as np
import pytest

import cases
from itkcontourinterpolation_b200 import build as B

TYPE = {np.dtype(np.uint8): "uchar", np.dtype(np.uint16): "ushort", np.dtype(np.int16): "short"}


@pytest.fixture(scope="module")
def driver():
    return B.build_cpp_driver()


def test_api_semantics_and_no_device_error(driver):
    r = subprocess.run([driver, "--api"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "api ok" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert "no device" in r.stdout and "no CPU fallback" in r.stdout


def test_usage_and_bad_input_fail_like_the_reference_driver(driver, tmp_path):
    assert subprocess.run([driver], capture_output=True).returncode != 0                      # usage -> EXIT_FAILURE
    r = subprocess.run([driver, str(tmp_path / "missing.raw"), str(tmp_path / "o.raw"), "4", "4", "4", "ushort"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Error:" in r.stderr                                         # exception -> EXIT_FAILURE


def run_driver(driver, tmp_path, vol, algo=None, axis=None, label=None):
    src, dst = tmp_path / "in.raw", tmp_path / "out.raw"
    vol.tofile(src)
    z, y, x = vol.shape
    cmd = [driver, str(src), str(dst), str(x), str(y), str(z), TYPE[vol.dtype]]
    for a in (algo, axis, label):
        if a is None:
            break
        cmd.append(str(a))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    slices = set(tuple(int(t) for t in line.split()) for line in r.stdout.splitlines() if line.strip())
    return np.fromfile(dst, vol.dtype).reshape(vol.shape), slices


DRIVER_ALGOS = {"B": dict(use_distance_transform=False, use_ball=True), "C": dict(use_distance_transform=False, use_ball=False),
                "T": dict(use_distance_transform=True, use_ball=True), "D": dict(use_distance_transform=True, use_ball=False)}


@pytest.mark.gpu
def test_regression_matrix_through_the_cpp_class(driver, tmp_path, seven_labels, many_to_many):
    from oracle import oracle_lib as O
    inputs = {"SevenLabels": seven_labels.astype(np.uint8), "ManyToMany": many_to_many}
    h = cases.handcrafted()
    for name in ("Empty", "NoSlices", "OneToThree", "RingAndIsland", "ThreeAxis", "DriftOddGaps", "BorderClip"):
        inputs[name] = h[name]
    inputs["TwoAxisTwoLabel_short"] = h["TwoAxisTwoLabel"].astype(np.int16)
    for name, vol in inputs.items():
        for algo, kw in DRIVER_ALGOS.items():
            out, slices = run_driver(driver, tmp_path, vol, algo)
            assert np.array_equal(out, O.interpolate(vol, **kw)), (name, algo)
            trip, _ = O.detect(vol)
            assert slices == set(map(tuple, trip.tolist())), (name, algo)
    # default algorithm of the driver is B; axis and label arguments (test/CMakeLists.txt:110-125)
    vol = h["ThreeAxis"]
    out, _ = run_driver(driver, tmp_path, vol)
    assert np.array_equal(out, O.interpolate(vol, **DRIVER_ALGOS["B"]))
    for axis in (0, 1, 2):
        out, _ = run_driver(driver, tmp_path, vol, "T", axis)
        assert np.array_equal(out, O.interpolate(vol, axis=axis, **DRIVER_ALGOS["T"])), axis
    vol = h["TwoAxisTwoLabel"]
    for label in (5, 9, 77):
        out, _ = run_driver(driver, tmp_path, vol, "C", -1, label)
        assert np.array_equal(out, O.interpolate(vol, label=label, **DRIVER_ALGOS["C"])), label


@pytest.mark.gpu
def test_example_program_interpolate_then_median_smooth(driver, tmp_path, seven_labels, many_to_many):
    # examples/mciExample.cxx: Image<unsigned char, 3>, default settings, MedianImageFilter radius 2 (default) or given
    from oracle import oracle_lib as O
    for vol, radius in ((seven_labels.astype(np.uint8), None), (many_to_many.astype(np.uint8), 1), (many_to_many, 3)):
        src, dst = tmp_path / "in.raw", tmp_path / "out.raw"
        vol.tofile(src)
        z, y, x = vol.shape
        cmd = [driver, "--example", str(src), str(dst), str(x), str(y), str(z), TYPE[vol.dtype]] + ([str(radius)] if radius else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        got = np.fromfile(dst, vol.dtype).reshape(vol.shape)
        assert np.array_equal(got, O.median_image_filter(O.interpolate(vol), radius or 2))
