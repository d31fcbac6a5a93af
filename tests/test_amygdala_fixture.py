"""The reference's third real input, 64816L_amygdala_int.nii.gz (wasm/test/data/input/, also an input of its
regression suite), through the reference's own regression matrix (test/CMakeLists.txt:143-160): algorithms B / C / T x
{all axes, one axis, one axis + one label}, pixel type short like its driver instantiates (test/...Test.cxx:110-114).
CPU: the oracle against the frozen hashes + invariants, and the NIfTI reader.  GPU: CUDA path against the oracle."""
import gzip
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import cases
from oracle import oracle_lib as O

GOLD = json.load(open(os.path.join(cases.GOLDEN, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_fixture_is_what_the_reference_ships():
    v = cases.amygdala()
    assert v.shape == (210, 400, 170) and v.dtype == np.int16
    assert sorted(np.unique(v).tolist()) == [0, 1, 2, 3, 4]
    src = "/root/reference/wasm/test/data/input/64816L_amygdala_int.nii.gz"
    if os.path.exists(src):  # only in the build container
        from itkcontourinterpolation_b200.nifti import read_nifti
        assert np.array_equal(read_nifti(src), v)


def test_nifti_reader_round_trip(tmp_path):
    from itkcontourinterpolation_b200.nifti import read_nifti
    rng = np.random.default_rng(2)
    for dt, code, end in ((np.int16, 4, "<"), (np.uint8, 2, "<"), (np.uint16, 512, ">")):
        vol = rng.integers(0, 100, (5, 6, 7)).astype(dt)
        hdr = bytearray(352)
        hdr[0:4] = struct.pack(end + "i", 348)
        hdr[40:56] = struct.pack(end + "8h", 3, 7, 6, 5, 1, 1, 1, 1)
        hdr[70:74] = struct.pack(end + "2h", code, 8 * np.dtype(dt).itemsize)
        hdr[108:112] = struct.pack(end + "f", 352.0)
        hdr[344:348] = b"n+1\0"
        body = vol.astype(np.dtype(dt).newbyteorder(end)).tobytes()
        for name, data in (("a.nii", bytes(hdr) + body), ("a.nii.gz", gzip.compress(bytes(hdr) + body))):
            p = tmp_path / name
            p.write_bytes(data)
            assert np.array_equal(read_nifti(str(p)), vol)
    (tmp_path / "bad.nii").write_bytes(b"\0" * 400)
    with pytest.raises(ValueError):
        read_nifti(str(tmp_path / "bad.nii"))


def test_oracle_regression_matrix_frozen_and_invariants():
    v = cases.amygdala()
    for algo in "BCT":
        for axis, label in cases.AMYGDALA_ARGS:
            out = O.interpolate(v, axis=axis, label=label, threads=os.cpu_count() or 1, **cases.ALGOS[algo])
            assert sha(out) == GOLD["64816L_amygdala_int_%s_%d_%d" % (algo, axis, label)]["output"], (algo, axis, label)
            assert np.array_equal(out[v != 0], v[v != 0])                      # output contains the input (X:1623-1636)
            if label:
                assert set(np.unique(out[out != v]).tolist()) <= {label}       # only the chosen label is interpolated
    # restricting the axis can only remove interpolated voxels of that algorithm's all-axes run's candidates
    allax = O.interpolate(v, **cases.ALGOS["B"])
    per_axis = [O.interpolate(v, axis=a, **cases.ALGOS["B"]) for a in (0, 1, 2)]
    combined = np.maximum.reduce(per_axis)                                     # cross-axis combine = max (X:754-763)
    assert np.array_equal(combined, allax)


@pytest.mark.gpu
def test_gpu_regression_matrix_matches_oracle():
    from itkcontourinterpolation_b200 import MorphologicalContourInterpolator
    v = cases.amygdala()
    f = MorphologicalContourInterpolator(0)
    try:
        for algo in "BCTD":
            kw = cases.ALGOS[algo]
            for axis, label in cases.AMYGDALA_ARGS:
                f.SetUseDistanceTransform(kw["use_distance_transform"])
                f.SetUseBallStructuringElement(kw["use_ball"])
                f.SetAxis(axis)
                f.SetLabel(label)
                f.SetInput(v)
                out = f.Update()
                if algo != "D":
                    assert sha(out) == GOLD["64816L_amygdala_int_%s_%d_%d" % (algo, axis, label)]["output"], (algo, axis, label)
                else:
                    assert np.array_equal(out, O.interpolate(v, axis=axis, label=label, **kw)), (algo, axis, label)
    finally:
        f.close()
