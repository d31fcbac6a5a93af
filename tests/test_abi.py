"""The C-ABI shared library: builds, loads, exports every symbol include/mci_b200.h declares, and
refuses to run without a GPU (no CPU fallback).  No compute calls here.  CPU only."""
import ctypes
import os
import re

import pytest

from itkcontourinterpolation_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mci_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mci_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmci_b200.so does not export " + n
    bound = sorted(s[0] for s in L.SYMBOLS)
    assert bound == names, "ctypes binding and header disagree"


def test_struct_layouts_match_header_sizes():
    assert ctypes.sizeof(L.LabelInfo) == 32
    assert ctypes.sizeof(L.LabeledSlice) == 16
    assert ctypes.sizeof(L.SliceDesc) == 32
    assert ctypes.sizeof(L.ComponentInfo) == 40
    assert ctypes.sizeof(L.Pair) == 8
    assert ctypes.sizeof(L.Job) == 32
    assert ctypes.sizeof(L.Params) == 20
    assert ctypes.sizeof(L.FilterOptions) == 40
    assert ctypes.sizeof(L.FilterStats) == 14 * 8


def test_default_options_mirror_reference_defaults():
    # H:229-235: label 0, axis -1, heuristic on, distance transform on, cross SE, auto slices, extrapolation on
    o = L.FilterOptions()
    L.lib().mci_filter_default_options(ctypes.byref(o))
    assert (o.label, o.axis, o.heuristic_alignment, o.use_distance_transform, o.use_ball_structuring_element,
            o.use_custom_slice_positions, o.use_extrapolation, o.shard_index, o.shard_count) == (0, -1, 1, 1, 0, 0, 1, 0, 1)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    rc = L.lib().mci_create(ctypes.byref(ctx), 0)
    assert rc != 0 and not ctx
    from itkcontourinterpolation_b200 import MorphologicalContourInterpolator
    with pytest.raises(L.MciError):
        MorphologicalContourInterpolator()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "itkcontourinterpolation_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synthetic.py", f


def test_entry_points_reject_bad_arguments_without_touching_a_device():
    # status codes, never exceptions or crashes ("All functions return int status (0 = ok), never throw")
    lib = L.lib()
    opt = L.FilterOptions()
    lib.mci_filter_default_options(ctypes.byref(opt))
    dims = (ctypes.c_int64 * 3)(4, 4, 4)
    buf = (ctypes.c_uint16 * 64)()
    assert lib.mci_filter_run(None, ctypes.byref(opt), buf, buf, dims, L.U16, None, 0, None) != 0
    assert lib.mci_filter_run_resident(None, ctypes.byref(opt), None, 0, None) != 0
    assert lib.mci_filter_run_batch(None, 1, ctypes.byref(opt), None, None, 0, dims, L.U16, None) != 0
    one = (ctypes.c_void_p * 1)(None)
    assert lib.mci_filter_run_batch(one, 1, ctypes.byref(opt), None, None, 0, dims, L.U16, None) != 0   # null context in the list
    assert lib.mci_median_filter_run(None, buf, buf, dims, L.U16, 1) != 0
    assert lib.mci_median_filter_resident(None, 1) != 0
    assert lib.mci_set_wait_mode(None, 1) != 0
    assert lib.mci_synchronize(None) != 0
    assert lib.mci_upload_volume(None, buf, dims, L.U16) != 0
    assert lib.mci_filter_schedule_pairs(None, 1, 1, None, 0, None, None, 0, None) != 0
    assert lib.mci_last_error(None) == b"null context"
    lib.mci_destroy(None)       # no-op
    lib.mci_host_free(None)     # no-op


def test_cpp_header_declares_the_reference_interface():
    # every public method of the reference class (H:85-224) exists in the C++ mirror
    text = open(os.path.join(ROOT, "include", "mci_b200_filter.hpp")).read()
    for name in ("SetLabel", "GetLabel", "SetAxis", "GetAxis", "SetHeuristicAlignment", "GetHeuristicAlignment",
                 "SetUseDistanceTransform", "GetUseDistanceTransform", "SetUseCustomSlicePositions", "GetUseCustomSlicePositions",
                 "SetUseExtrapolation", "SetUseBallStructuringElement", "GetUseBallStructuringElement", "DetermineSliceOrientations",
                 "ClearLabeledSliceIndices", "SetLabeledSliceIndices", "GetLabeledSliceIndices", "SliceSetType", "SliceIndicesType",
                 "New()", "Update()", "GetOutput()", "SetInput"):
        if name.startswith(("Set", "Get")) and name.endswith(("Label", "Axis", "Alignment", "Transform", "Positions", "Element")) \
                and name not in ("SetUseExtrapolation",):
            stem = name[3:]
            assert ("MCI_B200_SET_GET(%s," % stem) in text or name in text, name
        else:
            assert name in text, name
