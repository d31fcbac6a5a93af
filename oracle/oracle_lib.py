"""ctypes loader for the CPU oracle (oracle/mci_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmci_oracle.so")

PTYPES = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.int16): 2}


class Params(ctypes.Structure):
    _fields_ = [
        ("label", ctypes.c_int64),
        ("axis", ctypes.c_int32),
        ("heuristic_alignment", ctypes.c_int32),
        ("use_distance_transform", ctypes.c_int32),
        ("use_ball", ctypes.c_int32),
        ("use_extrapolation", ctypes.c_int32),
        ("use_custom_slice_positions", ctypes.c_int32),
        ("threads", ctypes.c_int32),
        ("shard_index", ctypes.c_int32),
        ("shard_count", ctypes.c_int32),
    ]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("segments", "root_jobs", "nodes", "align_calls", "align_evals", "extrapolations", "one_to_n")] + \
               [(n, ctypes.c_double) for n in ("t_detect", "t_cc", "t_interp", "t_final", "t_total")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    srcs = [os.path.join(_HERE, n) for n in ("mci_oracle.cpp", "median_filter_oracle.cpp")]
    have_src = all(os.path.exists(s) for s in srcs)
    if force or not os.path.exists(_SO) or (have_src and os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.mci_oracle_run.restype = ctypes.c_int
        L.mci_oracle_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                                     ctypes.POINTER(Params), ctypes.POINTER(Stats), ctypes.c_void_p, ctypes.c_int64]
        L.mci_oracle_last_error.restype = ctypes.c_char_p
        L.mci_oracle_detect.restype = ctypes.c_int64
        L.mci_oracle_detect.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_int64)]
        L.mci_oracle_cc2d.restype = ctypes.c_int64
        L.mci_oracle_cc2d.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.mci_oracle_edt2.restype = None
        L.mci_oracle_edt2.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.mci_oracle_dilate.restype = None
        L.mci_oracle_dilate.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
        L.mci_oracle_median.restype = ctypes.c_int
        L.mci_oracle_median.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.mci_oracle_median_image_filter.restype = ctypes.c_int
        L.mci_oracle_median_image_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def interpolate(vol, label=0, axis=-1, heuristic_alignment=True, use_distance_transform=True, use_ball=False,
                use_extrapolation=True, custom_slices=None, threads=1, return_stats=False, shard_index=0, shard_count=1):
    """Run the restated filter on `vol` (numpy array indexed [z, y, x], x fastest)."""
    vol = np.ascontiguousarray(vol)
    assert vol.ndim == 3 and vol.dtype in PTYPES
    out = np.empty_like(vol)
    dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    p = Params(int(label), int(axis), int(heuristic_alignment), int(use_distance_transform), int(use_ball),
               int(use_extrapolation), int(custom_slices is not None), int(threads), int(shard_index), int(shard_count))
    st = Stats()
    cust = None
    ncust = 0
    if custom_slices is not None:
        cust = np.ascontiguousarray(np.asarray(custom_slices, dtype=np.int64).reshape(-1, 3))
        ncust = cust.shape[0]
    rc = lib().mci_oracle_run(vol.ctypes.data, out.ctypes.data, dims, PTYPES[vol.dtype], ctypes.byref(p), ctypes.byref(st),
                              cust.ctypes.data if cust is not None and ncust else None, ncust)
    if rc != 0:
        raise OracleError(lib().mci_oracle_last_error().decode())
    return (out, st.as_dict()) if return_stats else out


def detect(vol, axis=-1):
    """DetermineSliceOrientations only: (triplets[axis,label,slice], boxes[label, idx3, size3])."""
    vol = np.ascontiguousarray(vol, dtype=np.uint16)
    dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    cap = 1 << 20
    trip = np.zeros((cap, 3), np.int64)
    boxes = np.zeros((70000, 7), np.int64)
    nb = ctypes.c_int64(0)
    n = lib().mci_oracle_detect(vol.ctypes.data, dims, 1, axis, trip.ctypes.data, cap, boxes.ctypes.data, 70000, ctypes.byref(nb))
    assert 0 <= n <= cap
    return trip[:n].copy(), boxes[:nb.value].copy()


def cc2d(mask, fully=False):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    ids = np.zeros(m.shape, np.uint16)
    n = lib().mci_oracle_cc2d(m.ctypes.data, m.shape[1], m.shape[0], int(fully), ids.ctypes.data)
    return ids, n


def edt2(mask, brute=False):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    d2 = np.zeros(m.shape, np.int64)
    lib().mci_oracle_edt2(m.ctypes.data, m.shape[1], m.shape[0], int(brute), d2.ctypes.data)
    return d2


def dilate(mask, ball=False):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    o = np.zeros_like(m)
    lib().mci_oracle_dilate(m.ctypes.data, m.shape[1], m.shape[0], int(ball), o.ctypes.data)
    return o


def median(i_mask, j_mask, dtype=np.uint16, use_distance_transform=True, use_ball=False):
    a = np.ascontiguousarray(i_mask, dtype=np.uint8)
    b = np.ascontiguousarray(j_mask, dtype=np.uint8)
    o = np.zeros_like(a)
    rc = lib().mci_oracle_median(a.ctypes.data, b.ctypes.data, a.shape[1], a.shape[0], PTYPES[np.dtype(dtype)],
                                 int(use_distance_transform), int(use_ball), o.ctypes.data)
    if rc != 0:
        raise OracleError(lib().mci_oracle_last_error().decode())
    return o


def median_image_filter(vol, radius=1, threads=1):
    """itk::MedianImageFilter with SetRadius(radius), as examples/mciExample.cxx:29-39 applies it (restated)."""
    vol = np.ascontiguousarray(vol)
    assert vol.ndim == 3 and vol.dtype in PTYPES
    out = np.empty_like(vol)
    dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    rc = lib().mci_oracle_median_image_filter(vol.ctypes.data, out.ctypes.data, dims, PTYPES[vol.dtype], int(radius), int(threads))
    if rc != 0:
        raise OracleError("median_image_filter: bad arguments (%d)" % rc)
    return out
