"""ctypes binding of libmci_b200.so (include/mci_b200.h).  Fails loudly: there is no CPU fallback."""
import ctypes
import os

from . import build as _build

_lib = None

OK = 0
U8, U16, I16 = 0, 1, 2
JOB_ONE_TO_ONE, JOB_ONE_TO_N, JOB_EXTRAPOLATE = 0, 1, 2


class MciError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mci_b200 error %d: %s" % (code, msg))
        self.code = code


class LabelInfo(ctypes.Structure):
    _fields_ = [("label", ctypes.c_int64), ("bbox_index", ctypes.c_int32 * 3), ("bbox_size", ctypes.c_int32 * 3)]


class LabeledSlice(ctypes.Structure):
    _fields_ = [("axis", ctypes.c_int32), ("slice", ctypes.c_int32), ("label", ctypes.c_int64)]


class SliceDesc(ctypes.Structure):
    _fields_ = [("axis", ctypes.c_int32), ("slice", ctypes.c_int32), ("label", ctypes.c_int64),
                ("u0", ctypes.c_int32), ("v0", ctypes.c_int32), ("w", ctypes.c_int32), ("h", ctypes.c_int32)]


class ComponentInfo(ctypes.Structure):
    _fields_ = [("count", ctypes.c_int64), ("sum_u", ctypes.c_int64), ("sum_v", ctypes.c_int64),
                ("min_u", ctypes.c_int32), ("min_v", ctypes.c_int32), ("max_u", ctypes.c_int32), ("max_v", ctypes.c_int32)]


class SegmentDesc(ctypes.Structure):
    _fields_ = [("slice_i", ctypes.c_int32), ("slice_j", ctypes.c_int32)]


class Pair(ctypes.Structure):
    _fields_ = [("segment", ctypes.c_int32), ("i_id", ctypes.c_uint16), ("j_id", ctypes.c_uint16)]


class Job(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("slice_a", ctypes.c_int32), ("a_id", ctypes.c_int32),
                ("slice_b", ctypes.c_int32), ("n_b", ctypes.c_int32), ("b_ids_offset", ctypes.c_int32),
                ("pos_a", ctypes.c_int32), ("pos_b", ctypes.c_int32)]


class Params(ctypes.Structure):
    _fields_ = [("heuristic_alignment", ctypes.c_int32), ("use_distance_transform", ctypes.c_int32),
                ("use_ball", ctypes.c_int32), ("min_align_iters", ctypes.c_int32), ("max_align_iters", ctypes.c_int32)]


class FilterOptions(ctypes.Structure):
    _fields_ = [("label", ctypes.c_int64), ("axis", ctypes.c_int32), ("heuristic_alignment", ctypes.c_int32),
                ("use_distance_transform", ctypes.c_int32), ("use_ball_structuring_element", ctypes.c_int32),
                ("use_custom_slice_positions", ctypes.c_int32), ("use_extrapolation", ctypes.c_int32),
                ("shard_index", ctypes.c_int32), ("shard_count", ctypes.c_int32)]


class KernelTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("launches", ctypes.c_int64), ("total_ms", ctypes.c_double)]


class FilterStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("n_labels", "n_annotated_slices", "n_segments", "n_pairs", "n_jobs",
                                               "n_components", "n_launches")] + \
               [(n, ctypes.c_double) for n in ("ms_upload", "ms_detect", "ms_components", "ms_pairs", "ms_interpolate",
                                               "ms_download", "ms_total")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/mci_b200.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
_I = ctypes.c_int
SYMBOLS = [
    ("mci_create", _I, [ctypes.POINTER(_P), _I]),
    ("mci_destroy", None, [_P]),
    ("mci_last_error", ctypes.c_char_p, [_P]),
    ("mci_host_alloc", _P, [ctypes.c_size_t]),
    ("mci_host_free", None, [_P]),
    ("mci_launch_count", ctypes.c_int64, [_P]),
    ("mci_stream", _P, [_P]),
    ("mci_synchronize", _I, [_P]),
    ("mci_set_wait_mode", _I, [_P, _I]),
    ("mci_profile_enable", _I, [_P, _I]),
    ("mci_profile_read", _I, [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    ("mci_upload_volume", _I, [_P, _P, ctypes.POINTER(ctypes.c_int64), _I]),
    ("mci_attach_device_volume", _I, [_P, _P, _P, ctypes.POINTER(ctypes.c_int64), _I]),
    ("mci_download_volume", _I, [_P, _P]),
    ("mci_device_output", _P, [_P]),
    ("mci_volume_dims", _I, [_P, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_I)]),
    ("mci_detect_slices", _I, [_P, _I, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    ("mci_get_detection", _I, [_P, _P, _P]),
    ("mci_label_slices", _I, [_P, ctypes.c_int32, _P, _I, _P]),
    ("mci_get_components", _I, [_P, _P]),
    ("mci_get_component_image", _I, [_P, ctypes.c_int32, _P]),
    ("mci_match_pairs", _I, [_P, ctypes.c_int32, _P, ctypes.POINTER(ctypes.c_int32)]),
    ("mci_get_pairs", _I, [_P, _P]),
    ("mci_label_and_match", _I, [_P, ctypes.c_int32, _P, _I, ctypes.c_int32, _P, _P, ctypes.POINTER(ctypes.c_int32)]),
    ("mci_interpolate", _I, [_P, ctypes.c_int32, _P, _P, ctypes.c_int32, ctypes.POINTER(Params)]),
    ("mci_get_translations", _I, [_P, _P]),
    ("mci_get_align_stats", _I, [_P, _P]),
    ("mci_combine_max", _I, [_P, _P, _P, ctypes.c_int64, _I]),
    ("mci_get_emitted", _I, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int64)]),
    ("mci_copy_emitted", _I, [_P, _P, ctypes.c_int64]),
    ("mci_apply_masks", _I, [_P, _P, ctypes.c_int64, _P, ctypes.c_int64]),
    ("mci_median_filter_resident", _I, [_P, _I]),
    ("mci_median_filter_run", _I, [_P, _P, _P, ctypes.POINTER(ctypes.c_int64), _I, _I]),
    ("mci_filter_default_options", None, [ctypes.POINTER(FilterOptions)]),
    ("mci_filter_run", _I, [_P, ctypes.POINTER(FilterOptions), _P, _P, ctypes.POINTER(ctypes.c_int64), _I, _P,
                            ctypes.c_int64, ctypes.POINTER(FilterStats)]),
    ("mci_filter_run_resident", _I, [_P, ctypes.POINTER(FilterOptions), _P, ctypes.c_int64, ctypes.POINTER(FilterStats)]),
    ("mci_filter_run_with_detection", _I, [_P, ctypes.POINTER(FilterOptions), _P, ctypes.c_int32, _P, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(FilterStats)]),
    ("mci_copy_input_range", _I, [_P, ctypes.c_int64, ctypes.c_int64]),
    ("mci_filter_run_batch", _I, [_P, ctypes.c_int32, ctypes.POINTER(FilterOptions), _P, _P, ctypes.c_int32,
                                  ctypes.POINTER(ctypes.c_int64), _I, _P]),
    ("mci_filter_detect", _I, [_P, ctypes.POINTER(FilterOptions), ctypes.POINTER(ctypes.c_int64)]),
    ("mci_filter_get_labeled_slices", _I, [_P, _P]),
    ("mci_filter_schedule_pairs", _I, [_P, ctypes.c_int32, _I, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), _P,
                                       ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    ("mci_filter_forget", None, [_P]),
]


def library_path():
    return _build.LIB


def lib():
    """Load (building first if the sources are newer) the shared library."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the ABI and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
