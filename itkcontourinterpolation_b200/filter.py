"""Python mirror of itk::MorphologicalContourInterpolator for 3-D label volumes (numpy [z, y, x]).

Same setters, defaults and execution contract as the reference filter
(include/itkMorphologicalContourInterpolator.h:85-235), running the slice-pair interpolation path on
a B200 through the C ABI of include/mci_b200.h.  No CPU fallback: constructing the filter without a
usable CUDA device raises.
"""
import ctypes

import numpy as np

from . import _lib as L

_PTYPE = {np.dtype(np.uint8): L.U8, np.dtype(np.uint16): L.U16, np.dtype(np.int16): L.I16}


class MorphologicalContourInterpolator:
    def __init__(self, device=0):
        self._lib = L.lib()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.mci_create(ctypes.byref(self._ctx), int(device))
        if rc != L.OK:
            raise L.MciError(rc, "mci_create failed (no usable CUDA device; there is no CPU fallback)")
        self._opt = L.FilterOptions()
        self._lib.mci_filter_default_options(ctypes.byref(self._opt))
        self._custom = {}  # (axis, label) -> sorted slice indices
        self._input = None
        self._output = None
        self.last_stats = None

    def close(self):
        if self._ctx:
            self._lib.mci_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- itkSetMacro / itkGetMacro mirrors (H:85-165) ----
    def SetLabel(self, v): self._opt.label = int(v)
    def GetLabel(self): return int(self._opt.label)
    def SetAxis(self, v): self._opt.axis = int(v)
    def GetAxis(self): return int(self._opt.axis)
    def SetHeuristicAlignment(self, v): self._opt.heuristic_alignment = int(bool(v))
    def GetHeuristicAlignment(self): return bool(self._opt.heuristic_alignment)
    def SetUseDistanceTransform(self, v): self._opt.use_distance_transform = int(bool(v))
    def GetUseDistanceTransform(self): return bool(self._opt.use_distance_transform)
    def SetUseCustomSlicePositions(self, v): self._opt.use_custom_slice_positions = int(bool(v))
    def GetUseCustomSlicePositions(self): return bool(self._opt.use_custom_slice_positions)
    def SetUseExtrapolation(self, v): self._opt.use_extrapolation = int(bool(v))
    def SetUseBallStructuringElement(self, v): self._opt.use_ball_structuring_element = int(bool(v))
    def GetUseBallStructuringElement(self): return bool(self._opt.use_ball_structuring_element)

    # ---- slice-index API (H:171-224) ----
    def ClearLabeledSliceIndices(self):
        self._custom = {}

    def SetLabeledSliceIndices(self, axis, label, indices):
        self._custom[(int(axis), int(label))] = sorted(set(int(i) for i in indices))

    def GetLabeledSliceIndices(self, axis=None, label=None):
        """All indices as {axis: {label: [slices]}}; or the list for one (axis, label)."""
        if self._opt.use_custom_slice_positions or not self._detected:
            table = {}
            for (a, l), s in self._custom.items():
                table.setdefault(a, {})[l] = list(s)
        else:
            table = self._detected
        if axis is None:
            return table
        return list(table.get(int(axis), {}).get(int(label), []))

    _detected = None

    def DetermineSliceOrientations(self):
        """Slice detection only (X:128-201); fills GetLabeledSliceIndices()."""
        vol = self._require_input()
        self._upload(vol)
        n = ctypes.c_int64(0)
        self._check(self._lib.mci_filter_detect(self._ctx, ctypes.byref(self._opt), ctypes.byref(n)))
        self._detected = self._fetch_labeled_slices(n.value)
        return self._detected

    # ---- pipeline ----
    def SetInput(self, vol):
        vol = np.ascontiguousarray(vol)
        if vol.ndim != 3 or vol.dtype not in _PTYPE:
            raise TypeError("expected a 3-D uint8/uint16/int16 label volume indexed [z, y, x]")
        self._input = vol
        self._output = None
        self._detected = None  # slice sets of a previous input say nothing about this one

    def Update(self):
        vol = self._require_input()
        out = np.empty_like(vol)
        dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
        cust, ncust = None, 0
        if self._opt.use_custom_slice_positions:
            rows = [(a, l, s) for (a, l), ss in sorted(self._custom.items()) for s in ss]
            cust = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
            ncust = cust.shape[0]
        st = L.FilterStats()
        rc = self._lib.mci_filter_run(self._ctx, ctypes.byref(self._opt), vol.ctypes.data, out.ctypes.data, dims,
                                      _PTYPE[vol.dtype], cust.ctypes.data if ncust else None, ncust, ctypes.byref(st))
        self._check(rc)
        self.last_stats = st.as_dict()
        self._output = out
        if not self._opt.use_custom_slice_positions:
            # like the reference, whose GenerateData leaves m_LabeledSlices filled (X:1573-1576): what the run detected
            self._detected = self._fetch_labeled_slices(int(st.n_annotated_slices))
        return out

    def _fetch_labeled_slices(self, n):
        trip = np.zeros((max(n, 1), 3), np.int64)
        self._check(self._lib.mci_filter_get_labeled_slices(self._ctx, trip.ctypes.data))
        table = {}
        for a, l, s in trip[:n]:
            table.setdefault(int(a), {}).setdefault(int(l), []).append(int(s))
        return table

    def GetOutput(self):
        if self._output is None:
            self.Update()
        return self._output

    # ---- helpers ----
    def _require_input(self):
        if self._input is None:
            raise RuntimeError("SetInput() has not been called")
        return self._input

    def _upload(self, vol):
        dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
        self._check(self._lib.mci_upload_volume(self._ctx, vol.ctypes.data, dims, _PTYPE[vol.dtype]))

    def _check(self, rc):
        if rc != L.OK:
            raise L.MciError(rc, self._lib.mci_last_error(self._ctx).decode())


def morphological_contour_interpolator(vol, label=0, axis=-1, heuristic_alignment=True, use_distance_transform=True,
                                       use_ball_structuring_element=False, use_extrapolation=True, device=0,
                                       return_stats=False):
    """Functional form, like itk.morphological_contour_interpolator(image, ...) (examples/mciExample.py:9)."""
    f = MorphologicalContourInterpolator(device)
    try:
        f.SetLabel(label)
        f.SetAxis(axis)
        f.SetHeuristicAlignment(heuristic_alignment)
        f.SetUseDistanceTransform(use_distance_transform)
        f.SetUseBallStructuringElement(use_ball_structuring_element)
        f.SetUseExtrapolation(use_extrapolation)
        f.SetInput(vol)
        out = f.Update()
        return (out, f.last_stats) if return_stats else out
    finally:
        f.close()


class VolumeSeriesInterpolator:
    """Throughput mode for a series of independent label volumes of one shape and pixel type
    (mci_filter_run_batch): `depth` contexts -- own stream, own host thread, own pinned in/out buffers each -- so that
    the H2D copy of one volume, the kernels of another and the D2H copy of a third overlap.  Every volume gets exactly
    the result MorphologicalContourInterpolator.Update() gives it."""

    def __init__(self, shape, dtype=np.uint16, depth=3, device=0, **options):
        self._lib = L.lib()
        self._shape = tuple(int(v) for v in shape)
        self._dtype = np.dtype(dtype)
        if len(self._shape) != 3 or self._dtype not in _PTYPE:
            raise TypeError("expected a 3-D uint8/uint16/int16 volume shape [z, y, x]")
        self._depth = int(depth)
        self._ctxs, self._in, self._out = [], [], []
        self._opt = L.FilterOptions()
        self._lib.mci_filter_default_options(ctypes.byref(self._opt))
        for k, v in options.items():
            if not hasattr(self._opt, k) or k.startswith("shard"):
                raise TypeError("unknown filter option %r" % k)
            setattr(self._opt, k, int(v))
        self._nbytes = int(np.prod(self._shape)) * self._dtype.itemsize
        for _ in range(self._depth):
            c = ctypes.c_void_p()
            rc = self._lib.mci_create(ctypes.byref(c), int(device))
            if rc != L.OK:
                self.close()
                raise L.MciError(rc, "mci_create failed (no usable CUDA device; there is no CPU fallback)")
            self._ctxs.append(c)
            for lst in (self._in, self._out):
                p = self._lib.mci_host_alloc(self._nbytes)
                if not p:
                    self.close()
                    raise MemoryError("pinned host allocation failed")
                lst.append(p)

    def close(self):
        for p in self._in + self._out:
            self._lib.mci_host_free(p)
        for c in self._ctxs:
            self._lib.mci_destroy(c)
        self._ctxs, self._in, self._out = [], [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, volumes):
        """Interpolates every volume of the sequence; returns the list of outputs (numpy arrays)."""
        volumes = [np.ascontiguousarray(v) for v in volumes]
        for v in volumes:
            if v.shape != self._shape or v.dtype != self._dtype:
                raise TypeError("every volume must have shape %r and dtype %s" % (self._shape, self._dtype))
        dims = (ctypes.c_int64 * 3)(self._shape[2], self._shape[1], self._shape[0])
        outs = []
        d = self._depth
        ctx_arr = (ctypes.c_void_p * d)(*[c.value for c in self._ctxs])
        for g0 in range(0, len(volumes), d):  # one group of `depth` volumes per call: the pinned slots are reused
            group = volumes[g0:g0 + d]
            for k, v in enumerate(group):
                ctypes.memmove(self._in[k], v.ctypes.data, self._nbytes)
            n = len(group)
            in_arr = (ctypes.c_void_p * n)(*self._in[:n])
            out_arr = (ctypes.c_void_p * n)(*self._out[:n])
            rc = self._lib.mci_filter_run_batch(ctx_arr, d, ctypes.byref(self._opt), in_arr, out_arr, n, dims,
                                                _PTYPE[self._dtype], None)
            if rc != L.OK:
                msgs = [self._lib.mci_last_error(c).decode() for c in self._ctxs]
                raise L.MciError(rc, "; ".join(m for m in msgs if m))
            for k in range(n):
                buf = (ctypes.c_char * self._nbytes).from_address(self._out[k])
                outs.append(np.frombuffer(buf, dtype=self._dtype).reshape(self._shape).copy())
        return outs


def _dims(vol):
    return (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])


def _checked_volume(vol):
    vol = np.ascontiguousarray(vol)
    if vol.ndim != 3 or vol.dtype not in _PTYPE:
        raise TypeError("expected a 3-D uint8/uint16/int16 label volume indexed [z, y, x]")
    return vol


def median_image_filter(vol, radius=1, device=0):
    """itk::MedianImageFilter with SetRadius(radius) -- the label smoothing the reference's example applies to the
    interpolator's output (examples/mciExample.cxx:29-39): box neighbourhood (2r+1)^3, ZeroFluxNeumann boundary."""
    vol = _checked_volume(vol)
    f = MorphologicalContourInterpolator(device)
    try:
        out = np.empty_like(vol)
        f._check(f._lib.mci_median_filter_run(f._ctx, vol.ctypes.data, out.ctypes.data, _dims(vol), _PTYPE[vol.dtype], int(radius)))
        return out
    finally:
        f.close()


def interpolate_and_smooth(vol, smoothing_radius=2, device=0, **options):
    """The whole of examples/mciExample.cxx: interpolate with the filter's settings (keyword names of
    mci_filter_options, e.g. axis=2, use_distance_transform=0), then median-smooth with `smoothing_radius` -- the
    interpolated volume stays in HBM between the two steps."""
    vol = _checked_volume(vol)
    f = MorphologicalContourInterpolator(device)
    try:
        for k, v in options.items():
            if not hasattr(f._opt, k) or k.startswith("shard"):
                raise TypeError("unknown filter option %r" % k)
            setattr(f._opt, k, int(v))
        f._upload(vol)
        st = L.FilterStats()
        f._check(f._lib.mci_filter_run_resident(f._ctx, ctypes.byref(f._opt), None, 0, ctypes.byref(st)))
        f._check(f._lib.mci_median_filter_resident(f._ctx, int(smoothing_radius)))
        out = np.empty_like(vol)
        f._check(f._lib.mci_download_volume(f._ctx, out.ctypes.data))
        return out
    finally:
        f.close()
