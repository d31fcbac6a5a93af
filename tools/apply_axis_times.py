"""Diagnostics: per-axis kernel times of a multi-axis config (which axis pays for the strided volume writes)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
vol, _, cfg = make_config(name, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt))
opt.use_distance_transform = int(cfg["use_dt"]); opt.use_ball_structuring_element = int(cfg["ball"])
dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats()
assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
for axis in (-1, 0, 1, 2):
    opt.axis = axis
    for it in range(5):
        if it == 2:
            lib.mci_profile_enable(ctx, 1)
        rc = lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)); assert rc == 0, lib.mci_last_error(ctx)
    kt = (L.KernelTime * 64)(); nk = ctypes.c_int32(0); lib.mci_profile_read(ctx, kt, 64, ctypes.byref(nk))
    lib.mci_profile_enable(ctx, 0)
    top = sorted(((kt[k].name.decode(), kt[k].total_ms / 3) for k in range(nk.value)), key=lambda t: -t[1])[:6]
    print("axis %2d: %.2f ms/step, %d nodes/jobs=%d  " % (axis, st.ms_total, st.n_segments, st.n_jobs), ", ".join("%s %.2f" % t for t in top), flush=True)
