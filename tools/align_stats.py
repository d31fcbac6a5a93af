"""Diagnostics: per-job cost of the Align kernel on a synthetic config (run on the GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
vol, _, cfg = make_config(name, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt))
opt.axis = cfg["axis"]; opt.use_distance_transform = int(cfg["use_dt"]); opt.use_ball_structuring_element = int(cfg["ball"])
out = np.empty_like(vol); dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats()
for _ in range(2):
    rc = lib.mci_filter_run(ctx, ctypes.byref(opt), vol.ctypes.data, out.ctypes.data, dims, L.U16, None, 0, ctypes.byref(st))
    assert rc == 0, lib.mci_last_error(ctx)
n = st.n_jobs
a = np.zeros((n, 4), np.int32); assert lib.mci_get_align_stats(ctx, a.ctypes.data) == 0
order = np.argsort(-a[:, 2])
print("jobs", n, "total pops", a[:, 0].sum(), "max cycles", a[:, 2].max(), "median cycles", int(np.median(a[:, 2])))
print("pops waves cycles kind*1000+nB")
for k in order[:15]: print(a[k])
print("by kind:", {int(k): (int((a[:,3]//1000==k).sum()), int(a[a[:,3]//1000==k][:,2].mean()) if (a[:,3]//1000==k).any() else 0) for k in (0,1,2)})
