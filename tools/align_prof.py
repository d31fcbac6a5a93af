import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
vol, _, cfg = make_config(sys.argv[1], 1.0, want_dense=False)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt)); opt.axis = cfg["axis"]
out = np.empty_like(vol); dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats()
for _ in range(2):
    assert lib.mci_filter_run(ctx, ctypes.byref(opt), vol.ctypes.data, out.ctypes.data, dims, L.U16, None, 0, ctypes.byref(st)) == 0
a = np.zeros((st.n_jobs, 4), np.int32); assert lib.mci_get_align_stats(ctx, a.ctypes.data) == 0
sc = (a[:, 1] & 0xFFFF) * 16; rp = ((a[:, 1] >> 16) & 0xFFFF) * 16
print("jobs", len(a), "median cycles", int(np.median(a[:, 2])), "median scoring", int(np.median(sc)), "median replay", int(np.median(rp)), "median pops", int(np.median(a[:, 0])))
o = np.argsort(-a[:, 2])[:8]
for k in o: print(a[k, 0], a[k, 2], sc[k], rp[k], a[k, 3])
