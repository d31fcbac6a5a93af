"""Diagnostics: host-side phase times of mci_filter_run_resident (run on the GPU box)."""
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
dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats()
assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
acc = {}
N = 20
for it in range(N + 3):
    rc = lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)); assert rc == 0, lib.mci_last_error(ctx)
    if it >= 3:
        for k, v in st.as_dict().items():
            if k.startswith("ms_"): acc[k] = acc.get(k, 0) + v / N
print({k: round(v, 3) for k, v in acc.items()}, {k: v for k, v in st.as_dict().items() if k.startswith("n_")})
