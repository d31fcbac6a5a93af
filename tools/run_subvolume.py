"""Kernel table of ONE rank's share of a slab-sharded C5 run, on one GPU: the planes [z0, z1] of rank `r` of `n`
are run as a volume of their own (same shapes, jobs and kernels as the rank sees; boundary segments aside).
    python tools/run_subvolume.py [rank] [world]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.distributed import slab_bounds
from itkcontourinterpolation_b200.synthetic import CONFIGS, make_config
r = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = CONFIGS["C5"]
z0, z1 = slab_bounds(cfg["dims"][2], n)[r]
vol, _, _ = make_config("C5", 1.0, z_range=(z0, z1 + 1), want_dense=False)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt)); opt.axis = 2
dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0]); st = L.FilterStats()
assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
for it in range(3):
    assert lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)) == 0, lib.mci_last_error(ctx)
print("resident %.2f ms" % st.ms_total, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.as_dict().items()})
lib.mci_profile_enable(ctx, 1)
for it in range(3):
    assert lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)) == 0
kt = (L.KernelTime * 64)(); nk = ctypes.c_int32(0); lib.mci_profile_read(ctx, kt, 64, ctypes.byref(nk))
for k in sorted(range(nk.value), key=lambda k: -kt[k].total_ms): print("  %-18s %3d launches %.3f ms/step" % (kt[k].name.decode(), kt[k].launches // 3, kt[k].total_ms / 3))
