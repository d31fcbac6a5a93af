"""Label smoothing step (examples/mciExample.cxx:29-39) on the interpolated workload volume: kernel time from CUDA
events on the library's stream, HBM fraction against MEASURED_PEAKS.json, the oracle (restated itk::MedianImageFilter,
all host threads) beside it.  Usage: python tools/median_bench.py [C3|C4|C5] [scale] [radius] [check]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
radius = int(sys.argv[3]) if len(sys.argv) > 3 else 2
check = int(sys.argv[4]) if len(sys.argv) > 4 else 1
vol, _, cfg = make_config(name, scale)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt))
opt.axis = cfg["axis"]; opt.use_distance_transform = int(cfg["use_dt"]); opt.use_ball_structuring_element = int(cfg["ball"])
dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats()
assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
N = 10
for it in range(N + 3):
    if it == 3:
        lib.mci_profile_enable(ctx, 1)
    rc = lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)); assert rc == 0, lib.mci_last_error(ctx)
    rc = lib.mci_median_filter_resident(ctx, radius); assert rc == 0, lib.mci_last_error(ctx)
    assert lib.mci_synchronize(ctx) == 0
kt = (L.KernelTime * 64)(); nk = ctypes.c_int32(0); lib.mci_profile_read(ctx, kt, 64, ctypes.byref(nk))
times = {kt[k].name.decode(): kt[k].total_ms / max(kt[k].launches, 1) for k in range(nk.value)}
ms = times["k_median3d"]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
alg = 2.0 * vol.nbytes
res = {"workload": name, "scale": scale, "radius": radius, "k_median3d_ms": ms, "Mvoxel_s": vol.size / ms / 1e3,
       "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / ms / 1e6, "frac_of_measured_hbm_peak": alg / ms / 1e6 / peak}
out = np.empty_like(vol)
assert lib.mci_download_volume(ctx, out.ctypes.data) == 0
if check:
    from oracle import oracle_lib as O
    interp = O.interpolate(vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"], threads=os.cpu_count())
    t0 = time.time(); ref = O.median_image_filter(interp, radius, threads=os.cpu_count()); t = time.time() - t0
    res["oracle_s"] = t; res["oracle_threads"] = os.cpu_count(); res["oracle_Mvoxel_s"] = vol.size / t / 1e6
    res["bit_exact"] = bool(np.array_equal(ref, out)); res["changed_voxels"] = int((ref != interp).sum())
print(json.dumps(res))
