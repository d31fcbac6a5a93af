"""Run one synthetic config on the GPU, check against the oracle (optional) and print timings."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itkcontourinterpolation_b200 import _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
name = sys.argv[1]; scale = float(sys.argv[2]); check = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t0 = time.time(); vol, _, cfg = make_config(name, scale); print("generated", vol.shape, "%.1fs" % (time.time() - t0), flush=True)
lib = L.lib(); ctx = ctypes.c_void_p(); assert lib.mci_create(ctypes.byref(ctx), 0) == 0
opt = L.FilterOptions(); lib.mci_filter_default_options(ctypes.byref(opt))
opt.axis = cfg["axis"]; opt.use_distance_transform = int(cfg["use_dt"]); opt.use_ball_structuring_element = int(cfg["ball"])
dims = (ctypes.c_int64 * 3)(*cfg["dims"]); st = L.FilterStats(); out = np.empty_like(vol)
for it in range(3):
    t0 = time.time()
    rc = lib.mci_filter_run(ctx, ctypes.byref(opt), vol.ctypes.data, out.ctypes.data, dims, L.U16, None, 0, ctypes.byref(st))
    assert rc == 0, lib.mci_last_error(ctx)
    print("run %d: %.1f ms wall" % (it, (time.time() - t0) * 1e3), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.as_dict().items()}, flush=True)
assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
lib.mci_profile_enable(ctx, 1)
for it in range(3):
    rc = lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)); assert rc == 0, lib.mci_last_error(ctx)
print("resident: %.2f ms" % st.ms_total, "Mvoxel/s %.0f" % (vol.size / st.ms_total / 1e3))
kt = (L.KernelTime * 64)(); nk = ctypes.c_int32(0); lib.mci_profile_read(ctx, kt, 64, ctypes.byref(nk))
for k in sorted(range(nk.value), key=lambda k: -kt[k].total_ms)[:24]: print("  %-18s %3d launches %.3f ms/step" % (kt[k].name.decode(), kt[k].launches // 3, kt[k].total_ms / 3))
assert np.array_equal(out[vol != 0], vol[vol != 0]); print("interpolated voxels:", int((out != vol).sum()))
if check:
    from oracle import oracle_lib as O
    t0 = time.time(); ref, ost = O.interpolate(vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"], threads=os.cpu_count(), return_stats=True)
    print("oracle %.2fs (%d threads)" % (time.time() - t0, os.cpu_count()), {k: ost[k] for k in ("segments", "root_jobs", "nodes", "extrapolations", "one_to_n")})
    print("BIT-EXACT" if np.array_equal(ref, out) else "MISMATCH %d voxels" % int((ref != out).sum()))
