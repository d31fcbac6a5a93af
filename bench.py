#!/usr/bin/env python
"""bench.py -- throughput of the slice-pair interpolation path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3|C4|C5] [--scale S]

One "step" = one GenerateData-equivalent (slice detection, connected components, component matching,
Align, 1-to-N splits, every level of the Interpolate1to1 recursion, output composition) on one synthetic
sparse-annotated label volume.  Metric: interpolated Mvoxel/s per label-volume = X*Y*Z / time.

  value : volume already resident in HBM, result left in HBM (mci_filter_run_resident)
  e2e   : through the reference-facing calls with HOST buffers (pinned host in -> pinned host out), H2D and D2H
          copies of every volume inside the timed region: mci_filter_run_batch with --depth volumes in flight
          (throughput; the headline), and mci_filter_run one volume at a time (latency; e2e.one_volume_at_a_time)

N > 1 (torchrun, one rank per GPU): `value` = independent label volumes, one per GPU, no data-path collective (voxels
all ranks processed / max-over-ranks time; scaling "weak"); `sharded` = ONE C5 volume cut into one z-slab per GPU
(strong scaling, the north-star split), with the same volume on one GPU measured beside it.
N = 1 also reports C4 and C5 (`extra.workloads`), checked against the frozen oracle output of tests/golden/large.json.
--impl reference times the CPU restatement of the reference (oracle/, all host threads) on rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from itkcontourinterpolation_b200.synthetic import make_config  # noqa: E402

WORKLOAD_DESC = {
    "C3": "synthetic 512x512x256 uint16, 20 labels annotated every 8th z-slice, UseDistanceTransform=on, axis 2",
    "C4": "synthetic 1024x1024x512 uint16, 20 labels annotated sparsely along 3 axes, ball dilations",
    "C5": "synthetic 2048x2048x1024 uint16, 200 labels annotated every 8th z-slice, UseDistanceTransform=on, axis 2",
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs.  The region lasts tens of milliseconds, so the
    samples come from NVML in a thread (one query every millisecond); nvidia-smi -lms is the fall-back."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.stop = threading.Event()
        self.thread = None
        self.proc = None
        self.rows = []

    def _nvml_loop(self, nv, h):
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=lambda: [self.rows.append([c.strip() for c in l.split(",")]) for l in self.proc.stdout],
                                 daemon=True).start()
            except Exception:
                self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.thread:
            self.thread.join(timeout=1)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        reasons = [name for bit, name in self.REASONS if self.mask & bit]
        sm, mx = list(self.sm), self.max_mhz
        for r in self.rows:  # nvidia-smi fall-back
            try:
                sm.append(float(r[0]))
                mx = max(mx or 0.0, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active") and name not in reasons:
                    reasons.append(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_time(vol, cfg, threads, runs):
    from oracle import oracle_lib as O
    times, out = [], None
    for _ in range(runs + 1):  # first run = warm-up
        t0 = time.perf_counter()
        out = O.interpolate(vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"], threads=threads)
        times.append(time.perf_counter() - t0)
    return statistics.median(times[1:]), out


def frozen(key):
    """tests/golden/large.json: sha256 + per-plane CRC-32 of the ORACLE's output for the full-size configurations"""
    p = os.path.join(ROOT, "tests", "golden", "large.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key)


def plane_crcs(a):
    import zlib
    return [zlib.crc32(np.ascontiguousarray(a[z]).data) for z in range(a.shape[0])]


def extra_workload(lib, L, torch, local_rank, name, steps, warmup):
    """one more BASELINE.json configuration on this GPU: resident and host-to-host times, checked against the frozen
    oracle output (per-plane CRC-32)"""
    vol, _, cfg = make_config(name, 1.0, want_dense=False)
    gold = frozen(name)
    ctx = ctypes.c_void_p()
    if lib.mci_create(ctypes.byref(ctx), local_rank):
        raise SystemExit("mci_create failed")

    def check(rc):
        if rc:
            raise SystemExit("mci error %d: %s" % (rc, lib.mci_last_error(ctx).decode()))

    try:
        opt = L.FilterOptions()
        lib.mci_filter_default_options(ctypes.byref(opt))
        opt.axis = cfg["axis"]
        opt.use_distance_transform = int(cfg["use_dt"])
        opt.use_ball_structuring_element = int(cfg["ball"])
        dims = (ctypes.c_int64 * 3)(*cfg["dims"])
        nbytes = vol.nbytes
        h_in, h_out = lib.mci_host_alloc(nbytes), lib.mci_host_alloc(nbytes)
        ctypes.memmove(h_in, vol.ctypes.data, nbytes)
        stats = L.FilterStats()
        stream = torch.cuda.ExternalStream(lib.mci_stream(ctx), device=torch.device("cuda", local_rank))

        def timed(fn):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) / steps

        ms_e2e = timed(lambda: check(lib.mci_filter_run(ctx, ctypes.byref(opt), h_in, h_out, dims, L.U16, None, 0, ctypes.byref(stats))))
        got = np.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_uint16)), shape=vol.shape)
        bit_exact = None
        if gold is not None:
            bit_exact = bool(plane_crcs(got) == gold["output_plane_crc32"] and plane_crcs(vol) == gold["input_plane_crc32"])
            if not bit_exact:
                raise SystemExit("%s: GPU output differs from the frozen oracle output -- no number reported" % name)
        job = stats.as_dict()
        check(lib.mci_upload_volume(ctx, h_in, dims, L.U16))
        check(lib.mci_synchronize(ctx))
        check(lib.mci_profile_enable(ctx, 2))  # CUDA events around the HBM-streaming kernel only
        l0 = lib.mci_launch_count(ctx)
        ms_res = timed(lambda: check(lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(stats))))
        launches = (lib.mci_launch_count(ctx) - l0) // (steps + warmup)
        kt = (L.KernelTime * 8)()
        nk = ctypes.c_int32(0)
        check(lib.mci_profile_read(ctx, kt, 8, ctypes.byref(nk)))
        check(lib.mci_profile_enable(ctx, 0))
        stream_ms = None
        for k in range(min(nk.value, 8)):
            if kt[k].name.decode() == "k_copy_flag" and kt[k].launches:
                stream_ms = kt[k].total_ms / kt[k].launches
        nvox = int(vol.size)
        peak, _ = measured_peak()
        lib.mci_host_free(h_in)
        lib.mci_host_free(h_out)
        return {"workload": "%s: %s" % (name, WORKLOAD_DESC[name]), "dims": list(cfg["dims"]), "steps": steps, "warmup": warmup,
                "ms_per_step": ms_res, "value": nvox / (ms_res * 1e-3) / 1e6, "unit": "Mvoxel/s",
                "e2e": {"ms_per_step": ms_e2e, "value": nvox / (ms_e2e * 1e-3) / 1e6, "call": "mci_filter_run, one volume at a time",
                        "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
                "whole_step_frac": 2.0 * nbytes / (ms_res * 1e-3) / 1e9 / peak, "gpu_launches_per_step": int(launches),
                "streaming_kernel": None if stream_ms is None else {
                    "kernel": "k_copy_flag", "avg_launch_ms": stream_ms, "achieved": 2.0 * nbytes / (stream_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "frac": 2.0 * nbytes / (stream_ms * 1e-3) / 1e9 / peak,
                    "note": "volumes of 512 MB and more go through the TMA unit (cp.async.bulk, k_copy_flag_tma)"},
                "bit_exact": bit_exact, "checked_against": "tests/golden/large.json (oracle output, CRC-32 per z-plane)" if gold else None,
                "job": {k: job[k] for k in ("n_labels", "n_annotated_slices", "n_segments", "n_jobs")}}
    finally:
        lib.mci_destroy(ctx)


def sharded_c5(torch, dist, L, local_rank, rank, world, scale, steps, warmup, lib):
    """ONE C5 volume, one z-slab per GPU (itkcontourinterpolation_b200.distributed.SlabShardedInterpolator): every rank
    generates, uploads, interpolates and downloads only its planes (+ the annotated planes bounding the segments that
    reach into its slab).  Returns the `sharded` object of the JSON line (rank 0) or None."""
    from itkcontourinterpolation_b200.distributed import SlabShardedInterpolator, slab_bounds
    from itkcontourinterpolation_b200.synthetic import CONFIGS
    cfg = CONFIGS["C5"]
    dims = tuple(max(8, int(round(d * scale))) for d in cfg["dims"])
    X, Y, Z = dims
    z0, z1 = slab_bounds(Z, world)[rank]
    nth = cfg["nth"][2]
    w0, w1 = max(0, z0 - nth), min(Z, z1 + nth + 1)  # annotated every nth plane: the bounding planes are that close
    win, _, _ = make_config("C5", scale, z_range=(w0, w1), want_dense=False)
    gold = frozen("C5" if scale == 1.0 else ("C5_half" if scale == 0.5 else "-"))
    f = SlabShardedInterpolator(local_rank, dims, torch.uint16, margin=nth + 1)
    try:
        f.set_options(axis=2, use_distance_transform=int(cfg["use_dt"]), use_ball_structuring_element=int(cfg["ball"]))
        host = torch.from_numpy(win).pin_memory()
        out = torch.empty((z1 - z0, Y, X), dtype=torch.uint16).pin_memory()

        def barrier():
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

        def timed(fn):
            for _ in range(warmup):
                fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(f.stream)
            for _ in range(steps):
                fn()
            e1.record(f.stream)
            e1.synchronize()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) / steps

        ms_e2e = timed(lambda: f.run_host(host, w0, out))
        ok = True
        if gold is not None:
            ok = plane_crcs(out.numpy()) == gold["output_plane_crc32"][z0:z1]
        # resident layout: the slab plus one ghost plane per side (slice detection reads z +- 1), as a stencil code would
        # hold it; whatever else a rank needs (annotated planes bounding segments that cross its slab) comes over NVLink
        f.load_slab(host[z0 - w0:z1 - w0], host[z0 - 1 - w0] if z0 > 0 else None, host[z1 - w0] if z1 < Z else None)
        f.nvlink_bytes = f.table_bytes = 0
        ms_res = timed(f.run_resident)
        per_step = lambda v: v / float(steps + warmup)
        nv, tb = per_step(f.nvlink_bytes), per_step(f.table_bytes)
        if gold is not None:
            ok = ok and plane_crcs(f.output_slab().cpu().numpy()) == gold["output_plane_crc32"][z0:z1]
        flags = torch.tensor([1 if ok else 0, int(nv), int(tb), f.last_stats.get("n_segments", 0)], device="cuda", dtype=torch.int64)
        allf = torch.empty((world, 4), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allf.view(-1), flags)
        allf = allf.cpu().numpy()
    finally:
        f.close()
    del host, out, win
    # the same volume on ONE GPU, measured by rank 0 in the same run (the others wait)
    one = None
    if rank == 0:
        vol, _, _ = make_config("C5", scale, want_dense=False)
        ctx = ctypes.c_void_p()
        if lib.mci_create(ctypes.byref(ctx), local_rank):
            raise SystemExit("mci_create failed")
        opt = L.FilterOptions()
        lib.mci_filter_default_options(ctypes.byref(opt))
        opt.axis = 2
        opt.use_distance_transform = int(cfg["use_dt"])
        opt.use_ball_structuring_element = int(cfg["ball"])
        cd = (ctypes.c_int64 * 3)(*dims)
        h_in, h_out = lib.mci_host_alloc(vol.nbytes), lib.mci_host_alloc(vol.nbytes)
        ctypes.memmove(h_in, vol.ctypes.data, vol.nbytes)
        st = L.FilterStats()
        stream = torch.cuda.ExternalStream(lib.mci_stream(ctx), device=torch.device("cuda", local_rank))

        def t1(fn):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1) / steps

        def must(rc):
            if rc:
                raise SystemExit("mci error %d: %s" % (rc, lib.mci_last_error(ctx).decode()))

        one_e2e = t1(lambda: must(lib.mci_filter_run(ctx, ctypes.byref(opt), h_in, h_out, cd, L.U16, None, 0, ctypes.byref(st))))
        must(lib.mci_upload_volume(ctx, h_in, cd, L.U16))
        must(lib.mci_synchronize(ctx))
        one_res = t1(lambda: must(lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st))))
        one = (one_res, one_e2e, int(st.n_segments))
        lib.mci_host_free(h_in)
        lib.mci_host_free(h_out)
        lib.mci_destroy(ctx)
    dist.barrier()
    if rank != 0:
        return None
    nvox = X * Y * Z
    bit_exact = bool(allf[:, 0].min() == 1) if gold is not None else None
    if bit_exact is False:
        raise SystemExit("sharded C5: some rank's planes differ from the frozen oracle output -- no number reported")
    return {"workload": "C5: %s" % WORKLOAD_DESC["C5"] + ("" if scale == 1.0 else " (scaled by %g)" % scale), "dims": list(dims),
            "partition": "one z-slab of %d planes per GPU; segments assigned by the slab their mid slices fall in; every output "
                         "voxel is produced by exactly one rank (no combine step)" % (Z // world),
            "steps": steps, "warmup": warmup, "scaling": "strong",
            "ms_per_volume": ms_res, "value": nvox / (ms_res * 1e-3) / 1e6, "unit": "Mvoxel/s",
            "e2e": {"ms_per_volume": ms_e2e, "value": nvox / (ms_e2e * 1e-3) / 1e6,
                    "h2d_bytes_per_step_per_gpu": int((z1 - z0 + 2) * X * Y * 2), "d2h_bytes_per_step_per_gpu": int((z1 - z0) * X * Y * 2)},
            "one_gpu": {"ms_per_volume": one[0], "e2e_ms_per_volume": one[1], "n_segments": one[2]},
            "speedup_vs_1gpu": one[0] / ms_res, "e2e_speedup_vs_1gpu": one[1] / ms_e2e,
            "bit_exact": bit_exact, "checked_against": "tests/golden/large.json (oracle output, CRC-32 per z-plane)" if gold else None,
            "resident_layout": "slab + one ghost plane per side in HBM",
            "collective": "one all-gather of the per-rank detection tables (label bounding boxes, slice sets); NCCL send/recv of "
                          "whole annotated planes from the owner's HBM only for segments whose bounding plane lies beyond the "
                          "ghost plane (none in C5: the slabs end on annotated planes)",
            "bytes_on_nvlink": {"planes_per_step_max_rank": int(allf[:, 1].max()), "planes_per_step_all_ranks": int(allf[:, 1].sum()),
                                "tables_per_step_per_rank": int(allf[:, 2].max())},
            "segments_per_rank": [int(v) for v in allf[:, 3]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--depth", type=int, default=0, help="volumes in flight per GPU in the e2e throughput measurement "
                    "(0 = 3 on one GPU, fewer when several GPUs share the host's I/O path: 2 on two, 1 beyond)")
    ap.add_argument("--no-check", action="store_true", help="skip the bit-exact check against the oracle before timing")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 lines (extra.workloads) of a single-GPU run")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the slab-sharded C5 volume (sharded)")
    ap.add_argument("--sharded-scale", type=float, default=1.0, help="scale of the sharded C5 volume (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)

    vol, _, cfg = make_config(args.workload, args.scale)
    # N > 1: every rank interpolates its own copy of the same workload volume (independent label volumes, fixed work
    # per GPU); sharding ONE volume over the ranks is tools/run_sharded.py
    nvox = int(vol.size)
    cores = os.cpu_count() or 1
    config = {"workload": "%s: %s" % (args.workload, WORKLOAD_DESC[args.workload]), "dims": list(cfg["dims"]),
              "labels": cfg["n_labels"], "annotated_every": list(cfg["nth"]), "axis": cfg["axis"],
              "use_distance_transform": bool(cfg["use_dt"]), "ball": bool(cfg["ball"]), "seed": cfg["seed"],
              "l2_policy": "inputs larger than L2 (volume in+out = %d MB per step)" % (2 * vol.nbytes // 2 ** 20),
              "parallelism": "one label volume per GPU, no data-path collective" if world > 1 else "single GPU"}

    if args.impl == "reference":
        # the reference's own CPU implementation cannot be built here (needs ITK 5.4); its restatement (oracle/)
        # runs on all host cores, rank 0 only
        if rank != 0:
            return 0
        from oracle import oracle_lib as O
        kw = dict(axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"], threads=cores)
        for _ in range(args.warmup):
            O.interpolate(vol, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.interpolate(vol, **kw)
        dt = (time.perf_counter() - t0) / args.steps
        v = nvox / dt / 1e6
        line = {"impl": "reference", "metric": "interpolated Mvoxel/s per label-volume", "value": v, "unit": "Mvoxel/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Mvoxel/s", "cores": cores, "kind": "port",
                                 "sample": "full %s volume per step, %d steps" % (args.workload, args.steps)},
                "e2e": {"value": v, "unit": "Mvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # libraries (NCCL's version banner) write to stdout; rank 0 must print ONE JSON line there: keep the real stdout
    # aside and point fd 1 at stderr for everything else
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from itkcontourinterpolation_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    try:
        # host threads and the pinned buffers they first touch stay on the NUMA node of this rank's GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = L.lib()
    ctx = ctypes.c_void_p()
    rc = lib.mci_create(ctypes.byref(ctx), local_rank)
    if rc:
        raise SystemExit("mci_create failed: %d" % rc)

    def check(rc):
        if rc:
            raise SystemExit("mci error %d: %s" % (rc, lib.mci_last_error(ctx).decode()))

    opt = L.FilterOptions()
    lib.mci_filter_default_options(ctypes.byref(opt))
    opt.axis = cfg["axis"]
    opt.use_distance_transform = int(cfg["use_dt"])
    opt.use_ball_structuring_element = int(cfg["ball"])
    dims = (ctypes.c_int64 * 3)(*cfg["dims"])
    nbytes = vol.nbytes
    h_in = lib.mci_host_alloc(nbytes)
    h_out = lib.mci_host_alloc(nbytes)
    ctypes.memmove(h_in, vol.ctypes.data, nbytes)
    stats = L.FilterStats()

    # ---- correctness first: bit-exact against the oracle (rank 0 also times the oracle = cpu_baseline)
    check(lib.mci_filter_run(ctx, ctypes.byref(opt), h_in, h_out, dims, L.U16, None, 0, ctypes.byref(stats)))
    gpu_out = np.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_uint16)), shape=vol.shape).copy()
    cpu_baseline = None
    if rank == 0:
        runs = 3 if args.workload == "C3" and args.scale <= 1.0 else 1
        t_cpu, ref = oracle_time(vol, cfg, cores, runs)
        if not args.no_check and not np.array_equal(ref, gpu_out):
            raise SystemExit("GPU output differs from the oracle in %d voxels -- no number reported" % int((ref != gpu_out).sum()))
        cpu_baseline = {"value": nvox / t_cpu / 1e6, "unit": "Mvoxel/s", "cores": cores, "kind": "port",
                        "sample": "full %s volume, 1 warm-up + %d timed runs (median), restated reference, baseline only"
                                  % (args.workload, runs)}
    job_stats = stats.as_dict()

    stream = torch.cuda.ExternalStream(lib.mci_stream(ctx), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as cs:
            e0.record(stream)
            for _ in range(args.steps):
                fn()
            e1.record(stream)
            e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / args.steps, cs.summary()

    # ---- value: volume resident in HBM
    check(lib.mci_upload_volume(ctx, h_in, dims, L.U16))
    check(lib.mci_synchronize(ctx))

    def step_resident():
        check(lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(stats)))

    def read_kernel_times():
        kt = (L.KernelTime * 64)()
        nk = ctypes.c_int32(0)
        check(lib.mci_profile_read(ctx, kt, 64, ctypes.byref(nk)))
        return {kt[k].name.decode(): (int(kt[k].launches), float(kt[k].total_ms)) for k in range(min(nk.value, 64))}

    # timed region: CUDA events around the HBM-streaming kernel only (2 events per step; the roofline entry)
    check(lib.mci_profile_enable(ctx, 2))
    l0 = lib.mci_launch_count(ctx)
    ms_res, clocks = timed(step_resident)
    launches = (lib.mci_launch_count(ctx) - l0) // (args.steps + args.warmup)
    stream_kernel = read_kernel_times()
    # second pass of the same steps with events around EVERY launch: the per-kernel table (adds ~0.2 ms per step)
    check(lib.mci_profile_enable(ctx, 1))
    ms_res_profiled, _ = timed(step_resident)
    kernels = read_kernel_times()
    check(lib.mci_profile_enable(ctx, 0))
    total_kernel_ms = sum(v[1] for v in kernels.values()) or 1.0

    # ---- e2e: host buffers through the reference-facing call
    def step_e2e():
        check(lib.mci_filter_run(ctx, ctypes.byref(opt), h_in, h_out, dims, L.U16, None, 0, ctypes.byref(stats)))

    ms_e2e, clocks_e2e = timed(step_e2e)

    # ---- e2e, throughput mode: a series of volumes through mci_filter_run_batch -- `depth` contexts (own stream, own
    # host thread, own pinned in/out buffers each), so H2D of one volume, kernels of another and D2H of a third overlap.
    # Every volume still crosses PCIe in both directions inside the timed region.
    # one box's host memory / I/O path saturates near 105 GB/s whatever the GPU count: more volumes in
    # flight than that path can feed only make the copies of all GPUs fight
    depth = args.depth if args.depth > 0 else max(1, min(3, 4 // world))
    ctxs = [ctx]
    for _ in range(depth - 1):
        c = ctypes.c_void_p()
        if lib.mci_create(ctypes.byref(c), local_rank):
            raise SystemExit("mci_create failed")
        ctxs.append(c)
    ins, outs = [h_in], [h_out]
    for _ in range(depth - 1):
        a, b = lib.mci_host_alloc(nbytes), lib.mci_host_alloc(nbytes)
        ctypes.memmove(a, vol.ctypes.data, nbytes)
        ins.append(a)
        outs.append(b)
    ctx_arr = (ctypes.c_void_p * depth)(*[c.value for c in ctxs])

    def run_batch(n):
        in_arr = (ctypes.c_void_p * n)(*[ins[k % depth] for k in range(n)])
        out_arr = (ctypes.c_void_p * n)(*[outs[k % depth] for k in range(n)])
        rc = lib.mci_filter_run_batch(ctx_arr, depth, ctypes.byref(opt), in_arr, out_arr, n, dims, L.U16, None)
        if rc:
            raise SystemExit("mci_filter_run_batch failed: %d" % rc)

    for b in outs:
        ctypes.memset(b, 0, nbytes)
    run_batch(max(args.warmup, depth))
    batch_ok = True
    if rank == 0 and not args.no_check:
        for b in outs:  # every in-flight slot holds a bit-exact result
            got = np.ctypeslib.as_array(ctypes.cast(b, ctypes.POINTER(ctypes.c_uint16)), shape=vol.shape)
            if not np.array_equal(got, ref):
                raise SystemExit("batched GPU output differs from the oracle -- no number reported")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as cs_b:
        e0.record(stream)
        run_batch(args.steps)  # returns when the last D2H has completed
        e1.record(stream)
        e1.synchronize()
    barrier()
    ms_batch = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_batch], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_batch = float(t.item())
    ms_batch /= args.steps
    clocks_batch = cs_b.summary()

    extra = None
    if world == 1 and not args.no_extra and args.workload == "C3" and args.scale == 1.0:
        extra = {"workloads": [extra_workload(lib, L, torch, local_rank, name, 5, 3) for name in ("C4", "C5")]}
    sharded = None
    if world > 1 and not args.no_sharded:
        sharded = sharded_c5(torch, dist, L, local_rank, rank, world, args.sharded_scale, min(args.steps, 10), 3, lib)

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel = largest share of device time; the HBM-streaming kernel of this path is k_copy_flag
        # (reads every input voxel once, writes every output voxel once: 2*s*N algorithmic bytes per launch)
        dom = max(kernels.items(), key=lambda kv: kv[1][1])[0] if kernels else None
        table = {k: {"launches_per_step": v[0] / float(args.steps + args.warmup), "ms_per_step": v[1] / (args.steps + args.warmup),
                     "share": v[1] / total_kernel_ms} for k, v in sorted(kernels.items(), key=lambda kv: -kv[1][1])}
        roof = None
        if "k_copy_flag" in stream_kernel:
            n_l, t_ms = stream_kernel["k_copy_flag"]
            avg_ms = t_ms / n_l
            alg = 2.0 * vol.itemsize * nvox
            achieved = alg / (avg_ms * 1e-3) / 1e9
            traffic = None
            prof = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(prof):
                try:
                    traffic = json.load(open(prof)).get("k_copy_flag", {}).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            step_gbs = alg / (ms_res * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "scope": "whole step: the volume's algorithmic bytes (read every input voxel "
                    "once, write every output voxel once) over the step time; `kernel` is the kernel with the largest share of "
                    "device time, which works on L2-resident bit-packed shapes and moves next to no DRAM bytes itself",
                    "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_step": alg, "step_ms": ms_res,
                    "dominant_kernel_share": (kernels[dom][1] / total_kernel_ms) if dom else None,
                    "streaming_kernel": {"kernel": "k_copy_flag", "achieved": achieved, "frac": achieved / peak, "avg_launch_ms": avg_ms,
                                         "algorithmic_bytes_per_launch": alg, "traffic": traffic,
                                         "with_detect_marked_frac": alg / ((t_ms + kernels.get("k_detect_marked", (1, 0.0))[1]) / n_l * 1e-3) / 1e9 / peak}}
        line = {"metric": "interpolated Mvoxel/s per label-volume", "value": world * nvox / (ms_res * 1e-3) / 1e6, "unit": "Mvoxel/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u16", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": world * nvox / (ms_batch * 1e-3) / 1e6, "unit": "Mvoxel/s", "ms_per_step": ms_batch,
                        "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "clocks": clocks_batch,
                        "mode": "mci_filter_run_batch: %d volumes in flight per GPU (one context, stream and host thread each), "
                                "pinned host in -> pinned host out for every volume" % depth,
                        "one_volume_at_a_time": {"value": world * nvox / (ms_e2e * 1e-3) / 1e6, "ms_per_step": ms_e2e,
                                                 "call": "mci_filter_run", "clocks": clocks_e2e}},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "roofline": roof,
                "cpu_baseline": cpu_baseline, "kernels": table,
                "kernels_note": "second pass of the same steps with CUDA events around every launch (%.3f ms/step)" % ms_res_profiled,
                "job": {k: job_stats[k] for k in ("n_labels", "n_annotated_slices", "n_segments", "n_pairs", "n_jobs", "n_components")},
                "bit_exact_vs_oracle": (not args.no_check)}
        if extra is not None:
            line["extra"] = extra
        if sharded is not None:
            line["sharded"] = sharded
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    for a, b in zip(ins, outs):
        lib.mci_host_free(a)
        lib.mci_host_free(b)
    for c in ctxs:
        lib.mci_destroy(c)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
