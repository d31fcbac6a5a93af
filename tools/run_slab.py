"""torchrun tool: ONE C5 (or scaled) volume slab-sharded over the ranks, with the per-phase host times of a step.
    python -m torch.distributed.run --nproc-per-node N tools/run_slab.py [scale] [steps]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itkcontourinterpolation_b200.distributed import SlabShardedInterpolator, slab_bounds  # noqa: E402
from itkcontourinterpolation_b200.synthetic import CONFIGS, make_config  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cfg = CONFIGS["C5"]
    dims = tuple(max(8, int(round(d * scale))) for d in cfg["dims"])
    z0, z1 = slab_bounds(dims[2], world)[rank]
    w0, w1 = max(0, z0 - 8), min(dims[2], z1 + 9)
    win, _, _ = make_config("C5", scale, z_range=(w0, w1), want_dense=False)
    f = SlabShardedInterpolator(lr, dims, torch.uint16, margin=9)
    host = torch.from_numpy(win).pin_memory()
    out = torch.empty((z1 - z0, dims[1], dims[0]), dtype=torch.uint16).pin_memory()
    ghosts = len(sys.argv) > 3 and sys.argv[3] == "ghosts"
    f.load_slab(host[z0 - w0:z1 - w0], host[z0 - 1 - w0] if ghosts and z0 > 0 else None, host[z1 - w0] if ghosts and z1 < dims[2] else None)
    for k in range(steps):
        dist.barrier()
        t0 = time.perf_counter()
        f.run_resident()
        t1 = time.perf_counter()
        if k == steps - 1:
            print("rank %d resident %.2f ms  %s  stats %s" % (rank, (t1 - t0) * 1e3, {a: round(b, 2) for a, b in f.phase_ms.items()},
                                                           {a: (round(b, 2) if isinstance(b, float) else b) for a, b in f.last_stats.items()}), flush=True)
    for k in range(steps):
        dist.barrier()
        t0 = time.perf_counter()
        f.run_host(host, w0, out)
        t1 = time.perf_counter()
        if k == steps - 1:
            print("rank %d host-to-host %.2f ms" % (rank, (t1 - t0) * 1e3), flush=True)
    f.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
