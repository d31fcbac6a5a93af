"""torchrun entry: one label volume sharded by segment over the ranks (NCCL), checked against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_sharded.py [C3|C4|C5] [scale] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from itkcontourinterpolation_b200.distributed import ShardedInterpolator
from itkcontourinterpolation_b200.synthetic import make_config


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    mode = sys.argv[4] if len(sys.argv) > 4 else "masks"
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vol, _, cfg = make_config(name, scale)
    d = torch.from_numpy(vol).cuda()  # torch.uint16
    sh = ShardedInterpolator(local)
    opts = dict(axis=cfg["axis"], use_distance_transform=int(cfg["use_dt"]), use_ball_structuring_element=int(cfg["ball"]))
    out = sh.run(d, mode=mode, **opts)
    got = out.cpu().numpy()
    if rank == 0:
        from oracle import oracle_lib as O
        ref = O.interpolate(vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"], threads=os.cpu_count())
        print("bit-exact vs oracle:", bool(np.array_equal(got, ref)), "segments on rank 0:", sh.last_stats["n_segments"])
    for _ in range(2):
        sh.run(d, mode=mode, **opts)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        sh.run(d, mode=mode, **opts)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("sharded over %d GPUs (combine by %s): %.3f ms per volume (%.0f Mvoxel/s), strong scaling incl. the exchange"
              % (world, mode, t.item(), vol.size / t.item() / 1e3))
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
