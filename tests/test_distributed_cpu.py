"""World-size-2/3 gloo tests of the multi-GPU path's host logic: segment sharding (k % world == rank),
the z-slab all-to-all of partial outputs, the max combine and the gather.  The oracle stands in for
the CUDA path (no GPU here); itkcontourinterpolation_b200.distributed moves the same bytes over NCCL on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from itkcontourinterpolation_b200.distributed import exchange_and_combine, slab_bounds
from oracle import oracle_lib as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, kw, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = cases.handcrafted()[name] if name in cases.handcrafted() else cases.synthetic_small()[name][0]
        partial = O.interpolate(vol, shard_index=rank, shard_count=world, **kw)
        # torch has no uint16 maximum on CPU: the labels here are < 32768, int16 carries them
        t = torch.from_numpy(partial.astype(np.int16))
        full = exchange_and_combine(t, None, None, gather=True).numpy().astype(vol.dtype)
        slab = exchange_and_combine(t, None, None, gather=False).numpy().astype(vol.dtype)
        q.put((rank, full, slab))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,kw", [
    (2, "ThreeAxis", dict()),
    (2, "C4", dict(use_distance_transform=False, use_ball=True)),
    (3, "C3", dict(axis=2)),
])
def test_sharded_segments_exchange_and_max_combine(world, name, kw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vol = cases.handcrafted()[name] if name in cases.handcrafted() else cases.synthetic_small()[name][0]
    ref = O.interpolate(vol, **kw)
    bounds = slab_bounds(vol.shape[0], world)
    for rank, full, slab in results:
        assert np.array_equal(full, ref), "rank %d gathered volume differs" % rank
        lo, hi = bounds[rank]
        assert np.array_equal(slab, ref[lo:hi])


def test_shards_are_a_partition_of_the_segments():
    vol, cfg = cases.synthetic_small()["C4"]
    kw = dict(use_distance_transform=False, use_ball=True)
    full, st = O.interpolate(vol, return_stats=True, **kw)
    segs = 0
    acc = np.zeros_like(vol)
    for r in range(4):
        part, s = O.interpolate(vol, return_stats=True, shard_index=r, shard_count=4, **kw)
        segs += s["segments"]
        acc = np.maximum(acc, part)
    assert segs == st["segments"]
    assert np.array_equal(acc, full)


def test_slab_bounds_cover_everything():
    for nz in (1, 7, 8, 256):
        for w in (1, 2, 3, 8):
            b = slab_bounds(nz, w)
            assert b[0][0] == 0 and b[-1][1] == nz and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
