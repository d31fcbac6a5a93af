"""Sharded runs on one GPU: every shard of a 3-way split is computed in turn by the CUDA path, the partial
outputs are merged with mci_combine_max, and the result must equal the unsharded oracle output bit for bit.
(The N-process NCCL version of the same flow is tools/run_sharded.py; its exchange logic is covered on
gloo by tests/test_distributed_cpu.py.)"""
import ctypes

import numpy as np
import pytest
import torch

import cases
from itkcontourinterpolation_b200 import _lib as L
from oracle import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw_gpu,kw_cpu", [
    ("C4", dict(use_distance_transform=0, use_ball_structuring_element=1), dict(use_distance_transform=False, use_ball=True)),
    ("C3", dict(axis=2), dict(axis=2)),
    ("Branching", dict(axis=2), dict(axis=2)),
])
def test_three_shards_combine_to_the_full_result(name, kw_gpu, kw_cpu):
    vol, _ = cases.synthetic_small()[name]
    lib = L.lib()
    ctx = ctypes.c_void_p()
    assert lib.mci_create(ctypes.byref(ctx), 0) == 0
    try:
        opt = L.FilterOptions()
        lib.mci_filter_default_options(ctypes.byref(opt))
        for k, v in kw_gpu.items():
            setattr(opt, k, v)
        dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
        d_in = torch.from_numpy(vol.view(np.int16)).cuda()
        acc = None
        nseg = 0
        for r in range(3):
            opt.shard_index, opt.shard_count = r, 3
            part = torch.empty_like(d_in)
            assert lib.mci_attach_device_volume(ctx, d_in.data_ptr(), part.data_ptr(), dims, L.U16) == 0
            st = L.FilterStats()
            rc = lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st))
            assert rc == 0, lib.mci_last_error(ctx)
            nseg += st.n_segments
            ref_part = O.interpolate(vol, shard_index=r, shard_count=3, **kw_cpu)
            assert np.array_equal(part.cpu().numpy().view(np.uint16), ref_part), "shard %d" % r
            if acc is None:
                acc = part
            else:
                assert lib.mci_combine_max(ctx, acc.data_ptr(), part.data_ptr(), acc.numel(), L.U16) == 0
                assert lib.mci_synchronize(ctx) == 0
        full, st_full = O.interpolate(vol, return_stats=True, **kw_cpu)
        assert nseg == st_full["segments"]
        assert np.array_equal(acc.cpu().numpy().view(np.uint16), full)
    finally:
        lib.mci_destroy(ctx)


@pytest.mark.parametrize("name,kw_gpu,kw_cpu", [
    ("C4", dict(use_distance_transform=0, use_ball_structuring_element=1), dict(use_distance_transform=False, use_ball=True)),
    ("C3", dict(axis=2), dict(axis=2)),
])
def test_combine_by_mid_slice_shapes(name, kw_gpu, kw_cpu):
    """The NCCL-free core of the "masks" combine: shard 0's volume receives the mid slices of shards 1 and 2 as
    (record, bit rows) lists through mci_get_emitted / mci_apply_masks and must end up equal to the full result."""
    vol, _ = cases.synthetic_small()[name]
    lib = L.lib()
    dims = (ctypes.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
    d_in = torch.from_numpy(vol.view(np.int16)).cuda()
    ctxs, outs, lists = [], [], []
    try:
        for r in range(3):
            ctx = ctypes.c_void_p()
            assert lib.mci_create(ctypes.byref(ctx), 0) == 0
            ctxs.append(ctx)
            opt = L.FilterOptions()
            lib.mci_filter_default_options(ctypes.byref(opt))
            for k, v in kw_gpu.items():
                setattr(opt, k, v)
            opt.shard_index, opt.shard_count = r, 3
            out = torch.empty_like(d_in)
            outs.append(out)
            assert lib.mci_attach_device_volume(ctx, d_in.data_ptr(), out.data_ptr(), dims, L.U16) == 0
            st = L.FilterStats()
            assert lib.mci_filter_run_resident(ctx, ctypes.byref(opt), None, 0, ctypes.byref(st)) == 0, lib.mci_last_error(ctx)
            recs, words = ctypes.c_void_p(), ctypes.c_void_p()
            nr, nw = ctypes.c_int64(0), ctypes.c_int64(0)
            assert lib.mci_get_emitted(ctx, ctypes.byref(recs), ctypes.byref(nr), ctypes.byref(words), ctypes.byref(nw)) == 0
            lists.append((recs.value, nr.value, words.value, nw.value))
        assert sum(l[1] for l in lists) > 0
        for r in (1, 2):
            recs, nr, words, nw = lists[r]
            if nr:
                assert lib.mci_apply_masks(ctxs[0], recs, nr, words, nw) == 0
        assert lib.mci_synchronize(ctxs[0]) == 0
        full = O.interpolate(vol, **kw_cpu)
        assert np.array_equal(outs[0].cpu().numpy().view(np.uint16), full)
    finally:
        for ctx in ctxs:
            lib.mci_destroy(ctx)


# ---- slab-sharded path: one z-slab per process, several processes sharing this box's GPU, gloo for the plumbing ----
def _slab_volume(name):
    import test_slab_sharded_cpu as T
    if name == "Staggered":
        return T.volume(name)
    return cases.synthetic_small()[name][0]


def _slab_worker(rank, world, port, name, q):
    import os
    import torch.distributed as dist
    from itkcontourinterpolation_b200.distributed import SlabShardedInterpolator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = _slab_volume(name)
        Z, Y, X = vol.shape
        f = SlabShardedInterpolator(0, (X, Y, Z), torch.int16, margin=2)  # small margin: the regrow path runs too
        try:
            host = torch.from_numpy(vol.view(np.int16)).pin_memory()
            out = torch.empty((f.z1 - f.z0, Y, X), dtype=torch.int16).pin_memory()
            f.run_host(host, 0, out)
            e2e = out.numpy().view(np.uint16).copy()
            st = dict(f.last_stats)
            f.load_slab(host[f.z0:f.z1])
            res = f.run_resident().cpu().numpy().view(np.uint16).copy()
            # the same with ghost planes resident: no halo exchange
            f.load_slab(host[f.z0:f.z1], host[f.z0 - 1] if f.z0 > 0 else None, host[f.z1] if f.z1 < Z else None)
            res2 = f.run_resident().cpu().numpy().view(np.uint16).copy()
            assert np.array_equal(res, res2)
            q.put((rank, f.z0, f.z1, e2e, res, st, f.nvlink_bytes))
        finally:
            f.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "Staggered"), (3, "Staggered"), (3, "C5"), (2, "Branching")])
def test_slab_sharded_ranks_reproduce_their_planes(world, name):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vol = _slab_volume(name)
    ref, st_full = O.interpolate(vol, axis=2, return_stats=True)
    nseg = 0
    for rank, z0, z1, e2e, res, st, nv in results:
        assert np.array_equal(e2e, ref[z0:z1]), "host mode, rank %d" % rank
        assert np.array_equal(res, ref[z0:z1]), "resident mode, rank %d" % rank
        nseg += st["n_segments"]
    assert nseg >= st_full["segments"]  # a segment crossing a slab boundary is interpolated by both owners
