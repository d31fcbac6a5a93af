"""World-size-2/3 gloo tests of the SLAB-sharded multi-GPU path's host logic (itkcontourinterpolation_b200.distributed:
merge_detection, plan_slab, exchange_tables): every rank detects slices on its z-slab plus one halo plane, the tables
are all-gathered and merged, each rank plans which foreign annotated planes it needs, runs the filter on its local
plane range with the merged tables, and must reproduce its planes of the single-process result bit for bit.  No GPU
here: the literal restatement (tests/literal_mci.py) stands in for mci_filter_run_with_detection and the oracle for
mci_detect_slices; on GPUs SlabShardedInterpolator drives the same plan over NCCL (tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import literal_mci
from itkcontourinterpolation_b200.distributed import exchange_tables, merge_detection, plan_slab, slab_bounds
from oracle import oracle_lib as O


def volume(name):
    if name == "Staggered":
        # labels annotated at different z positions, so that segments cross every slab boundary differently
        w = np.zeros((14, 36, 40), np.uint16)
        for z, (cy, cx, r) in {0: (14, 14, 7), 5: (18, 20, 8), 12: (22, 26, 5)}.items():
            w[z][cases.disk(w[z].shape, cy, cx, r)] = 7
        for z, (cy, cx, r) in {2: (8, 30, 4), 9: (10, 28, 6)}.items():
            w[z][cases.disk(w[z].shape, cy, cx, r)] = 3
        w[1][cases.disk(w[1].shape, 30, 8, 3)] = 5
        w[13][cases.disk(w[13].shape, 28, 10, 4)] = 5
        return w
    return cases.handcrafted()[name]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def local_tables(vol, z0, z1):
    """what SlabShardedInterpolator._detect does: detection on slab + halo, halo entries dropped, global z"""
    Z = vol.shape[0]
    h0, h1 = (1 if z0 > 0 else 0), (1 if z1 < Z else 0)
    if z1 <= z0:
        return np.zeros((0, 7), np.int32), np.zeros((0, 2), np.int32)
    trip, boxes = O.detect(vol[z0 - h0:z1 + h1], axis=2)
    base = z0 - h0
    lab = np.array([[b[0], b[1], b[2], b[3] + base, b[4], b[5], b[6]] for b in boxes], np.int32).reshape(-1, 7)
    sl = np.array([[t[1], t[2] + base] for t in trip if t[0] == 2 and z0 <= t[2] + base < z1], np.int32).reshape(-1, 2)
    return lab, sl


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = volume(name)
        z0, z1 = slab_bounds(vol.shape[0], world)[rank]
        lab, sl = local_tables(vol, z0, z1)
        tables, _ = exchange_tables(lab, sl, None, "cpu", cap_labels=2, cap_slices=1)  # tiny slots: the exact re-gather runs
        boxes, slices = merge_detection(tables)
        zlo, zhi, needed = plan_slab(slices, z0, z1)
        # the local plane range holds ONLY the slab, its halo planes and the planned foreign planes
        local = np.zeros((zhi - zlo,) + vol.shape[1:], vol.dtype)
        have = set(range(max(z0 - 1, 0), min(z1 + 1, vol.shape[0]))) | set(needed)
        for z in have:
            if zlo <= z < zhi:
                local[z - zlo] = vol[z]
        lb = {int(r[0]): ([r[1], r[2], 0], [r[4], r[5], 0]) for r in boxes}
        ls = {}
        for l, z in slices:
            if zlo <= z < zhi:
                ls.setdefault(int(l), []).append(int(z - zlo))
        out = literal_mci.interpolate_with_detection(local, lb, ls, 2, (z0 - zlo, z1 - zlo), dt=True, ball=False)
        q.put((rank, z0, z1, out[z0 - zlo:z1 - zlo].copy(), needed))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "Staggered"), (3, "Staggered"), (3, "DriftOddGaps"), (2, "OneToThree")])
def test_slab_plan_reproduces_the_single_process_result(world, name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vol = volume(name)
    ref = O.interpolate(vol, axis=2)
    assert (ref != vol).any()
    crossing = 0
    for rank, z0, z1, slab, needed in results:
        assert np.array_equal(slab, ref[z0:z1]), "rank %d planes [%d, %d)" % (rank, z0, z1)
        crossing += len(needed)
    assert crossing > 0  # the case does exercise segments that cross a slab boundary


def test_plan_slab_lists_only_the_bounding_planes():
    slices = {1: [0, 8, 16, 24], 2: [3, 20], 4: [9, 10, 11]}
    assert plan_slab(slices, 8, 16) == (3, 21, [3, 16, 20])
    assert plan_slab(slices, 0, 8) == (0, 21, [8, 20])
    assert plan_slab(slices, 16, 24) == (3, 25, [3, 24])
    assert plan_slab(slices, 8, 16, label=1) == (8, 17, [16])
    assert plan_slab({4: [9, 10, 11]}, 0, 32) == (0, 32, [])  # adjacent annotated slices make no segment (X:1502)
    assert plan_slab(slices, 5, 5) == (5, 5, [])


def test_merge_detection_unions_boxes_and_slice_sets():
    a = (np.array([[7, 2, 3, 0, 4, 5, 1]]), np.array([[7, 0], [7, 5]]))
    b = (np.array([[7, 1, 6, 5, 2, 9, 2], [3, 0, 0, 6, 1, 1, 1]]), np.array([[7, 5], [3, 6]]))
    boxes, slices = merge_detection([a, b])
    assert boxes.tolist() == [[3, 0, 0, 6, 0, 0, 6], [7, 1, 3, 0, 5, 14, 6]]  # label, lo xyz, hi xyz
    assert slices.tolist() == [[3, 6], [7, 0], [7, 5]]


def test_far_plane_test_agrees_with_the_per_rank_plans():
    """the resident path with ghost planes plans only its own rank when no rank needs a plane beyond its ghost planes;
    that shortcut's test must say exactly what planning every rank says, uneven and empty slabs included"""
    from itkcontourinterpolation_b200.distributed import far_planes_needed, segments_of
    rng = np.random.default_rng(11)
    seen = set()
    for trial in range(400):
        nz = int(rng.integers(4, 90))
        rows = set()
        for lab in range(1, int(rng.integers(2, 7))):
            step = int(rng.integers(1, 9))
            for z in (rng.choice(nz, size=int(rng.integers(1, min(nz, 12) + 1)), replace=False) if trial % 2 else range(0, nz, step)):
                rows.add((lab, int(z)))
        slices = np.array(sorted(rows), dtype=np.int64).reshape(-1, 2)
        world = int(rng.integers(1, 7))
        cuts = sorted(int(c) for c in rng.integers(0, nz + 1, size=world - 1))
        if trial % 4 == 0:
            cuts = [c // 8 * 8 for c in cuts]
        bounds = list(zip([0] + cuts, cuts + [nz]))
        label = int(rng.integers(0, 3))
        segs = segments_of(slices, label)
        halos = [[z for z in (lo - 1, hi) if 0 <= z < nz and hi > lo] for lo, hi in bounds]
        plans = [plan_slab(slices, lo, hi, label) for lo, hi in bounds]
        far = any(z not in halos[q] for q in range(len(bounds)) for z in plans[q][2])
        assert far_planes_needed(segs, bounds) == far, (trial, bounds, slices.tolist(), label)
        seen.add(far)
    assert seen == {True, False}
