"""One label volume sharded across the GPUs of one box by (label, axis, slice-pair) segment.

Two ways to combine the partial results (ShardedInterpolator.run(mode=...)):

  "masks" (default): every rank all-gathers the mid slices the others produced AS BIT-PACKED SHAPES (48-byte record
      + A/8 bytes per slice: a few hundred MB for the 8.6 GB C5 volume) and replays their volume writes locally with
      mci_apply_masks.  Every rank ends with the full output; no voxel ever crosses NVLink.
  "slabs": z-slabs of the partial output volumes are exchanged and max-combined (described below); simple, but it
      moves the whole volume.

Every rank holds the input volume, runs slice detection, and interpolates only the segments
k with k % world == rank (mci_filter_options.shard_index / shard_count); its partial output is the input
copy plus its own mid slices.  One exchange step follows (the only inter-GPU traffic of the path):

  1. the volume is cut into `world` z-slabs; rank r sends slab q of its partial output to rank q
     (one all-to-all of bytes -- NCCL has no 16-bit integer reduction, and bytes need none);
  2. rank q combines the `world` partial slabs with the write rule out = max(out, label) (X:754-763);
  3. the finished slabs are gathered on every rank (all-gather) or left sharded.

`combine` and `exchange_and_combine` are backend-agnostic: tests/test_distributed_cpu.py drives them on
gloo with a CPU stand-in producing the partial outputs; on GPUs the NCCL backend moves the same bytes over NVLink.
"""
import numpy as np


def slab_bounds(nz, world):
    """z-range [lo, hi) owned by each rank (contiguous, sizes differ by at most one slice)."""
    base, extra = divmod(nz, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exchange_and_combine(partial, group=None, combine_max=None, gather=True):
    """partial: torch tensor [Z, Y, X] (this rank's partial output, any integer dtype of 1 or 2 bytes).
    Returns the combined volume (gather=True) or this rank's finished slab."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    Z = partial.shape[0]
    bounds = slab_bounds(Z, world)
    plane = partial[0].numel()
    maxz = max(hi - lo for lo, hi in bounds)
    item = partial.element_size()
    # fixed-size slots so that one all_to_all_single moves everything
    send = torch.zeros((world, maxz * plane), dtype=partial.dtype, device=partial.device)
    for q, (lo, hi) in enumerate(bounds):
        send[q, :(hi - lo) * plane] = partial[lo:hi].reshape(-1)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(torch.uint8).view(-1), send.view(torch.uint8).view(-1), group=group)
    lo, hi = bounds[rank]
    n = (hi - lo) * plane
    mine = recv[0, :n].clone() if n else recv[0, :0].clone()
    for q in range(1, world):
        if n == 0:
            break
        if combine_max is not None:
            combine_max(mine, recv[q, :n])
        else:
            torch.maximum(mine, recv[q, :n], out=mine) if mine.dtype not in (torch.uint16,) else \
                mine.copy_(torch.maximum(mine.to(torch.int32), recv[q, :n].to(torch.int32)).to(torch.uint16))
    if not gather:
        return mine.view(hi - lo, *partial.shape[1:])
    slot = torch.zeros(maxz * plane, dtype=partial.dtype, device=partial.device)
    slot[:n] = mine
    allslots = torch.empty((world, maxz * plane), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(allslots.view(torch.uint8).view(-1), slot.view(torch.uint8).view(-1), group=group) \
        if hasattr(dist, "all_gather_into_tensor") and partial.device.type == "cuda" else \
        dist.all_gather([allslots[q].view(torch.uint8) for q in range(world)], slot.view(torch.uint8), group=group)
    out = torch.empty_like(partial)
    for q, (lo, hi) in enumerate(bounds):
        out[lo:hi] = allslots[q, :(hi - lo) * plane].view(hi - lo, *partial.shape[1:])
    return out


class ShardedInterpolator:
    """GPU driver: one process per GPU (torch.distributed, NCCL).  Volumes are torch CUDA tensors [Z, Y, X]."""

    def __init__(self, device, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib as L
        self.L, self.ct, self.torch, self.dist = L, ctypes, torch, dist
        self.group = group
        self.device = torch.device("cuda", device)
        self.lib = L.lib()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.mci_create(ctypes.byref(self.ctx), device)
        if rc:
            raise L.MciError(rc, "mci_create failed (no usable CUDA device; there is no CPU fallback)")
        self.opt = L.FilterOptions()
        self.lib.mci_filter_default_options(ctypes.byref(self.opt))
        self.stream = torch.cuda.ExternalStream(self.lib.mci_stream(self.ctx), device=self.device)
        self._ptype = {torch.uint8: L.U8, torch.uint16: L.U16, torch.int16: L.I16}

    def close(self):
        if self.ctx:
            self.lib.mci_destroy(self.ctx)
            self.ctx = self.ct.c_void_p()

    def _check(self, rc):
        if rc:
            raise self.L.MciError(rc, self.lib.mci_last_error(self.ctx).decode())

    def _combine_by_masks(self, partial, world):
        """all-gather (records, bit rows) of every rank's mid slices, replay the others' into `partial`."""
        torch, dist, ct = self.torch, self.dist, self.ct
        recs, words = ct.c_void_p(), ct.c_void_p()
        nrec, nword = ct.c_int64(0), ct.c_int64(0)
        self._check(self.lib.mci_get_emitted(self.ctx, ct.byref(recs), ct.byref(nrec), ct.byref(words), ct.byref(nword)))
        REC = 48
        sizes = torch.tensor([nrec.value, nword.value], dtype=torch.int64, device=self.device)
        allsizes = torch.empty((world, 2), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allsizes.view(-1), sizes, group=self.group)
        allsizes = allsizes.cpu()
        max_rec, max_word = int(allsizes[:, 0].max()), int(allsizes[:, 1].max())
        if max_rec == 0:
            return partial
        slot = max_rec * REC + max_word * 4
        slot = (slot + 15) // 16 * 16
        send = torch.empty(slot, dtype=torch.uint8, device=self.device)
        if nrec.value:
            self._check(self.lib.mci_copy_emitted(self.ctx, send.data_ptr(), max_rec * REC))
        gathered = torch.empty((world, slot), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(gathered.view(-1), send, group=self.group)
        torch.cuda.current_stream(self.device).synchronize()
        rank = dist.get_rank(self.group)
        for q in range(world):
            n_r, n_w = int(allsizes[q, 0]), int(allsizes[q, 1])
            if q == rank or n_r == 0:
                continue
            base = gathered[q].data_ptr()
            self._check(self.lib.mci_apply_masks(self.ctx, base, n_r, base + max_rec * REC, n_w))
        self._check(self.lib.mci_synchronize(self.ctx))
        return partial

    def _combine(self, dst, src):
        # dst / src were produced by torch on ITS stream (clones, NCCL results); the library launches on the context's
        # own non-blocking stream, which has no implicit ordering with it
        self.torch.cuda.current_stream(self.device).synchronize()
        self._check(self.lib.mci_combine_max(self.ctx, dst.data_ptr(), src.data_ptr(), dst.numel(), self._ptype[dst.dtype]))

    def run(self, vol, gather=True, mode="masks", **options):
        """vol: CUDA tensor [Z, Y, X] present on every rank.  Returns the interpolated volume."""
        torch, dist, ct = self.torch, self.dist, self.ct
        for k, v in options.items():
            setattr(self.opt, k, int(v))
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.opt.shard_index, self.opt.shard_count = rank, world
        vol = vol.contiguous()
        partial = torch.empty_like(vol)
        dims = (ct.c_int64 * 3)(vol.shape[2], vol.shape[1], vol.shape[0])
        self._check(self.lib.mci_attach_device_volume(self.ctx, vol.data_ptr(), partial.data_ptr(), dims, self._ptype[vol.dtype]))
        st = self.L.FilterStats()
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.lib.mci_filter_run_resident(self.ctx, ct.byref(self.opt), None, 0, ct.byref(st)))
        self.last_stats = st.as_dict()
        if world == 1:
            return partial
        if mode == "masks":
            return self._combine_by_masks(partial, world)

        def combine(dst, src):
            # slices of a slot are not 16-byte aligned in general: stage through aligned temporaries when needed
            if dst.data_ptr() % 16 or src.data_ptr() % 16:
                a, b = dst.clone(), src.clone()
                self._combine(a, b)
                self._check(self.lib.mci_synchronize(self.ctx))
                dst.copy_(a)
            else:
                self._combine(dst, src)
                self._check(self.lib.mci_synchronize(self.ctx))

        return exchange_and_combine(partial, self.group, combine, gather)


# =====================================================================================================================
# Slab-sharded runs: nobody streams, uploads or downloads more than its own slab (+ the annotated slices that bound
# the segments reaching into it).  Single-axis runs along z (SetAxis(2)): the unit of parallelism of the reference is
# the segment (X:1523-1537) and a segment writes only the slices strictly between its two annotated slices
# (X:743-764), so the segments -- and with them every voxel write -- partition by z.
#
#   rank r owns planes [z0, z1) of the volume.
#   1. slice detection (X:128-201) on the slab plus one halo plane per side (the 6-neighbour test reads z +- 1);
#      entries of halo planes are dropped, they belong to the neighbour.
#   2. ONE small all-gather merges the per-rank tables: label bounding boxes (X:148-159) and slice sets (X:161-197).
#   3. every rank lists the segments with a mid slice in its slab; those that cross the slab boundary need annotated
#      planes a neighbour owns: fetched from the host volume (host mode) or from the owner's HBM over NVLink (resident
#      mode, NCCL send / recv of whole planes) -- the only voxel traffic between GPUs.
#   4. mci_filter_run_with_detection on the local plane range, restricted to those segments.  No combine step: every
#      output voxel is produced by exactly one rank.
# The plan functions are pure (tests/test_distributed_cpu.py drives them, and the table exchange, on gloo).
# =====================================================================================================================
LABEL_DT = np.dtype([("label", "<i8"), ("idx", "<i4", (3,)), ("size", "<i4", (3,))])   # mci_label_info
SLICE_DT = np.dtype([("axis", "<i4"), ("slice", "<i4"), ("label", "<i8")])                # mci_labeled_slice


def merge_detection(tables):
    """tables: per rank (labels [n,7] = label, bbox index xyz, bbox size xyz; slices [m,2] = label, z), global
    coordinates.  Returns (boxes [L,7] int64: label, lo xyz, hi xyz, sorted by label; slices [S,2] int64: (label, z)
    rows, unique, sorted by label then z).  Vectorised: C5 has 200 labels x 128 annotated planes."""
    labs = [np.asarray(t[0], dtype=np.int64).reshape(-1, 7) for t in tables]
    lab = np.concatenate(labs) if labs else np.zeros((0, 7), np.int64)
    if len(lab):
        # rows sorted by label, then one segmented reduction per bound (ufunc.at is an order of magnitude slower)
        lab = lab[np.argsort(lab[:, 0], kind="stable")]
        first = np.flatnonzero(np.concatenate(([True], lab[1:, 0] != lab[:-1, 0])))
        lo = np.minimum.reduceat(lab[:, 1:4], first, axis=0)
        hi = np.maximum.reduceat(lab[:, 1:4] + lab[:, 4:7] - 1, first, axis=0)
        boxes = np.concatenate([lab[first, 0:1], lo, hi], axis=1)
    else:
        boxes = np.zeros((0, 7), np.int64)
    sls = [np.asarray(t[1]).reshape(-1, 2) for t in tables]
    sl = np.concatenate(sls) if sls else np.zeros((0, 2), np.int64)
    if len(sl):
        # unique rows in lexicographic order (label, then z) through one sorted 64-bit key
        key = (sl[:, 0].astype(np.int64) << 32) | sl[:, 1].astype(np.int64)  # planes are >= 0
        key.sort()
        if len(key) > 1:
            key = key[np.concatenate(([True], key[1:] != key[:-1]))]
        sl = np.empty((len(key), 2), np.int64)
        sl[:, 0] = key >> 32
        sl[:, 1] = key & 0xFFFFFFFF
    else:
        sl = np.zeros((0, 2), np.int64)
    return boxes, sl


def segments_of(slices, label=0):
    """(i, j) of every segment of InterpolateAlong (X:1492-1519): consecutive annotated planes of one label that are
    not adjacent.  slices: merge_detection's [S,2] table."""
    if len(slices) < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    same = (slices[1:, 0] == slices[:-1, 0]) & (slices[:-1, 1] + 1 < slices[1:, 1])
    if label:
        same &= slices[1:, 0] == label
    return slices[:-1, 1][same], slices[1:, 1][same]


def plan_slab(slices, z0, z1, label=0):
    """Which planes rank [z0, z1) must hold: returns (zlo, zhi, needed) -- the contiguous range covering its slab and
    both annotated planes of every segment that has a mid slice in the slab, and the sorted annotated planes outside
    [z0, z1) among them.  slices: merge_detection's table (or {label: sorted z list})."""
    if isinstance(slices, dict):
        rows = [(lab, z) for lab, zs in slices.items() for z in zs]
        slices = np.unique(np.asarray(rows, dtype=np.int64).reshape(-1, 2), axis=0) if rows else np.zeros((0, 2), np.int64)
    if z1 <= z0:
        return z0, z1, []
    si, sj = slices if isinstance(slices, tuple) else segments_of(slices, label)  # (i, j) arrays: computed once by the caller
    # a segment has i + 1 <= j - 1, so its mid slices meet [z0, z1) iff i + 1 <= z1 - 1 and j - 1 >= z0
    idx = np.flatnonzero((si < z1 - 1) & (sj > z0))
    if not len(idx):
        return z0, z1, []
    si, sj = si[idx], sj[idx]
    ends = np.concatenate([si, sj])
    ends = np.unique(ends[(ends < z0) | (ends >= z1)])
    return int(min(z0, si.min())), int(max(z1, sj.max() + 1)), ends.tolist()


def far_planes_needed(segs, bounds):
    """True if some rank needs an annotated plane that is not the plane next to its slab, i.e. some segment reaches
    from at least two planes before a slab boundary to beyond it.  When False, a resident step with ghost planes has
    nothing to plan for the other ranks (C5: the slabs end on annotated planes), which saves planning every rank."""
    si, sj = segs
    cuts = np.asarray(sorted({int(b[0]) for b in bounds if b[1] > b[0]} | {int(b[1]) for b in bounds if b[1] > b[0]}), dtype=np.int64)
    if not len(si) or len(cuts) <= 2:
        return False
    cuts = cuts[1:-1]  # interior boundaries: a boundary c is the first plane of the upper slab
    # rank below c needs plane j > c (beyond its ghost plane c) when i < c - 1 and j > c;
    # rank above c needs plane i < c - 1 (before its ghost plane c - 1) when i < c - 1 and j > c
    return bool((np.searchsorted(cuts, sj - 1, "right") > np.searchsorted(cuts, si + 2, "left")).any())


def exchange_tables(labels, slices, group=None, device="cpu", cap_labels=512, cap_slices=4096, buffers=None):
    """All-gather of the per-rank detection tables (int32 rows).  One collective of fixed-size slots (50 KB per rank
    by default); a second, exact one only if some rank's tables overflow the slot.  buffers: dict the caller keeps
    between calls (pinned staging + device tensors, so that a step allocates nothing)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    labels = np.asarray(labels, dtype=np.int32).reshape(-1, 7)
    slices = np.asarray(slices, dtype=np.int32).reshape(-1, 2)
    on_gpu = str(device) != "cpu"

    def gather(cl, cs):
        n = 4 + 7 * cl + 2 * cs
        key = (cl, cs)
        if buffers is not None and buffers.get("key") == key:
            hs, hr, ds, dr = buffers["hs"], buffers["hr"], buffers["ds"], buffers["dr"]
        else:
            hs = torch.zeros(n, dtype=torch.int32)
            hr = torch.zeros((world, n), dtype=torch.int32)
            if on_gpu:
                hs, hr = hs.pin_memory(), hr.pin_memory()
                ds = torch.empty(n, dtype=torch.int32, device=device)
                dr = torch.empty((world, n), dtype=torch.int32, device=device)
            else:
                ds, dr = hs, hr
            if buffers is not None:
                buffers.update(key=key, hs=hs, hr=hr, ds=ds, dr=dr)
        buf = hs.numpy()
        buf[0], buf[1] = len(labels), len(slices)
        if len(labels) <= cl and len(slices) <= cs:
            buf[4:4 + 7 * len(labels)] = labels.reshape(-1)
            buf[4 + 7 * cl:4 + 7 * cl + 2 * len(slices)] = slices.reshape(-1)
        if on_gpu:
            ds.copy_(hs, non_blocking=True)
            dist.all_gather_into_tensor(dr.view(-1), ds, group=group)
            hr.copy_(dr, non_blocking=True)
            torch.cuda.current_stream(dr.device).synchronize()
        else:
            dist.all_gather([hr[q] for q in range(world)], hs, group=group)
        return hr.numpy()

    got = gather(cap_labels, cap_slices)
    nl, ns = int(got[:, 0].max()), int(got[:, 1].max())
    if nl > cap_labels or ns > cap_slices:
        cap_labels, cap_slices = max(nl, cap_labels), max(ns, cap_slices)
        got = gather(cap_labels, cap_slices)
    out = []
    for q in range(world):
        a, b = int(got[q, 0]), int(got[q, 1])
        out.append((got[q, 4:4 + 7 * a].reshape(-1, 7).copy(), got[q, 4 + 7 * cap_labels:4 + 7 * cap_labels + 2 * b].reshape(-1, 2).copy()))
    return out, int(got.nbytes)


class SlabShardedInterpolator:
    """One volume, one z-slab per GPU (one process per GPU, torch.distributed).  See the block comment above.

    host mode:      run_host(window, window_z0, out_slab)  -- pinned host planes in, pinned host slab out (end to end)
    resident mode:  load_slab(t) once, then run_resident() -- the slab stays in HBM, extra planes come from peers
    """

    def __init__(self, device, dims, dtype, group=None, margin=16):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib as L
        self.L, self.ct, self.torch, self.dist = L, ctypes, torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", device)
        self.X, self.Y, self.Z = (int(d) for d in dims)
        self.dtype = dtype
        self.bounds = slab_bounds(self.Z, self.world)
        self.z0, self.z1 = self.bounds[self.rank]
        self.lib = L.lib()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.mci_create(ctypes.byref(self.ctx), device)
        if rc:
            raise L.MciError(rc, "mci_create failed (no usable CUDA device; there is no CPU fallback)")
        self.opt = L.FilterOptions()
        self.lib.mci_filter_default_options(ctypes.byref(self.opt))
        self.opt.axis = 2
        # torch-side work (copies, NCCL) runs on a torch-owned stream; the library launches on the context's own
        # stream and synchronises it before it returns, so the only ordering needed is "torch work done before a
        # library call" (_lib_call).  (Wrapping the context's stream as a torch ExternalStream would leave pinned
        # tensors that were copied on it recording events on a destroyed stream after close().)
        self.stream = torch.cuda.Stream(device=self.device)
        self._ptype = {torch.uint8: L.U8, torch.uint16: L.U16, torch.int16: L.I16}[dtype]
        self.margin = 0
        self._alloc(margin)
        self.nvlink_bytes = 0
        self.table_bytes = 0
        self.last_stats = {}
        self._tab_buffers = {}
        # NCCL moves device memory (NVLink); any other backend (gloo, in tests) gets the planes staged through the host
        self._nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def close(self):
        if self.ctx:
            self.stream.synchronize()
            self.lib.mci_destroy(self.ctx)
            self.ctx = self.ct.c_void_p()
            self.buf_in = self.buf_out = self._attached = self._compact = None

    def _check(self, rc):
        if rc:
            raise self.L.MciError(rc, self.lib.mci_last_error(self.ctx).decode())

    def _lib_call(self, fn, *args):
        self.stream.synchronize()  # everything torch queued (uploads, received planes) has landed
        self._check(fn(self.ctx, *args))

    # device planes: global z lives at buffer plane z - z0 + margin
    def _alloc(self, margin):
        torch = self.torch
        old = (self.buf_in, self.margin) if self.margin else None
        n = (self.z1 - self.z0) + 2 * margin
        with torch.cuda.stream(self.stream):
            self.buf_in = torch.zeros((n, self.Y, self.X), dtype=self.dtype, device=self.device)
            self.buf_out = torch.zeros((n, self.Y, self.X), dtype=self.dtype, device=self.device)
            if old is not None:
                src, m = old
                self.buf_in[margin - m:margin - m + src.shape[0]] = src
        self.margin = margin

    def _p(self, z):
        return z - self.z0 + self.margin

    def set_options(self, **options):
        for k, v in options.items():
            setattr(self.opt, k, int(v))
        if self.opt.axis != 2:
            raise ValueError("slab-sharded runs interpolate along z (SetAxis(2)); use ShardedInterpolator otherwise")

    def _attach(self, zlo, zhi):
        a, b = self._p(zlo), self._p(zhi)
        tin, tout = self.buf_in[a:b], self.buf_out[a:b]
        self._compact = None
        if tin.data_ptr() % 16 or tout.data_ptr() % 16:  # odd plane sizes: work on aligned copies
            with self.torch.cuda.stream(self.stream):
                tin, tout = tin.clone(), tout.clone()
            self._compact = (tout, a, b)
        dims = (self.ct.c_int64 * 3)(self.X, self.Y, zhi - zlo)
        self._lib_call(self.lib.mci_attach_device_volume, tin.data_ptr(), tout.data_ptr(), dims, self._ptype)
        self._attached = (tin, tout)

    def _detach(self):
        if self._compact is not None:
            tout, a, b = self._compact
            with self.torch.cuda.stream(self.stream):
                self.buf_out[a:b] = tout
            self._compact = None

    def _detect(self):
        """slab + halo planes -> this rank's tables in global coordinates (halo entries dropped)"""
        ct = self.ct
        h0, h1 = (1 if self.z0 > 0 else 0), (1 if self.z1 < self.Z else 0)
        if self.z1 <= self.z0:
            return np.zeros((0, 7), np.int32), np.zeros((0, 2), np.int32)
        self._attach(self.z0 - h0, self.z1 + h1)
        nl, ns = ct.c_int32(0), ct.c_int32(0)
        self._lib_call(self.lib.mci_detect_slices, 2, ct.byref(nl), ct.byref(ns))
        labels = np.zeros(max(nl.value, 1), LABEL_DT)
        slices = np.zeros(max(ns.value, 1), SLICE_DT)
        self._check(self.lib.mci_get_detection(self.ctx, labels.ctypes.data, slices.ctypes.data))
        self._detach()
        labels, slices = labels[:nl.value], slices[:ns.value]
        base = self.z0 - h0
        lab = np.empty((len(labels), 7), np.int32)
        lab[:, 0] = labels["label"]
        lab[:, 1:4] = labels["idx"]
        lab[:, 3] += base
        lab[:, 4:7] = labels["size"]
        z = slices["slice"].astype(np.int64) + base
        keep = (slices["axis"] == 2) & (z >= self.z0) & (z < self.z1)
        sl = np.stack([slices["label"][keep].astype(np.int32), z[keep].astype(np.int32)], axis=1).reshape(-1, 2)
        return lab, sl

    def _gather_tables(self, lab, sl):
        if self.world == 1:
            return [(lab, sl)]
        with self.torch.cuda.stream(self.stream):
            tables, nbytes = exchange_tables(lab, sl, self.group, self.device if self._nccl else "cpu", buffers=self._tab_buffers)
        self.table_bytes += nbytes
        return tables

    def _run_local(self, boxes, slices, zlo, zhi):
        ct, L = self.ct, self.L
        self._attach(zlo, zhi)
        li = np.zeros(max(len(boxes), 1), LABEL_DT)
        n = len(boxes)
        li["label"][:n] = boxes[:, 0]
        li["idx"][:n] = boxes[:, 1:4]
        li["size"][:n] = boxes[:, 4:7] - boxes[:, 1:4] + 1
        li["idx"][:n, 2] = np.maximum(0, boxes[:, 3] - zlo)  # the z extent of a box is not used by runs along z
        li["size"][:n, 2] = 1
        rows = slices[(slices[:, 1] >= zlo) & (slices[:, 1] < zhi)]
        ls = np.zeros(max(len(rows), 1), SLICE_DT)
        ls["axis"][:len(rows)] = 2
        ls["slice"][:len(rows)] = rows[:, 1] - zlo
        ls["label"][:len(rows)] = rows[:, 0]
        st = L.FilterStats()
        self._lib_call(self.lib.mci_filter_run_with_detection, ct.byref(self.opt), li.ctypes.data, n, ls.ctypes.data, len(rows),
                       self.z0 - zlo, self.z1 - zlo, ct.byref(st))
        self._detach()
        self.last_stats = st.as_dict()

    def _ensure_margin(self, zlo, zhi):
        need = max(self.z0 - zlo, zhi - self.z1, 1)
        if need > self.margin:
            self._lib_call(self.lib.mci_synchronize)
            self._alloc(need)
            with self.torch.cuda.stream(self.stream):  # detection had initialised the output of the old buffers
                self.buf_out[self._p(self.z0):self._p(self.z1)] = self.buf_in[self._p(self.z0):self._p(self.z1)]
            return True
        return False

    # ---- host mode ------------------------------------------------------------------------------------------------
    def run_host(self, window, window_z0, out_slab, **options):
        """window: pinned host tensor holding planes [window_z0, window_z0 + len) of the volume -- at least the slab
        and whatever bounds the segments reaching into it (the whole volume always works).  out_slab: pinned host
        tensor [z1 - z0, Y, X] that receives this rank's planes of the interpolated volume."""
        torch = self.torch
        self.set_options(**options)
        z0, z1 = self.z0, self.z1

        def h2d(za, zb):
            if zb > za:
                if za < window_z0 or zb > window_z0 + window.shape[0]:
                    raise ValueError("host window [%d, %d) does not hold planes [%d, %d)" % (window_z0, window_z0 + window.shape[0], za, zb))
                with torch.cuda.stream(self.stream):
                    self.buf_in[self._p(za):self._p(zb)].copy_(window[za - window_z0:zb - window_z0], non_blocking=True)

        h0, h1 = (1 if z0 > 0 else 0), (1 if z1 < self.Z else 0)
        h2d(z0 - h0, z1 + h1)
        lab, sl = self._detect()
        bboxes, slices = merge_detection(self._gather_tables(lab, sl))
        zlo, zhi, needed = plan_slab(slices, z0, z1, int(self.opt.label))
        self._ensure_margin(zlo, zhi)
        for z in needed:
            if not z0 - h0 <= z < z1 + h1:
                h2d(z, z + 1)
        if z1 > z0:
            self._run_local(bboxes, slices, zlo, zhi)
            with torch.cuda.stream(self.stream):
                out_slab.copy_(self.buf_out[self._p(z0):self._p(z1)], non_blocking=True)
        self.stream.synchronize()
        return out_slab

    # ---- resident mode --------------------------------------------------------------------------------------------
    def load_slab(self, slab, ghost_lo=None, ghost_hi=None):
        """slab: tensor [z1 - z0, Y, X] (host or device) = this rank's planes of the input volume.  ghost_lo / ghost_hi:
        the plane just below / above the slab ([Y, X]); when the resident layout carries these ghost planes (both, or
        the one that exists at a volume end) run_resident skips the halo exchange -- slice detection reads z +- 1."""
        with self.torch.cuda.stream(self.stream):
            self.buf_in[self._p(self.z0):self._p(self.z1)].copy_(slab, non_blocking=True)
            if ghost_lo is not None and self.z0 > 0:
                self.buf_in[self._p(self.z0 - 1)].copy_(ghost_lo, non_blocking=True)
            if ghost_hi is not None and self.z1 < self.Z:
                self.buf_in[self._p(self.z1)].copy_(ghost_hi, non_blocking=True)
        self.stream.synchronize()
        self._ghosts_agreed = None
        self.has_ghosts = (ghost_lo is not None or self.z0 == 0) and (ghost_hi is not None or self.z1 == self.Z)

    def _all_have_ghosts(self):
        """every rank must take the same branch (the halo exchange is a collective): agreed once, at the first run"""
        if getattr(self, "_ghosts_agreed", None) is None:
            mine = 1 if getattr(self, "has_ghosts", False) else 0
            if self.world > 1:
                t = self.torch.tensor([mine], dtype=self.torch.int32, device=self.device if self._nccl else "cpu")
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
                mine = int(t.item())
            self._ghosts_agreed = bool(mine)
        return self._ghosts_agreed

    def output_slab(self):
        return self.buf_out[self._p(self.z0):self._p(self.z1)]

    def _owner(self, z):
        for q, (lo, hi) in enumerate(self.bounds):
            if lo <= z < hi:
                return q
        raise ValueError(z)

    def _fetch_planes(self, needs):
        """needs[q] = planes rank q lacks (every rank passes the same table): owners send, over NCCL / NVLink"""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return
        ops, landed = [], []
        plane_bytes = self.X * self.Y * self.buf_in.element_size()
        self.stream.synchronize()
        for q in range(self.world):
            for z in needs[q]:
                o = self._owner(z)
                if o == q or self.rank not in (o, q):
                    continue
                plane = self.buf_in[self._p(z)].view(torch.uint8)
                if o == self.rank:
                    ops.append(dist.P2POp(dist.isend, plane if self._nccl else plane.cpu(), q, self.group))
                elif q == self.rank:
                    dst = plane if self._nccl else torch.empty(plane.shape, dtype=torch.uint8)
                    landed.append((plane, dst))
                    ops.append(dist.P2POp(dist.irecv, dst, o, self.group))
                    self.nvlink_bytes += plane_bytes
        if ops:
            with torch.cuda.stream(self.stream):
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                if not self._nccl:
                    for plane, dst in landed:
                        plane.copy_(dst)

    def run_resident(self, **options):
        """the slab loaded by load_slab is interpolated in place in HBM; returns the device tensor of output planes"""
        import time
        self.set_options(**options)
        bounds = self.bounds
        t = [time.perf_counter()]
        halos = [[z for z in (lo - 1, hi) if 0 <= z < self.Z and hi > lo] for lo, hi in bounds]
        ghosts = self._all_have_ghosts()
        if not ghosts:
            self._fetch_planes(halos)
        self.stream.synchronize()
        t.append(time.perf_counter())
        lab, sl = self._detect()
        t.append(time.perf_counter())
        bboxes, slices = merge_detection(self._gather_tables(lab, sl))
        t.append(time.perf_counter())
        segs = segments_of(slices, int(self.opt.label))
        if ghosts and not far_planes_needed(segs, bounds):
            # every plane any rank needs from a neighbour is a ghost plane it already holds: only the own plan matters
            zlo, zhi, _ = plan_slab(segs, *bounds[self.rank])
            self._ensure_margin(zlo, zhi)
            extra = [[] for _ in range(self.world)]
        else:
            plans = [plan_slab(segs, lo, hi) for lo, hi in bounds]
            zlo, zhi, _ = plans[self.rank]
            if self._ensure_margin(zlo, zhi) and not ghosts:
                self._fetch_planes(halos)
            extra = [[z for z in plans[q][2] if z not in halos[q]] for q in range(self.world)]
        self._fetch_planes(extra)
        self.stream.synchronize()
        t.append(time.perf_counter())
        if self.z1 > self.z0:
            self._run_local(bboxes, slices, zlo, zhi)
        self.stream.synchronize()
        t.append(time.perf_counter())
        self.phase_ms = dict(zip(("halo_exchange", "detect", "gather_tables", "plan_and_fetch", "interpolate"),
                                 [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]))
        return self.output_slab()
