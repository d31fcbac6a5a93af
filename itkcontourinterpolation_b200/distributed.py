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
