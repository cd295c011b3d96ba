"""One volume sharded over the GPUs of a box (BASELINE.json configs 3 and 5).

A patch is indivisible (InstanceNorm and the top-k selections are global over it), so the unit of sharding is the
patch: the z-major patch list is cut into `world` contiguous chunks ("slabs").

predict_volume_sharded - the default, OWNER-COMPUTES form.  Rank r forwards its patches into probability slots in its
own HBM; rank g owns the rows x in owned_x_range(X, g, world) of the output and produces them with ONE kernel
(`dcl_gather_finalize_range`) that reads, for every owned voxel, the covering patches' probabilities from the slot they
were written to - local memory or a PEER's, mapped through CUDA IPC and pulled over NVLink inside the kernel.  Nothing is
accumulated across ranks and no accumulator is exchanged: each probability crosses NVLink at most once (about
604 MB / world per rank for the 18-patch plan instead of 179 MB x 5 reduce-scatters), the sums run in patch order exactly
as on one GPU, so the label map is BIT-IDENTICAL to the single-GPU result (tests assert equality).  Collectives:
one tiny all-reduce as the stream-ordered "slots written" barrier, one all-gather of the uint8 label rows (8.9 MB in
total) and one all-reduce of the 13 counters (which is also the "slots free again" barrier of the next volume).

predict_volume_sharded_accumulate - the round-1 form, kept for comparison: private full-size fp32 accumulators
(`dcl_accumulate_patches`) summed by NCCL reduce-scatter, owner finalises a flat voxel range.

The reference has no counterpart (its inference is single-GPU, test_overlap.py:78); the stitched result is defined by
predict_overlap.py:31-58 (+ the sum-then-divide blend of predict_cls.py:184-203) and must not depend on the number of
ranks.  The arithmetic callables default to the CUDA engine; tests inject CPU stand-ins to exercise this host logic
under the `gloo` backend.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._native import DclError, StitchMode


def partition_patches(n_patches: int, world: int):
    """[(first, count)] per rank: contiguous chunks of the (z-major) patch list, sizes differing by at most 1."""
    if world < 1 or n_patches < 0:
        raise DclError("partition_patches: bad arguments")
    base, extra = divmod(n_patches, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def owned_chunk(total_voxels: int, world: int) -> int:
    """Voxels per rank of the output partition (the last rank's range may be short)."""
    return -(-total_voxels // world)


def owned_range(total_voxels: int, rank: int, world: int):
    chunk = owned_chunk(total_voxels, world)
    v0 = min(rank * chunk, total_voxels)
    return v0, min(chunk, total_voxels - v0)


def _exchange(acc, wsum, rank, world, group):
    """Sums the per-rank accumulators; returns (acc_local (4, n), wsum_local (n,) | None, v0, n) for the owned range.
    reduce-scatter per channel plane when the backend has it and the volume divides evenly, else all-reduce."""
    total = acc.shape[1]
    v0, n = owned_range(total, rank, world)
    if world == 1:
        return acc, wsum, 0, total
    backend = dist.get_backend(group)
    if backend == "nccl" and total % world == 0:
        loc = torch.empty((5 if wsum is not None else 4, n), dtype=acc.dtype, device=acc.device)
        for c in range(4):
            dist.reduce_scatter_tensor(loc[c], acc[c], op=dist.ReduceOp.SUM, group=group)
        if wsum is not None:
            dist.reduce_scatter_tensor(loc[4], wsum, op=dist.ReduceOp.SUM, group=group)
        return loc[:4], (loc[4] if wsum is not None else None), v0, n
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    if wsum is not None:
        dist.all_reduce(wsum, op=dist.ReduceOp.SUM, group=group)
    return (acc[:, v0:v0 + n].contiguous(), wsum[v0:v0 + n].contiguous() if wsum is not None else None, v0, n)


def predict_volume_sharded_accumulate(engine, vol, mode=StitchMode.UNIFORM, starts=None, keep_scales=None, target=None,
                                      group=None, accumulate=None, finalize=None):
    """All ranks call this with the SAME volume / plan; returns dict(labels uint8 (X,Y,Zout), counts int64[13])
    identical on every rank.

    accumulate(vol, mode, starts, keep_scales, first, count, acc, wsum) and
    finalize(acc_local, wsum_local, labels_local, target_local, counts) default to the engine's CUDA kernels."""
    mode = StitchMode(mode)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if vol.dim() == 5:
        vol = vol[0]
    X, Y, Z = (int(v) for v in vol.shape[1:])
    weighted = mode in (StitchMode.UNIFORM, StitchMode.GAUSSIAN)
    zout = Z if weighted else 155
    n_patches = len(starts) if weighted else 8
    if weighted and not starts:
        raise DclError("weighted stitch modes need a patch list")
    total = X * Y * zout
    dev = vol.device
    acc = torch.zeros((4, total), dtype=torch.float32, device=dev)
    wsum = torch.zeros(total, dtype=torch.float32, device=dev) if weighted else None
    first, count = partition_patches(n_patches, world)[rank]
    accumulate = accumulate or engine.accumulate_patches
    if count > 0:
        accumulate(vol, mode, starts, keep_scales, first, count, acc, wsum)
    acc_l, wsum_l, v0, n = _exchange(acc, wsum, rank, world, group)

    chunk = owned_chunk(total, world)
    labels_l = torch.zeros(chunk, dtype=torch.uint8, device=dev)      # padded to the common chunk for all_gather
    counts = torch.zeros(13, dtype=torch.int64, device=dev)
    tgt_l = None
    if target is not None:
        tgt_l = target.reshape(-1)[v0:v0 + n].to(device=dev, dtype=torch.uint8).contiguous()
    if n > 0:
        if finalize is not None:
            finalize(acc_l, wsum_l, labels_l[:n], tgt_l, counts)
        else:
            engine.finalize_labels(acc_l, wsum_l, 0, n, labels_l[:n], target=tgt_l, counts=counts)
    if world > 1:
        gathered = torch.empty(world * chunk, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, labels_l, group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        labels = gathered[:total]
    else:
        labels = labels_l[:total]
    return {"labels": labels.reshape(X, Y, zout), "counts": counts, "patches": (first, count)}


# ------------------------------------------------------------------------------------------------------------------
# owner-computes form
# ------------------------------------------------------------------------------------------------------------------
PATCH = 128
SLOT_FLOATS = 4 * PATCH ** 3


def owned_x_range(X: int, rank: int, world: int):
    """Rows x0 <= x < x1 of the output volume owned by `rank`: contiguous blocks of ceil(X / world) x-planes (X is the
    slowest axis, so an owned block is one contiguous voxel range and the label all-gather needs no packing)."""
    step = -(-X // world)
    x0 = min(rank * step, X)
    return x0, min(x0 + step, X)


def assignment_order(starts):
    """The order in which the patches of a plan are DEALT to the ranks (contiguous chunks of this order): x-major, so
    that the patches of a rank share their x-range and the rank uploads one 128-plane slab of the input (76 MB of the
    143 MB volume for the 18-patch plan on 8 ranks; the z-major plan order itself would hand rank 3 the patches at
    x = 112 and x = 0, i.e. the whole volume).  Which rank computes a patch changes nothing in the result: the blend
    reads the slots by patch index, in plan order."""
    return sorted(range(len(starts)), key=lambda i: (int(starts[i][0]), int(starts[i][2]), int(starts[i][1])))


def patch_owner(n_patches: int, world: int):
    """[(rank, local slot index)] per patch of the plan, consistent with partition_patches."""
    out = []
    for r, (first, count) in enumerate(partition_patches(n_patches, world)):
        out += [(r, i) for i in range(count)]
    return out


class _PeerSlots:
    """This rank's slot buffer and the IPC-mapped slot buffers of its peers (set up once per engine / group / plan size
    and reused for every volume)."""

    def __init__(self, engine, group, n_patches, order):
        self.engine, self.group, self.n_patches = engine, group, n_patches
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        parts = partition_patches(n_patches, self.world)
        self.first, self.count = parts[self.rank]
        self.local = engine.slots_ensure(max(1, max(c for _, c in parts)))
        self.bases = [self.local] * self.world
        self.imported = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, engine.ipc_export(self.local), group=group)
            for r in range(self.world):
                if r != self.rank:
                    self.bases[r] = engine.ipc_import(handles[r])
                    self.imported.append(self.bases[r])
        owner = patch_owner(n_patches, self.world)             # by position in the assignment order
        pos = {p: k for k, p in enumerate(order)}
        self.ptrs = [self.bases[owner[pos[i]][0]] + owner[pos[i]][1] * SLOT_FLOATS * 4 for i in range(n_patches)]

    def close(self):
        for p in self.imported:
            try:
                self.engine.ipc_release(p)
            except Exception:
                pass
        self.imported = []


def _peer_slots(engine, group, n_patches, order):
    cache = engine.__dict__.setdefault("_shard_ctx", {})
    key = (id(group), n_patches, tuple(order))
    ctx = cache.get(key)
    if ctx is None or ctx.local != engine.slots_ensure(max(1, ctx.count)):     # (the buffer only ever grows)
        if ctx is not None:
            ctx.close()
        ctx = cache[key] = _PeerSlots(engine, group, n_patches, order)
    return ctx


def predict_volume_sharded(engine, vol, mode=StitchMode.UNIFORM, starts=None, keep_scales=None, target=None, group=None,
                           forward_to_slots=None, finalize_range=None):
    """All ranks call this with the SAME plan (and the same volume, of which a rank reads only its patches' boxes);
    returns dict(labels uint8 (X,Y,Z), counts int64[13]) identical on every rank and bit-identical to one GPU.

    Test hooks (CPU / gloo): forward_to_slots(vol, mode, starts, keep_scales, first, count) -> (count, 4, 128, 128, 128)
    tensor (starts / keep_scales arrive in assignment_order); finalize_range(slots: list of per-patch tensors, x0, x1, labels (X,Y,Z), target | None, counts).  With the
    hooks the slots travel by all_gather (the CPU stand-in for peer memory); without them the CUDA engine is used."""
    mode = StitchMode(mode)
    if mode not in (StitchMode.UNIFORM, StitchMode.GAUSSIAN) or not starts:
        raise DclError("predict_volume_sharded: weighted stitch modes with a patch list (the reference's 8-corner "
                       "crop-overwrite plan has no overlap to exchange: use predict_volume_sharded_accumulate)")
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if vol.dim() == 5:
        if vol.shape[0] != 1:
            raise DclError("one volume per call")
        vol = vol[0]
    X, Y, Z = (int(v) for v in vol.shape[1:])
    n_patches = len(starts)
    order = assignment_order(starts)                        # patches are dealt to the ranks x-major ...
    starts_a = [tuple(int(v) for v in starts[i]) for i in order]
    keep_a = None if keep_scales is None else np.asarray(keep_scales, dtype=np.float32).reshape(n_patches, 16)[order]
    pos = {p: k for k, p in enumerate(order)}               # ... and read back by plan index
    first, count = partition_patches(n_patches, world)[rank]
    x0, x1 = owned_x_range(X, rank, world)
    dev = vol.device
    step = -(-X // world)
    labels = torch.empty((world * step, Y, Z), dtype=torch.uint8, device=dev)      # padded to whole blocks for the gather
    counts = torch.zeros(13, dtype=torch.int64, device=dev)
    if target is not None:
        if tuple(target.shape) != (X, Y, Z):
            raise DclError(f"target must have shape {(X, Y, Z)}")
        if target.device != dev:      # a host target: only the rows this rank owns travel (the counters are per range)
            full = torch.empty((X, Y, Z), dtype=torch.uint8, device=dev)
            if x1 > x0:
                full[x0:x1].copy_(target[x0:x1].to(torch.uint8), non_blocking=True)
            target = full
        else:
            target = target.to(dtype=torch.uint8).contiguous()
    hooks = forward_to_slots is not None or finalize_range is not None
    if hooks:
        mine = forward_to_slots(vol, mode, starts_a, keep_a, first, count)
        slots = [None] * n_patches
        if world > 1:
            cmax = max(c for _, c in partition_patches(n_patches, world))
            pad = torch.zeros((cmax, 4, PATCH, PATCH, PATCH), dtype=torch.float32, device=dev)
            pad[:count] = mine
            allp = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(allp, pad, group=group)
            owner = patch_owner(n_patches, world)
            for i in range(n_patches):
                r, j = owner[pos[i]]
                slots[i] = allp[r][j]
        else:
            slots = [mine[pos[i]] for i in range(n_patches)]
        if x1 > x0:
            finalize_range(slots, x0, x1, labels[:X], target, counts)
    else:
        ctx = _peer_slots(engine, group, n_patches, order)
        if count > 0:
            engine.forward_patches_to_slots(vol, mode, starts_a, keep_a, first, count)
        if world > 1:      # stream-ordered barrier: every rank's slots are written before anybody reads them
            flag = torch.ones(1, dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.SUM, group=group)
        if x1 > x0:
            engine.gather_finalize_range((X, Y, Z), mode, starts, ctx.ptrs, x0, x1, labels[:X], target=target, counts=counts)
    if world > 1:
        dist.all_gather_into_tensor(labels.view(-1), labels[rank * step:(rank + 1) * step].reshape(-1).clone(), group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return {"labels": labels[:X], "counts": counts, "patches": (first, count), "patch_ids": order[first:first + count],
            "rows": (x0, x1)}


def own_x_ranges(starts, first, count, X):
    """Disjoint x ranges [(xa, xb)] of the input volume that the patches [first, first+count) of the plan read."""
    iv = sorted((max(0, int(starts[i][0])), min(X, int(starts[i][0]) + PATCH)) for i in range(first, first + count))
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def upload_own_region(vol_host, stage_dev, starts, group=None):
    """Host-to-device copy of ONLY the x-slabs of the (pinned) host volume that this rank's patches read, into the same
    place of a full-size device staging volume.  X is the slowest axis, so a slab of one modality is one contiguous
    block: one plain asynchronous copy per (modality, slab) - a strided 4-D copy would be staged through a host-side
    repack.  Returns the number of bytes copied (18 patches on 8 ranks: 76 MB on most ranks instead of 143 MB)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if vol_host.dim() == 5:
        vol_host = vol_host[0]
    X = int(vol_host.shape[1])
    first, count = partition_patches(len(starts), world)[rank]
    starts_a = [starts[i] for i in assignment_order(starts)]
    total = 0
    for xa, xb in own_x_ranges(starts_a, first, count, X):
        for c in range(int(vol_host.shape[0])):
            stage_dev[c, xa:xb].copy_(vol_host[c, xa:xb], non_blocking=True)
        total += int(vol_host.shape[0]) * (xb - xa) * int(vol_host.shape[2]) * int(vol_host.shape[3]) * 4
    return total
