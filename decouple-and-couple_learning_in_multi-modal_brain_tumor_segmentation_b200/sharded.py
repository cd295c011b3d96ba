"""One volume sharded over the GPUs of a box (BASELINE.json configs 3 and 5).

A patch is indivisible (InstanceNorm and the top-k selections are global over it), so the unit of
sharding is the patch: the z-major patch list is cut into `world` contiguous chunks ("slabs"), each rank
accumulates its patches into a private full-size fp32 accumulator (`dcl_accumulate_patches`), ONE exchange
step sums the overlapped logits so that rank g owns voxel range g (NCCL reduce-scatter over NVLink; an
all-reduce where the backend has no reduce-scatter), the owner normalises + arg-maxes + counts its range
(`dcl_finalize_labels`), and the uint8 labels are all-gathered / the 13 counters all-reduced.

The reference has no counterpart (its inference is single-GPU, test_overlap.py:78); the stitched result is
defined by predict_overlap.py:31-58 and must not depend on the number of ranks, which the tests check.
The arithmetic callables default to the CUDA engine; tests inject CPU stand-ins to exercise this host
logic under the `gloo` backend.
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


def predict_volume_sharded(engine, vol, mode=StitchMode.UNIFORM, starts=None, keep_scales=None, target=None,
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
