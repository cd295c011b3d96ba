"""Host logic of the single-volume multi-GPU path (dcl_b200/sharded.py) under the `gloo` backend with
world_size 2 on CPU: patch partition, accumulator exchange, owned-range finalise, label gather, counter reduce.
The CUDA kernels are replaced by oracle stand-ins (tests may use the oracle); the result must equal the
single-process oracle bit for bit and must not depend on the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcl_b200 import StitchMode, sharded
from oracle import stitch_oracle as S

SHAPE = (144, 144, 136)


def _patch_probs(i):
    """Deterministic per-patch 'probabilities' on a 1/64 grid, so that fp32 sums are exact in any order."""
    return (np.random.RandomState(100 + i).randint(0, 64, (4, 128, 128, 128)) / 64.0).astype(np.float32)


def _cpu_accumulate(vol, mode, starts, keep_scales, first, count, acc, wsum):
    X, Y, Z = SHAPE
    a = acc.view(4, X, Y, Z).numpy()
    w = wsum.view(X, Y, Z).numpy()
    for i in range(first, first + count):
        sx, sy, sz = starts[i]
        a[:, sx:sx + 128, sy:sy + 128, sz:sz + 128] += _patch_probs(i)
        w[sx:sx + 128, sy:sy + 128, sz:sz + 128] += 1.0


def _cpu_finalize(acc_l, wsum_l, labels_l, tgt_l, counts):
    probs = acc_l.numpy() / wsum_l.numpy()[None]
    lab = probs.argmax(0)
    labels_l.copy_(torch.from_numpy(lab.astype(np.uint8)))
    c = [int((lab == k).sum()) for k in range(4)]
    if tgt_l is not None:
        c += [v for trip in S.region_counts(lab, tgt_l.numpy()) for v in trip]
    else:
        c += [0] * 9
    counts += torch.tensor(c, dtype=torch.int64)


def _cpu_forward_to_slots(vol, mode, starts, keep_scales, first, count):
    """`starts` arrives in the assignment order of sharded.py; the stand-in probabilities go by PLAN index."""
    plan = S.patch_starts(SHAPE, 16)
    ids = [plan.index(tuple(starts[k])) for k in range(first, first + count)]
    return torch.from_numpy(np.stack([_patch_probs(i) for i in ids])) if count else torch.zeros((0, 4, 128, 128, 128))


def _cpu_finalize_range(slots, x0, x1, labels, target, counts):
    """Owner-computes stand-in: blend of the rows x0 <= x < x1 from per-patch slots (here tensors that travelled by
    all_gather; on the GPU pointers into local / peer memory), in patch order like gather_finalize_kernel."""
    X, Y, Z = SHAPE
    starts = S.patch_starts(SHAPE, 16)
    acc = np.zeros((4, x1 - x0, Y, Z), np.float32)
    w = np.zeros((x1 - x0, Y, Z), np.float32)
    for p, (sx, sy, sz) in zip(slots, starts):
        a, b = max(sx, x0), min(sx + 128, x1)
        if a < b:
            acc[:, a - x0:b - x0, sy:sy + 128, sz:sz + 128] += p.numpy()[:, a - sx:b - sx]
            w[a - x0:b - x0, sy:sy + 128, sz:sz + 128] += 1.0
    lab = (acc / w[None]).argmax(0)
    labels[x0:x1] = torch.from_numpy(lab.astype(np.uint8))
    c = [int((lab == k).sum()) for k in range(4)]
    c += [v for trip in S.region_counts(lab, target[x0:x1].numpy()) for v in trip] if target is not None else [0] * 9
    counts += torch.tensor(c, dtype=torch.int64)


def _expected(starts, target):
    want = S.accumulate_from_probs([_patch_probs(i) for i in range(len(starts))], starts, "uniform", shape=SHAPE)
    lab = S.labels_from_probs(want)
    counts = S.label_histogram(lab) + [v for trip in S.region_counts(lab, target) for v in trip]
    return lab.astype(np.uint8), counts


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        starts = S.patch_starts(SHAPE, 16)
        target = np.random.RandomState(7).randint(0, 4, SHAPE)
        vol = torch.zeros((4,) + SHAPE)
        out = sharded.predict_volume_sharded_accumulate(None, vol, StitchMode.UNIFORM, starts=starts,
                                                        target=torch.from_numpy(target), accumulate=_cpu_accumulate,
                                                        finalize=_cpu_finalize)
        own = sharded.predict_volume_sharded(None, vol, StitchMode.UNIFORM, starts=starts, target=torch.from_numpy(target),
                                             forward_to_slots=_cpu_forward_to_slots, finalize_range=_cpu_finalize_range)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), labels=out["labels"].numpy(), counts=out["counts"].numpy(),
                 patches=np.array(out["patches"]), own_labels=own["labels"].numpy(), own_counts=own["counts"].numpy(),
                 own_rows=np.array(own["rows"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_owner_ranges_and_patch_owners():
    for world in (1, 2, 3, 7, 8):
        for X in (240, 144, 10):
            rs = [sharded.owned_x_range(X, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == X
            for (a, b), (c, _) in zip(rs, rs[1:]):
                assert a <= b == c
    assert sharded.owned_x_range(240, 7, 8) == (210, 240)
    owners = sharded.patch_owner(18, 8)
    assert owners[:4] == [(0, 0), (0, 1), (0, 2), (1, 0)] and owners[-1] == (7, 1) and len(owners) == 18


def test_own_x_ranges_merge_overlaps():
    starts = S.patch_starts((240, 240, 155), 64)
    parts = sharded.partition_patches(len(starts), 8)
    rs = [sharded.own_x_ranges(starts, f, c, 240) for f, c in parts]
    assert rs[0] == [(0, 128)] and rs[3] == [(0, 240)] and rs[7] == [(112, 240)]
    assert sharded.own_x_ranges([(0, 0, 0), (200, 0, 0)], 0, 2, 400) == [(0, 128), (200, 328)]
    assert sharded.own_x_ranges(starts, 0, 0, 240) == []


def test_partition_and_ranges():
    assert sharded.partition_patches(18, 8) == [(0, 3), (3, 3), (6, 2), (8, 2), (10, 2), (12, 2), (14, 2), (16, 2)]
    assert sharded.partition_patches(8, 8) == [(i, 1) for i in range(8)]
    assert sharded.partition_patches(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]
    for world in (1, 2, 3, 8):
        for total in (8928000, 1001):
            ranges = [sharded.owned_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and sum(n for _, n in ranges) == total
            for (a, n), (b, _) in zip(ranges, ranges[1:]):
                assert a + n == b
    with pytest.raises(Exception):
        sharded.partition_patches(4, 0)


def test_two_rank_gloo_equals_single_process_oracle(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    starts = S.patch_starts(SHAPE, 16)
    assert len(starts) == 8
    target = np.random.RandomState(7).randint(0, 4, SHAPE)
    want_labels, want_counts = _expected(starts, target)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert got[0]["patches"].tolist() == [0, 4] and got[1]["patches"].tolist() == [4, 4]
    for g in got:                                   # every rank ends with the full, identical result - in both forms
        assert np.array_equal(g["labels"], want_labels)
        assert g["counts"].tolist() == want_counts
        assert np.array_equal(g["own_labels"], want_labels)
        assert g["own_counts"].tolist() == want_counts
    assert got[0]["own_rows"].tolist() == [0, 72] and got[1]["own_rows"].tolist() == [72, 144]


def test_single_process_path_needs_no_process_group():
    starts = S.patch_starts(SHAPE, 16)[:2]
    vol = torch.zeros((4,) + SHAPE)
    out = sharded.predict_volume_sharded_accumulate(None, vol, StitchMode.UNIFORM, starts=starts, accumulate=_cpu_accumulate,
                                                    finalize=lambda a, w, l, t, c: l.zero_())
    assert out["patches"] == (0, 2) and out["labels"].shape == SHAPE
    full = S.patch_starts(SHAPE, 16)
    own = sharded.predict_volume_sharded(None, vol, StitchMode.UNIFORM, starts=full, forward_to_slots=_cpu_forward_to_slots,
                                         finalize_range=_cpu_finalize_range)
    want_labels, _ = _expected(full, np.zeros(SHAPE, np.int64))
    assert own["patches"] == (0, 8) and own["rows"] == (0, 144) and np.array_equal(own["labels"].numpy(), want_labels)
    assert sorted(own["patch_ids"]) == list(range(8))


def test_assignment_order_groups_patches_by_x():
    starts = S.patch_starts((240, 240, 155), 64)
    order = sharded.assignment_order(starts)
    assert sorted(order) == list(range(18)) and [starts[i][0] for i in order] == [0] * 6 + [64] * 6 + [112] * 6
    dealt = [starts[i] for i in order]
    for rank, (f, c) in enumerate(sharded.partition_patches(18, 8)):
        assert sharded.own_x_ranges(dealt, f, c, 240) in ([(0, 128)], [(64, 192)], [(112, 240)]), rank
