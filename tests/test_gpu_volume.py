"""Sliding-window driver parity: stitch / accumulate / label / Dice kernels against the numpy oracle
(bit-exact for the integer and copy work) and against the reference goldens end to end."""
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_overlap50 import unpack2
from tests.util import check_digest, rel_err, volume_input, volume_target

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["FP32", "BF16X3"])
def engine(request, seed0_state_dict):
    import dcl_b200
    eng = dcl_b200.Engine(dcl_b200.Precision[request.param])
    eng.load_state_dict(seed0_state_dict)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def vol():
    return volume_input(0).cuda()


def _patch_probs(engine, vol, starts, keeps):
    out = []
    for (sx, sy, sz), k in zip(starts, keeps):
        p = engine.forward(vol[..., sx:sx + 128, sy:sy + 128, sz:sz + 128], k)
        out.append(p[0].cpu().numpy())
    return out


def test_reference_volume_end_to_end(engine, vol, golden_volume):
    from dcl_b200 import StitchMode
    from dcl_b200.engine import dice_from_counts
    g = golden_volume
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    out = engine.predict_volume(vol, StitchMode.REFERENCE, keep_scales=g["keep_scale"], target=tgt)
    torch.cuda.synchronize()
    assert out["probs"].shape == (1, 4, 240, 240, 155)
    check_digest("stitched", out["probs"], g, 1e-3)
    labels = out["labels"].cpu().numpy()
    counts = out["counts"].cpu().numpy()
    V = labels.size
    assert counts[:4].sum() == V
    assert np.abs(counts[:4] - g["labels_hist"]).sum() <= 2e-4 * V          # <= 1e-4 of voxels may flip
    samp = labels.ravel()[:: V // 4096][:4096]
    assert (samp != g["labels_sample"]).mean() <= 1e-3
    # the WHOLE label map against the unmodified reference's (tests/golden/make_golden_volume_labels.py): the 1e-4 budget itself
    full = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "volume_seed1000_labels.npz"))
    assert int(np.prod(full["shape"])) == V
    flips = float((labels.ravel() != unpack2(full["labels_packed"], V)).mean())
    print(f"reference tiling, {engine.precision.name}: label flips {flips:.2e} of {V} voxels")
    assert flips <= 1e-4
    assert np.allclose(dice_from_counts(counts), g["dice"], atol=1e-3)       # per-region Dice within 1e-3


@pytest.mark.parametrize("mode_name", ["REFERENCE", "ALIGNED"])
def test_stitch_and_labels_are_bit_exact(engine, vol, golden_volume, mode_name):
    """Copy / arg-max / counter kernels are integer-or-copy work: bit-exact against the oracle when fed
    the same per-patch probabilities."""
    from dcl_b200 import StitchMode
    from oracle import stitch_oracle as S
    keeps = golden_volume["keep_scale"]
    probs = _patch_probs(engine, vol, S.REFERENCE_STARTS, keeps)
    if mode_name == "REFERENCE":
        want = S.stitch_reference_from_probs(probs)
    else:   # aligned: z tail from patch-local 101:128
        want = S.stitch_reference_from_probs(probs)
        for p, (sx, sy, sz) in zip(probs, S.REFERENCE_STARTS):
            if sz == 27:
                xs = slice(0, 128) if sx == 0 else slice(128, 240)
                ys = slice(0, 128) if sy == 0 else slice(128, 240)
                px = slice(0, 128) if sx == 0 else slice(16, 128)
                py = slice(0, 128) if sy == 0 else slice(16, 128)
                want[:, xs, ys, 128:155] = p[:, px, py, 101:128]
    tgt_np = volume_target(0)
    out = engine.predict_volume(vol, StitchMode[mode_name], keep_scales=keeps,
                                target=torch.from_numpy(tgt_np.astype(np.uint8)).cuda())
    got = out["probs"][0].cpu().numpy()
    assert np.array_equal(got, want)
    labels = S.labels_from_probs(want)
    assert np.array_equal(out["labels"].cpu().numpy(), labels.astype(np.uint8))
    counts = out["counts"].cpu().numpy().tolist()
    assert counts[:4] == S.label_histogram(labels)
    assert counts[4:] == [v for c in S.region_counts(labels, tgt_np) for v in c]


@pytest.mark.parametrize("mode_name,stride", [("UNIFORM", 64), ("GAUSSIAN", 64), ("UNIFORM", 96)])
def test_weighted_accumulate_matches_oracle(engine, vol, mode_name, stride):
    from dcl_b200 import StitchMode, patch_starts
    from oracle import stitch_oracle as S
    starts = patch_starts((240, 240, 155), stride)
    assert starts == S.patch_starts((240, 240, 155), stride)
    keeps = np.ones((len(starts), 16), np.float32)
    probs = _patch_probs(engine, vol, starts, keeps)
    want = S.accumulate_from_probs(probs, starts, mode_name.lower())
    tgt_np = volume_target(0)
    out = engine.predict_volume(vol, StitchMode[mode_name], starts=starts, keep_scales=keeps,
                                target=torch.from_numpy(tgt_np.astype(np.uint8)).cuda())
    got = out["probs"][0].cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-6
    # labels are the arg-max of the library's own normalised probabilities (bit-exact), counters follow
    labels = S.labels_from_probs(got)
    assert np.array_equal(out["labels"].cpu().numpy(), labels.astype(np.uint8))
    counts = out["counts"].cpu().numpy().tolist()
    assert counts[:4] == S.label_histogram(labels)
    assert counts[4:] == [v for c in S.region_counts(labels, tgt_np) for v in c]
    assert np.abs(got.sum(0) - 1).max() < 1e-5       # blended probabilities still sum to one


@pytest.mark.parametrize("mode_name", ["UNIFORM", "GAUSSIAN"])
def test_gather_form_is_bit_identical_to_accumulate_form(engine, vol, mode_name, monkeypatch):
    """the default gather-form stitch (one slot per patch + gather_finalize_kernel) against accumulate + finalize"""
    from dcl_b200 import StitchMode, patch_starts
    starts = patch_starts((240, 240, 155), 64)
    keeps = np.ones((len(starts), 16), np.float32)
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    monkeypatch.setenv("DCL_GATHER", "0")
    ref = engine.predict_volume(vol, StitchMode[mode_name], starts=starts, keep_scales=keeps, target=tgt)
    monkeypatch.delenv("DCL_GATHER")
    got = engine.predict_volume(vol, StitchMode[mode_name], starts=starts, keep_scales=keeps, target=tgt)
    assert torch.equal(got["probs"], ref["probs"])
    assert torch.equal(got["labels"], ref["labels"]) and torch.equal(got["counts"], ref["counts"])
    only_labels = engine.predict_volume(vol, StitchMode[mode_name], starts=starts, keep_scales=keeps, want_probs=False)
    assert torch.equal(only_labels["labels"], ref["labels"])


def test_host_entry_point_equals_device_entry_point(engine, vol, golden_volume):
    from dcl_b200 import StitchMode
    keeps = golden_volume["keep_scale"]
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8))
    dev = engine.predict_volume(vol, StitchMode.REFERENCE, keep_scales=keeps, target=tgt.cuda())
    host = engine.predict_volume_host(volume_input(0)[0].pin_memory(), StitchMode.REFERENCE, keep_scales=keeps,
                                      target_host=tgt.pin_memory())
    assert np.array_equal(host["labels"].numpy(), dev["labels"].cpu().numpy())
    assert np.array_equal(host["counts"], dev["counts"].cpu().numpy())


def test_dropin_tailor_and_concat(seed0_state_dict, vol, golden_volume):
    """predict_overlap.tailor_and_concat(x, missing_modal, model) through the drop-in modules."""
    import predict_overlap
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    model = get_cls_wise_former("brats", True, "fixed", 0)
    model.load_state_dict(seed0_state_dict)
    model.eval()
    model.compute_aux = False
    model.deterministic = True
    with torch.no_grad():
        y = predict_overlap.tailor_and_concat(vol, None, model)
    assert y.shape == (1, 4, 240, 240, 155)
    assert (y.sum(1) - 1).abs().max().item() < 1e-5
    # with the always-on dropout replayed from the CPU generator state the golden volume is reproduced
    class Replay(type(model)):
        pass
    it = iter(torch.from_numpy(golden_volume["keep_scale"]))
    model.deterministic = False
    model.draw_keep_scale = lambda n, device: next(it).reshape(1, 16)
    with torch.no_grad():
        y = predict_overlap.tailor_and_concat(vol, None, model)
    check_digest("stitched", y, golden_volume, 1e-3)


def test_sharded_path_world1_equals_predict_volume(engine, vol):
    """dcl_accumulate_patches + dcl_finalize_labels (the building blocks of the multi-GPU path) reproduce
    dcl_predict_volume bit for bit when a single rank owns every patch."""
    from dcl_b200 import StitchMode, patch_starts, sharded
    starts = patch_starts((240, 240, 155), 96)
    keeps = np.ones((len(starts), 16), np.float32)
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    want = engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt, want_probs=False)
    for fn in (sharded.predict_volume_sharded, sharded.predict_volume_sharded_accumulate):     # owner-computes / reduce-scatter
        got = fn(engine, vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt)
        torch.cuda.synchronize()
        assert torch.equal(got["labels"], want["labels"]), fn.__name__
        assert torch.equal(got["counts"], want["counts"]), fn.__name__


def _sharded_rank(rank, world, port, out_dir, sd, precision="FP32"):
    import os
    import torch.distributed as dist
    import dcl_b200
    from dcl_b200 import StitchMode, patch_starts, sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        eng = dcl_b200.Engine(dcl_b200.Precision[precision])
        eng.load_state_dict(sd)
        starts = patch_starts((240, 240, 155), 96)
        keeps = np.ones((len(starts), 16), np.float32)
        v = volume_input(0).cuda()
        tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
        out = sharded.predict_volume_sharded(eng, v, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt)
        out2 = sharded.predict_volume_sharded(eng, v, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt)   # reuse of the slots
        acc = sharded.predict_volume_sharded_accumulate(eng, v, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt)
        torch.cuda.synchronize()
        assert torch.equal(out["labels"], out2["labels"]) and torch.equal(out["counts"], out2["counts"])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), labels=out["labels"].cpu().numpy(),
                 counts=out["counts"].cpu().numpy(), acc_labels=acc["labels"].cpu().numpy(), acc_counts=acc["counts"].cpu().numpy())
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_two_gpus_nccl_equals_one_gpu(engine, vol, seed0_state_dict, tmp_path):
    """One volume, patches split over 2 ranks.  Owner-computes form (per-rank slots, peer reads over NVLink inside
    gather_finalize_kernel): the label map and the counters EQUAL the single-GPU result bit for bit.  Accumulate form
    (NCCL reduce-scatter of the accumulators): may differ only where fp32 re-association of the overlap sums flips a
    near-tie (<= 1e-4 of voxels)."""
    import socket
    import torch.multiprocessing as mp
    from dcl_b200 import StitchMode, patch_starts
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    sd = {k: v.cpu() for k, v in seed0_state_dict.items()}
    mp.spawn(_sharded_rank, args=(2, port, str(tmp_path), sd, engine.precision.name), nprocs=2, join=True)
    starts = patch_starts((240, 240, 155), 96)
    keeps = np.ones((len(starts), 16), np.float32)
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    want = engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt, want_probs=False)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["labels"], r1["labels"]) and np.array_equal(r0["counts"], r1["counts"])
    wl = want["labels"].cpu().numpy()
    assert np.array_equal(r0["labels"], wl) and np.array_equal(r0["counts"], want["counts"].cpu().numpy())
    assert (r0["acc_labels"] != wl).mean() <= 1e-4
    assert np.abs(r0["acc_counts"][:4] - want["counts"].cpu().numpy()[:4]).sum() <= 2e-4 * wl.size


def test_tta_matches_oracle_algebra(engine, vol):
    """dcl_predict_volume_tta (flip -> tiling -> un-flip -> softmax -> mean, predict_cls.py:180-203) against the oracle
    algebra fed with the library's own stitched outputs of the flipped volumes."""
    from dcl_b200 import StitchMode
    from oracle import stitch_oracle as S
    keeps = np.ones((8, 8, 16), np.float32)
    keeps[3, 2, 5] = 0.0
    x155 = vol[..., :155]
    stitched = []
    for f, dims in enumerate(S.TTA_FLIPS):
        xf = x155.flip(dims=dims).contiguous() if dims else x155.contiguous()     # test-side input preparation only
        stitched.append(engine.predict_volume(xf, StitchMode.REFERENCE, keep_scales=keeps[f])["probs"][0].cpu().numpy())
    want = S.tta_average_from_stitched(stitched)
    tgt_np = volume_target(0)
    out = engine.predict_volume_tta(vol, keep_scales=keeps, target=torch.from_numpy(tgt_np.astype(np.uint8)).cuda())
    got = out["probs"][0].cpu().numpy()
    assert got.shape == (4, 240, 240, 155)
    assert np.abs(got - want).max() < 1e-6
    labels = S.labels_from_probs(got)
    assert np.array_equal(out["labels"].cpu().numpy(), labels.astype(np.uint8))
    counts = out["counts"].cpu().numpy().tolist()
    assert counts[:4] == S.label_histogram(labels)
    assert counts[4:] == [v for c in S.region_counts(labels, tgt_np) for v in c]


def test_tta_reference_golden(engine, vol):
    """End to end against the UNMODIFIED reference run on CPU (tests/golden/make_golden_tta.py): averaged probabilities
    within 1e-3, <= 1e-4 of the voxels with a different label, Dice within 1e-3."""
    import os
    from dcl_b200.engine import dice_from_counts
    path = os.path.join(os.path.dirname(__file__), "golden", "volume_tta_seed1000.npz")
    if not os.path.exists(path):
        pytest.skip("TTA golden not generated")
    g = np.load(path)
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    out = engine.predict_volume_tta(vol, keep_scales=g["keep_scale"].reshape(8, 8, 16), target=tgt)
    torch.cuda.synchronize()
    check_digest("tta", out["probs"], g, 1e-3)
    counts = out["counts"].cpu().numpy()
    V = 240 * 240 * 155
    assert np.abs(counts[:4] - g["labels_hist"]).sum() <= 2e-4 * V
    samp = out["labels"].cpu().numpy().ravel()[:: V // 4096][:4096]
    assert (samp != g["labels_sample"]).mean() <= 1e-3
    assert np.allclose(dice_from_counts(counts), g["dice"], atol=1e-3)


def test_bad_arguments_are_refused(engine, vol):
    """A patch list that leaves voxels uncovered (weight sum 0 -> 0/0), a batch of more than one volume and a target of
    the wrong shape are errors, not silent garbage."""
    import dcl_b200
    from dcl_b200 import StitchMode, patch_starts
    starts = patch_starts((240, 240, 155), 64)
    with pytest.raises(dcl_b200.DclError, match="does not cover"):
        engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts[:-1])
    with pytest.raises(dcl_b200.DclError, match="does not cover"):
        engine.predict_volume(vol, StitchMode.GAUSSIAN, starts=[(0, 0, 0), (112, 112, 27)])
    with pytest.raises(dcl_b200.DclError, match="batch of 2"):
        engine.predict_volume(torch.cat([vol, vol]), StitchMode.REFERENCE)
    with pytest.raises(dcl_b200.DclError, match="target"):
        engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts, target=torch.zeros(240, 240, 150, dtype=torch.uint8))
    out = engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts, want_probs=False)      # the handle still works
    assert out["labels"].shape == (240, 240, 155)


def test_owner_computes_building_blocks_on_one_gpu(engine, vol):
    """dcl_forward_patches_to_slots + dcl_gather_finalize_range (the multi-GPU owner-computes blocks) on ONE GPU: the
    volume blended in three x-ranges from explicit slot pointers - the second half of the patches living in a second
    buffer, as a peer's slots would - equals dcl_predict_volume bit for bit (labels, counters, probabilities)."""
    from dcl_b200 import StitchMode, patch_starts
    starts = patch_starts((240, 240, 155), 96)
    n = len(starts)
    keeps = np.ones((n, 16), np.float32)
    keeps[2, 4] = 0.0
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8)).cuda()
    want = engine.predict_volume(vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt)
    slot_bytes = 4 * 128 ** 3 * 4
    base = engine.slots_ensure(n)
    engine.forward_patches_to_slots(vol, StitchMode.UNIFORM, starts, keeps, 0, n)
    torch.cuda.synchronize()
    other = torch.empty((n - n // 2) * slot_bytes // 4, dtype=torch.float32, device="cuda")      # "peer" memory
    import ctypes as C
    from dcl_b200 import _native as N
    cudart = C.CDLL("libcudart.so")
    assert cudart.cudaMemcpy(C.c_void_p(other.data_ptr()), C.c_void_p(base + (n // 2) * slot_bytes),
                             C.c_size_t(other.numel() * 4), 3) == 0          # cudaMemcpyDeviceToDevice
    ptrs = [base + i * slot_bytes if i < n // 2 else other.data_ptr() + (i - n // 2) * slot_bytes for i in range(n)]
    labels = torch.zeros((240, 240, 155), dtype=torch.uint8, device="cuda")
    probs = torch.zeros((1, 4, 240, 240, 155), dtype=torch.float32, device="cuda")
    counts = torch.zeros(13, dtype=torch.int64, device="cuda")
    for x0, x1 in ((0, 77), (77, 78), (78, 240)):
        engine.gather_finalize_range((240, 240, 155), StitchMode.UNIFORM, starts, ptrs, x0, x1, labels, target=tgt, counts=counts,
                                     probs_out=probs)
    torch.cuda.synchronize()
    assert torch.equal(labels, want["labels"]) and torch.equal(counts, want["counts"]) and torch.equal(probs, want["probs"])
    _ = N
