"""bf16 mode (tcgen05 kernels, bf16 operands / fp32 accumulate) against the fp32 reference goldens.
Tolerance: north_star states 2e-2 for bf16 logits; measured per tensor as max|a-b| / max|b| (SURVEY H3:
the discrete top-k selections make element-wise relative error meaningless under low precision)."""
import numpy as np
import pytest
import torch

from tests.test_gpu_forward import STAGES, STAGE_SHAPES
from tests.util import config1_input, rel_err, strided_sample

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def engine_bf16(seed0_state_dict):
    import dcl_b200
    eng = dcl_b200.Engine(dcl_b200.Precision.BF16, want_aux=False, keep_stages=True)
    eng.load_state_dict(seed0_state_dict)
    yield eng
    eng.close()


def test_bf16_forward_within_tolerance_of_reference_goldens(engine_bf16, golden_patch):
    g = golden_patch
    probs = engine_bf16.forward(config1_input().cuda(), g["keep_scale"])
    torch.cuda.synchronize()
    errs = {}
    for name in STAGES:
        t = engine_bf16.read_stage(name).reshape(STAGE_SHAPES[name])
        errs[name] = rel_err(strided_sample(t), g[f"{name}/sample"])
    errs["probs"] = rel_err(strided_sample(probs), g["probs/sample"])
    overlap = {tag: len(set(idx.tolist()) & set(g["topk_" + tag].tolist())) for tag, idx in engine_bf16.read_topk().items()}
    print("bf16 stage rel errs:", {k: f"{v:.1e}" for k, v in errs.items()})
    print("bf16 top-k overlap (of 128):", overlap)
    assert errs["probs"] < BF16_TOL, errs
    for name in ("init", "x1_1", "x2_1", "x3_1", "x4"):      # encoder: no discrete selection upstream
        assert errs[name] < BF16_TOL, (name, errs[name])
    assert min(overlap.values()) >= 112                       # at most a few boundary tokens swap
    assert (probs.sum(1) - 1).abs().max().item() < 1e-5
    lab = probs[0].argmax(0).to(torch.uint8).cpu().numpy()
    hist = np.bincount(lab.ravel(), minlength=4)
    print("bf16 label histogram:", hist.tolist(), "reference:", g["labels_hist"].tolist())


def test_bf16_forward_is_deterministic(engine_bf16, golden_patch):
    x = config1_input().cuda()
    a = engine_bf16.forward(x, golden_patch["keep_scale"])
    b = engine_bf16.forward(x, golden_patch["keep_scale"])
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_bf16_volume_graph_replay_matches_per_patch_forwards(engine_bf16):
    """The sliding-window driver in bf16 mode replays the forward as a CUDA graph with per-patch arguments in device
    memory, accumulates with the vectorised kernel and (host entry point) uploads the volume in x-slabs while the first
    patches already run.  All of that must be pure plumbing: the blended probabilities equal the oracle blend of the
    per-patch forwards (same kernels, bit-reproducible), host and device entry points agree bit for bit."""
    from dcl_b200 import StitchMode, patch_starts
    from oracle import stitch_oracle as S
    from tests.util import volume_input, volume_target
    vol_h = volume_input(0)
    vol = vol_h.cuda()
    starts = patch_starts((240, 240, 155), 96)
    keeps = np.ones((len(starts), 16), np.float32)
    keeps[1, 3] = 0.0
    keeps[2, 7] = 1.25
    probs = []
    for (sx, sy, sz), k in zip(starts, keeps):
        probs.append(engine_bf16.forward(vol[..., sx:sx + 128, sy:sy + 128, sz:sz + 128], k)[0].cpu().numpy())
    want = S.accumulate_from_probs(probs, starts, "uniform")
    tgt = torch.from_numpy(volume_target(0).astype(np.uint8))
    dev = engine_bf16.predict_volume(vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt.cuda())
    got = dev["probs"][0].cpu().numpy()
    assert np.abs(got - want).max() < 2e-6
    labels = S.labels_from_probs(got)
    assert np.array_equal(dev["labels"].cpu().numpy(), labels.astype(np.uint8))
    host = engine_bf16.predict_volume_host(vol_h[0].pin_memory(), StitchMode.UNIFORM, starts=starts, keep_scales=keeps,
                                           target_host=tgt.pin_memory())
    assert np.array_equal(host["labels"].numpy(), dev["labels"].cpu().numpy())
    assert np.array_equal(host["counts"], dev["counts"].cpu().numpy())
    # the accumulate form (DCL_GATHER=0, what the multi-GPU path uses) gives the same bits as the default gather form
    import os
    os.environ["DCL_GATHER"] = "0"
    try:
        acc = engine_bf16.predict_volume(vol, StitchMode.UNIFORM, starts=starts, keep_scales=keeps, target=tgt.cuda())
    finally:
        del os.environ["DCL_GATHER"]
    assert torch.equal(acc["probs"], dev["probs"]) and torch.equal(acc["labels"], dev["labels"])
    assert torch.equal(acc["counts"], dev["counts"])


def test_bf16_production_schedule_equals_stage_keeping_schedule(engine_bf16, seed0_state_dict, golden_patch):
    """Without keep_stages the last DeBlock tail is folded into endconv's load; the probabilities must not change by a bit."""
    import dcl_b200
    eng = dcl_b200.Engine(dcl_b200.Precision.BF16)
    eng.load_state_dict(seed0_state_dict)
    x = config1_input().cuda()
    a = engine_bf16.forward(x, golden_patch["keep_scale"])
    b = eng.forward(x, golden_patch["keep_scale"])      # eager launches
    c = eng.forward(x, golden_patch["keep_scale"])      # captured graph
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(b, c)
    eng.close()
