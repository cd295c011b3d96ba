"""The 13 discrete top-k selections of a patch (cls_wise_former.py:345-376, :552-555) are the one place where the result
is NOT a continuous function of the arithmetic: two token scores closer than the arithmetic's error may be selected in
either order, and a token that enters / leaves the set rewrites a 2x2x1 / 4x2x2 block of features downstream.  The
parity-grade modes must therefore select the reference's sets - and where they do not, the disagreement must be a true
near-tie of the REFERENCE's own fp32 scores (a case in which the reference on other hardware / another BLAS flips as
well: SURVEY H3 measured 6e-6 between oneDNN and native convolutions).  Six inputs, deterministic dropout mask."""
import pytest
import torch

from tests.util import rel_err, topk_disagreement

pytestmark = pytest.mark.gpu
TAGS = [f"{k}_{s}" for k in ("01", "02", "04") for s in ("ee", "es", "ss", "se")] + ["fusion"]
NEAR_TIE = 1e-4          # |score difference| / std(score) below which a selection is considered tied


@pytest.mark.parametrize("mode", ["F16X3", "FP32"])
def test_topk_sets_match_the_oracle_or_are_near_ties(mode, seed0_state_dict):
    import dcl_b200
    from oracle import clswiseformer_oracle as O
    eng = dcl_b200.Engine(dcl_b200.Precision[mode])
    eng.load_state_dict(seed0_state_dict)
    report = []
    try:
        for seed in (1, 2, 3, 4, 5, 6):
            torch.manual_seed(seed)
            x = torch.randn(1, 4, 128, 128, 128)
            stages = {}
            ref = O.forward(seed0_state_dict, x, torch.ones(1, 16), False, stages)[0]
            probs = eng.forward(x.cuda(), None).cpu()
            n_diff, margin = topk_disagreement(eng.read_topk(), stages, TAGS)
            err = rel_err(probs.numpy(), ref.numpy())
            flips = (probs[0].argmax(0) != ref[0].argmax(0)).float().mean().item()
            report.append((seed, n_diff, margin, err, flips))
            if n_diff == 0:
                assert err < 1e-3 and flips <= 1e-4, (mode, seed, err, flips)
            else:      # a disagreement is admissible only at a near-tie of the reference's own scores
                assert margin < NEAR_TIE, (mode, seed, n_diff, margin)
                assert err < 2e-2 and flips < 2e-3, (mode, seed, err, flips)
    finally:
        eng.close()
        print(f"[{mode}] (seed, differing selections, worst margin/std, probs rel err, label flips):")
        for r in report:
            print("   seed %d: %d  %.1e  %.2e  %.2e" % r)
    # the split-operand mode must not disagree more often than on one input in six
    assert sum(1 for r in report if r[1]) <= 1


def test_split_mode_agrees_with_the_ffma_mode_on_many_inputs(seed0_state_dict):
    """24 random patches with random dropout masks, split-fp16 (tcgen05) against the fp32 FFMA mode of the same library
    (no oracle in the loop, so many inputs are cheap): identical top-k sets and probabilities within 1e-4 on every
    input; label disagreement <= 1e-5 of the voxels on average."""
    import numpy as np
    import dcl_b200
    a = dcl_b200.Engine(dcl_b200.Precision.F16X3)
    b = dcl_b200.Engine(dcl_b200.Precision.FP32)
    a.load_state_dict(seed0_state_dict)
    b.load_state_dict(seed0_state_dict)
    rng = np.random.RandomState(5)
    worst, flips, set_diffs = 0.0, [], 0
    try:
        for seed in range(100, 124):
            torch.manual_seed(seed)
            x = (torch.randn(1, 4, 128, 128, 128) * (0.5 + 1.5 * rng.rand()) + rng.randn() * 0.3).cuda()
            keep = (rng.rand(16) < 0.8).astype(np.float32) / 0.8
            pa = a.forward(x, keep)
            ta = a.read_topk()
            pb = b.forward(x, keep)
            tb = b.read_topk()
            torch.cuda.synchronize()
            set_diffs += sum(1 for t in TAGS if set(ta[t].tolist()) != set(tb[t].tolist()))
            worst = max(worst, float((pa - pb).abs().max() / pb.abs().max()))
            flips.append(float((pa[0].argmax(0) != pb[0].argmax(0)).float().mean()))
    finally:
        a.close()
        b.close()
    print(f"split-fp16 vs FFMA over 24 inputs: worst probs rel err {worst:.2e}, mean label disagreement {np.mean(flips):.2e}, "
          f"max {np.max(flips):.2e}, differing top-k selections {set_diffs} of {24 * 13}")
    assert set_diffs == 0
    assert worst < 1e-4 and np.mean(flips) <= 1e-5
