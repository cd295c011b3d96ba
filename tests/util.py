"""Shared helpers of the test-suite."""
import numpy as np
import torch

N_SAMPLE = 4096


def strided_sample(t):
    """The sub-sampling rule of tests/golden/make_golden.py::digest."""
    a = t.detach().flatten()
    step = max(1, a.numel() // N_SAMPLE)
    return a[::step][:N_SAMPLE].float().cpu().numpy()


def rel_err(a, b):
    """max|a-b| / max|b| (the per-tensor relative error SURVEY H3 recommends)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def check_digest(name, t, golden, tol):
    g = {k: golden[f"{name}/{k}"] for k in ("shape", "mean", "std", "absmax", "l2", "sample")}
    assert tuple(t.shape) == tuple(int(v) for v in g["shape"]), name
    err = rel_err(strided_sample(t), g["sample"])
    assert err <= tol, f"{name}: sampled rel err {err:.3e} > {tol}"
    a = t.detach().double().flatten()
    l2 = float(a.norm())
    assert abs(l2 - float(g["l2"])) <= tol * float(g["l2"]) + 1e-12, f"{name}: l2 {l2} vs {float(g['l2'])}"
    return err


def config1_input():
    torch.manual_seed(1)
    return torch.randn(1, 4, 128, 128, 128)


def volume_input(i=0):
    torch.manual_seed(1000 + i)
    return torch.randn(1, 4, 240, 240, 155)


def volume_target(i=0):
    return np.random.RandomState(i).randint(0, 4, (240, 240, 155))


COUPLER_ORDER = {"coupler_01": ("01_ee", "01_ss"), "coupler_02": ("02_ee", "02_ss"), "coupler_04": ("04_ee", "04_ss"),
                 "coupler_fusion": ("fusion",)}


def coupler_rows_in_golden_order(name, t, topk, golden):
    """A coupler output is (1, blocks x 129, 512): per block one class-token row, then one row per SELECTED token in the
    order torch.topk returned them (cls_wise_former.py:345-376).  Only the selected SET is defined by the algorithm's
    result (attention is permutation invariant in its keys and the scatter-back goes by index, :458-543); two
    near-tied scores may come out in either order.  This puts our token rows into the golden's order (rows whose token
    the golden did not select are reported back) so that stage digests can be compared row by row."""
    t = t.reshape(-1, 512).clone()
    missing = 0
    for b, tag in enumerate(COUPLER_ORDER[name]):
        ours = topk[tag].tolist()
        want = golden["topk_" + tag].tolist()
        pos = {tok: i for i, tok in enumerate(ours)}
        base = b * 129 + 1
        block = t[base:base + 128].clone()
        for i, tok in enumerate(want):
            if tok in pos:
                t[base + i] = block[pos[tok]]
            else:
                missing += 1
                t[base + i] = float("nan")
    return t.reshape(1, -1, 512), missing


def topk_disagreement(ours, oracle_stages, tags):
    """Compares our 13 top-k index sets with the oracle's.  Returns (n_differing_selections, worst_margin): for every
    token that one side selected and the other did not, |score(token) - score(128th selected)| / std(score) in the
    ORACLE's fp32 scores - how close to a tie the disagreement is (0 = exact tie)."""
    n_diff, worst = 0, 0.0
    for tag in tags:
        a, b = set(int(v) for v in ours[tag]), set(int(v) for v in oracle_stages["topk_" + tag].tolist())
        if a == b:
            continue
        n_diff += 1
        sc = oracle_stages["score_" + tag].double()
        kth = sc[oracle_stages["topk_" + tag][-1]]
        for tok in a ^ b:
            worst = max(worst, float((sc[tok] - kth).abs() / sc.std()))
    return n_diff, worst
