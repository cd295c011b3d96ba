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
