"""Per-operator parity on the GPU: the library's kernels against torch.nn.functional (CPU fp32)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _conv_case(c0, c1, cout, dims, stride, norm, act, residual, seed, impl=0):
    from dcl_b200.engine import op_conv3d_k3
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(c0, *dims, generator=g)
    x1 = torch.randn(c1, *dims, generator=g) if c1 else None
    cin = c0 + c1
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    b = torch.randn(cout, generator=g)
    xin = torch.cat([x0, x1], 0) if c1 else x0
    mean = rstd = None
    ref_in = xin
    if norm:
        mean = torch.randn(cin, generator=g) * 0.3
        rstd = torch.rand(cin, generator=g) + 0.5
        ref_in = (xin - mean.view(-1, 1, 1, 1)) * rstd.view(-1, 1, 1, 1)
    if act == 1:
        ref_in = F.relu(ref_in)
    elif act == 2:
        ref_in = F.leaky_relu(ref_in, 0.01)
    ref = F.conv3d(ref_in[None], w, b, stride=stride, padding=1)[0]
    res = torch.randn(ref.shape, generator=g) if residual else None
    if residual:
        ref = ref + res
    dev = "cuda"
    y = op_conv3d_k3(x0.to(dev), w.to(dev), b.to(dev), x1.to(dev) if c1 else None, stride,
                     (mean.to(dev), rstd.to(dev)) if norm else None, act, res.to(dev) if residual else None, impl=impl)
    torch.cuda.synchronize()
    return y.cpu(), ref


@pytest.mark.parametrize("case", [
    dict(c0=4, c1=0, cout=16, dims=(32, 32, 32), stride=1, norm=False, act=0, residual=False),
    dict(c0=16, c1=0, cout=16, dims=(32, 32, 64), stride=1, norm=True, act=1, residual=True),
    dict(c0=16, c1=0, cout=32, dims=(32, 32, 32), stride=2, norm=False, act=0, residual=False),
    dict(c0=32, c1=64, cout=96, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),
    dict(c0=32, c1=0, cout=8, dims=(16, 16, 16), stride=1, norm=True, act=2, residual=False),
    dict(c0=8, c1=0, cout=2, dims=(20, 12, 36), stride=1, norm=False, act=0, residual=False),
    dict(c0=5, c1=3, cout=19, dims=(9, 17, 18), stride=2, norm=True, act=2, residual=True),
    dict(c0=128, c1=0, cout=128, dims=(16, 16, 16), stride=1, norm=True, act=1, residual=True),
])
def test_conv3d_k3_fp32(case):
    y, ref = _conv_case(seed=7, **case)
    assert y.shape == ref.shape
    assert rel_err(y.numpy(), ref.numpy()) < 2e-5


def test_instnorm_stats():
    from dcl_b200.engine import op_instnorm_stats
    g = torch.Generator().manual_seed(3)
    for shape, shift in (((16, 64, 64, 64), 0.0), ((96, 32, 32, 32), 5.0), ((7, 9, 11, 13), -40.0)):
        x = torch.randn(shape, generator=g) * 0.7 + shift
        mean, rstd = op_instnorm_stats(x.cuda())
        xd = x.double().flatten(1)
        m = xd.mean(1)
        r = 1.0 / torch.sqrt(xd.var(1, unbiased=False) + 1e-5)
        std = xd.std(1, unbiased=False)
        assert ((mean.cpu().double() - m).abs() / (m.abs() + std)).max().item() < 1e-6
        assert rel_err(rstd.cpu().numpy(), r.numpy()) < 1e-5


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("c,g,act,residual", [(16, 128, 1, True), (16, 128, 2, False), (32, 64, 2, True), (32, 64, 0, False)])
def test_conv3d_k3_tcgen05_bf16(c, g, act, residual):
    """The tcgen05 rolling kernel against torch conv3d.  The kernel works on B-format (bf16) activations:
    input and residual are rounded to bf16 on entry, the normalised input is rounded again before the MMA,
    accumulation is fp32 and the output is stored as bf16 - the reference applies the same roundings, so what
    is left is summation order plus one bf16 ulp (2^-9) of the output.  Also checks the fused statistics."""
    from dcl_b200.engine import op_conv3d_k3
    gen = torch.Generator().manual_seed(11 + c)
    x = torch.randn(c, g, g, g, generator=gen) * 1.7 + 0.4
    w = torch.randn(c, c, 3, 3, 3, generator=gen) / (c * 27) ** 0.5
    b = torch.randn(c, generator=gen)
    mean = torch.randn(c, generator=gen) * 0.3
    rstd = torch.rand(c, generator=gen) + 0.5
    xin = (_bf16_round(x) - mean.view(-1, 1, 1, 1)) * rstd.view(-1, 1, 1, 1)
    xin = F.relu(xin) if act == 1 else (F.leaky_relu(xin, 0.01) if act == 2 else xin)
    ref = F.conv3d(_bf16_round(xin)[None], _bf16_round(w), b, padding=1)[0]
    res = torch.randn(ref.shape, generator=gen) if residual else None
    if residual:
        ref = ref + _bf16_round(res)
    stats = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    y = op_conv3d_k3(x.cuda(), w.cuda(), b.cuda(), None, 1, (mean.cuda(), rstd.cuda()), act,
                     res.cuda() if residual else None, impl=2, stats=stats)
    torch.cuda.synchronize()
    y = y.cpu()
    assert rel_err(y.numpy(), ref.numpy()) < 4e-3
    assert float((y - ref).abs().mean() / ref.abs().mean()) < 2e-3
    assert float((y - _bf16_round(ref)).abs().mean() / ref.abs().mean()) < 2e-4     # identical up to rare ulp flips
    st = stats.cpu().view(c, 2)
    rd = ref.double().flatten(1)
    assert rel_err(st[:, 0].numpy(), rd.sum(1).numpy()) < 1e-4
    assert rel_err(st[:, 1].numpy(), (rd * rd).sum(1).numpy()) < 1e-4


@pytest.mark.parametrize("case", [
    dict(c0=4, c1=0, cout=16, dims=(32, 32, 32), stride=1, norm=False, act=0, residual=False),
    dict(c0=16, c1=0, cout=32, dims=(32, 32, 32), stride=2, norm=False, act=0, residual=False),
    dict(c0=32, c1=64, cout=96, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),
    dict(c0=32, c1=0, cout=8, dims=(16, 16, 16), stride=1, norm=True, act=2, residual=False),
    dict(c0=8, c1=0, cout=2, dims=(20, 12, 36), stride=1, norm=False, act=0, residual=False),
    dict(c0=5, c1=3, cout=19, dims=(9, 17, 18), stride=2, norm=True, act=2, residual=True),
    dict(c0=128, c1=0, cout=128, dims=(16, 16, 16), stride=1, norm=True, act=1, residual=True),
    dict(c0=64, c1=0, cout=64, dims=(32, 32, 32), stride=1, norm=True, act=2, residual=False),
    dict(c0=256, c1=0, cout=384, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),
    dict(c0=32, c1=0, cout=32, dims=(6, 64, 64), stride=1, norm=True, act=1, residual=True),
    dict(c0=16, c1=0, cout=16, dims=(5, 4, 128), stride=1, norm=True, act=2, residual=False),
    dict(c0=128, c1=0, cout=256, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),
    dict(c0=16, c1=0, cout=32, dims=(128, 128, 128), stride=2, norm=False, act=0, residual=False),   # rolling stride-2 kernel
    dict(c0=32, c1=0, cout=64, dims=(64, 64, 64), stride=2, norm=False, act=0, residual=False),   # EnDown2 (rolling stride-2 kernel unless DCL_S2GEN=0)
    dict(c0=32, c1=0, cout=32, dims=(64, 64, 64), stride=2, norm=False, act=0, residual=False),   # conv_64_to_32
])
def test_conv3d_k3_gemm_bf16(case):
    """The general tcgen05 implicit-GEMM kernel (prep + cp.async im2col) against torch conv3d on
    bf16-rounded operands."""
    from dcl_b200.engine import op_conv3d_k3
    g = torch.Generator().manual_seed(23)
    c0, c1, cout, dims, stride = case["c0"], case["c1"], case["cout"], case["dims"], case["stride"]
    cin = c0 + c1
    x0 = torch.randn(c0, *dims, generator=g)
    x1 = torch.randn(c1, *dims, generator=g) if c1 else None
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    b = torch.randn(cout, generator=g)
    xin = torch.cat([x0, x1], 0) if c1 else x0
    # stride-1 convs on 16/32/64/128-wide rows run on the slab kernel, which stages the RAW input as bf16 and
    # applies the norm in shared memory (one more bf16 rounding than the im2col GEMM path, whose prep kernel
    # normalises in fp32 first)
    slab = stride == 1 and dims[2] in (16, 32, 64, 128) and (dims[1] * dims[2]) % 128 == 0
    if slab:
        xin = _bf16_round(xin)
    mean = rstd = None
    if case["norm"]:
        mean = torch.randn(cin, generator=g) * 0.3
        rstd = torch.rand(cin, generator=g) + 0.5
        xin = (xin - mean.view(-1, 1, 1, 1)) * rstd.view(-1, 1, 1, 1)
    xin = F.relu(xin) if case["act"] == 1 else (F.leaky_relu(xin, 0.01) if case["act"] == 2 else xin)
    ref = F.conv3d(_bf16_round(xin)[None], _bf16_round(w), b, stride=stride, padding=1)[0]
    res = torch.randn(ref.shape, generator=g) if case["residual"] else None
    if res is not None:
        ref = ref + res
    y = op_conv3d_k3(x0.cuda(), w.cuda(), b.cuda(), x1.cuda() if c1 else None, stride,
                     (mean.cuda(), rstd.cuda()) if case["norm"] else None, case["act"],
                     res.cuda() if res is not None else None, impl=2)
    torch.cuda.synchronize()
    y = y.cpu()
    assert y.shape == ref.shape
    # the rolling stride-2 kernel stores its output as bf16 (B-format): one more rounding of 2^-9
    import os
    s2gen = os.environ.get("DCL_S2GEN", "1") != "0" and (c0, dims[0]) == (32, 64) and cout in (32, 64)
    bf16_out = stride == 2 and ((c0, cout, dims[0]) == (16, 32, 128) or s2gen)
    assert rel_err(y.numpy(), ref.numpy()) < (4e-3 if bf16_out else 2e-3)
    assert float((y - ref).abs().mean() / ref.abs().mean()) < (2e-3 if bf16_out else 1e-4)
    if bf16_out:
        assert float((y - _bf16_round(ref)).abs().mean() / ref.abs().mean()) < 2e-4     # identical up to rare ulp flips


@pytest.mark.parametrize("case", [
    dict(c0=16, c1=0, cout=16, dims=(128, 128, 128), stride=1, norm=True, act=1, residual=True),    # rolling kernel, split strips
    dict(c0=16, c1=0, cout=16, dims=(128, 128, 128), stride=1, norm=False, act=0, residual=False),
    dict(c0=32, c1=0, cout=32, dims=(64, 64, 64), stride=1, norm=True, act=2, residual=True),       # slab kernel at 64^3
    dict(c0=64, c1=0, cout=64, dims=(32, 32, 32), stride=1, norm=True, act=1, residual=True),       # slab
    dict(c0=32, c1=64, cout=96, dims=(32, 32, 32), stride=1, norm=False, act=0, residual=False),    # slab, two sources
    dict(c0=128, c1=0, cout=128, dims=(16, 16, 16), stride=1, norm=True, act=2, residual=False),    # slab, two passes
    dict(c0=256, c1=0, cout=384, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),   # slab, four passes
    dict(c0=128, c1=0, cout=256, dims=(16, 16, 16), stride=1, norm=False, act=0, residual=False),
    dict(c0=16, c1=0, cout=32, dims=(128, 128, 128), stride=2, norm=False, act=0, residual=False),  # rolling stride-2 kernel
    dict(c0=32, c1=0, cout=64, dims=(64, 64, 64), stride=2, norm=False, act=0, residual=False),     # im2col GEMM
    dict(c0=64, c1=0, cout=128, dims=(32, 32, 32), stride=2, norm=False, act=0, residual=False),    # im2col GEMM (EnDown3)
    dict(c0=5, c1=3, cout=19, dims=(9, 17, 18), stride=2, norm=True, act=2, residual=True),         # ragged, padded channels
    dict(c0=32, c1=0, cout=8, dims=(16, 16, 16), stride=1, norm=True, act=2, residual=False),       # auxiliary-head shapes
    dict(c0=8, c1=0, cout=2, dims=(20, 12, 36), stride=1, norm=False, act=0, residual=False),
])
def test_conv3d_k3_split_fp16(case):
    """DCL_F16X3 kernels (split operands, fp16 hi + fp16 lo: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on tcgen05, fp32 accumulate)
    against the PLAIN fp32 torch convolution - no operand rounding is granted: 22 significant bits per operand, B-format
    outputs stored as hi + lo."""
    y, ref = _conv_case(seed=31, impl=1, **case)
    assert y.shape == ref.shape
    e_max, e_mean = rel_err(y.numpy(), ref.numpy()), float((y - ref).abs().mean() / ref.abs().mean())
    print(f"split-fp16 conv {case['c0'] + case['c1']}->{case['cout']} {case['dims']} s{case['stride']}: max {e_max:.2e} mean {e_mean:.2e}")
    assert e_max < 2e-5
    assert e_mean < 1e-5      # (the fp32 CPU reference itself accumulates K = 432 .. 6912 terms in fp32)
