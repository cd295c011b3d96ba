"""Host-side model of the split-operand arithmetic of DCL_F16X3 (csrc/tc_common.cuh: split_x2; csrc/conv_tc.cu:
tc_pack_weights): v = hi + lo with hi = fp16(v), lo = fp16(v - hi); a product is a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
accumulated in fp32; weights are stored times 2^k so that their lo halves stay normal fp16 numbers.  numpy only - this
pins the error model DESIGN.md section 4 quotes (it is what decides whether the 13 discrete top-k selections of a patch
come out as the fp32 reference's)."""
import numpy as np


def split(v):
    v = np.clip(np.asarray(v, dtype=np.float32), -65504.0, 65504.0)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def pow2_scale(w):
    """The k of tc_pack_weights: max|w| * 2^k in [2^13, 2^14)."""
    _, e = np.frexp(np.abs(w).max())
    return int(np.clip(14 - e, -8, 30))


def split_dot(a, w, scale_weights):
    k = pow2_scale(w) if scale_weights else 0
    ah, al = split(a)
    wh, wl = split(w * np.float32(2.0 ** k))
    acc = (ah * wh + al * wh + ah * wl).sum(-1)        # the three products the GEMM / slab / stride-2 kernels issue
    return acc * 2.0 ** -k


def test_split_representation_has_22_bits_on_o1_values_and_saturates():
    rng = np.random.RandomState(0)
    v = (rng.randn(200000) * 3).astype(np.float32)
    hi, lo = split(v)
    err = np.abs(hi + lo - v.astype(np.float64))
    assert (err <= np.maximum(np.abs(v) * 2.0 ** -21, 6.0e-8)).all()       # 2^-22 relative, or fp16's subnormal step
    hi, lo = split(np.array([1e6, -3e5], np.float32))
    assert hi.tolist() == [65504.0, -65504.0] and np.isfinite(lo).all()


def test_weight_scale_keeps_the_lo_halves_normal():
    rng = np.random.RandomState(1)
    k_terms = 27 * 128                                     # a 128-channel 3x3x3 convolution
    a = rng.randn(64, k_terms).astype(np.float32)          # InstanceNorm'ed activations
    w = (rng.rand(k_terms).astype(np.float32) - 0.5) * 2 / np.sqrt(k_terms)      # kaiming-uniform-like, |w| <= 0.017
    exact = (a.astype(np.float64) * w.astype(np.float64)).sum(-1)
    scale = np.sqrt((exact ** 2).mean())
    err_plain = np.abs(split_dot(a, w, False) - exact).max() / scale
    err_scaled = np.abs(split_dot(a, w, True) - exact).max() / scale
    assert err_scaled < 4e-7                                # fp32-class: the fp32 accumulator itself is ~2^-24 * sqrt(K)
    assert err_plain > 3 * err_scaled                       # unscaled, w_lo ~ 1e-5 is subnormal: ~19 bits of w survive
    assert 13 <= np.log2(np.abs(w).max() * 2.0 ** pow2_scale(w)) < 14


def test_dropping_a_product_is_not_an_option():
    """a_hi*w_hi alone is plain fp16 (2^-11 per operand); without a_hi*w_lo the weights keep 11 bits."""
    rng = np.random.RandomState(2)
    a = rng.randn(32, 1728).astype(np.float32)
    w = ((rng.rand(1728) - 0.5) * 0.05).astype(np.float32)
    exact = (a.astype(np.float64) * w.astype(np.float64)).sum(-1)
    scale = np.sqrt((exact ** 2).mean())
    ah, al = split(a)
    wh, wl = split(w * 2.0 ** pow2_scale(w))
    one = (ah * wh).sum(-1) * 2.0 ** -pow2_scale(w)
    two = (ah * wh + al * wh).sum(-1) * 2.0 ** -pow2_scale(w)
    three = split_dot(a, w, True)
    e1, e2, e3 = (np.abs(x - exact).max() / scale for x in (one, two, three))
    assert e1 > 1e-4 and e2 > 5e-5 and e3 < 4e-7
