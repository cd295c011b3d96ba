"""CPU model of the kh pairing of the 32-channel rolling kernel (csrc/conv_tc.cu, RollCfg::KH2, weight layout 4).

A plane of the 64^3 layers is staged as ROWS = TH + 2 rows of W = 64 voxels; an M = 128 tile is two rows.  The A tile that
starts on staged row s covers input rows (s, s+1): for EVEN s = 2t it is the kh = 0 operand of output tile t and the kh = 2
operand of tile t-1, so ONE MMA against the stacked [W_kh2 | W_kh0] matrix feeds both (adjacent accumulator blocks); for
ODD s = 2t+1 it is the kh = 1 operand of tile t.  This test re-states (a) the packed weight layout of tc_pack_weights and
(b) the MMA schedule of the kernel in numpy - matrices addressed exactly as the UMMA descriptors address them (start row,
rows per K chunk = leading byte offset / 16) - and checks that the accumulators equal the direct convolution, for the bf16
geometry (32 input channels per launch, one image) and the split-fp16 one (16 per launch, [hi | lo] rows, two launches)."""
import numpy as np
import pytest

W, TH, NT, ROWS = 64, 8, 4, 10


def pack_layout4(w, x3):
    """w[cout][cin][kd][kh][kw] -> 16-byte rows (8 input channels each), as tc_pack_weights lays them out.
    Returns rows[n_rows][8] and, for the split mode, the hi / lo parts stored as separate rows."""
    cout, cin = w.shape[:2]
    cl = 16 if x3 else 32                    # input channels per launch
    kcl = cl // 8
    cn = (2 if x3 else 1) * cout             # rows of one matrix
    rows = np.zeros(((cin // cl) * 9 * kcl * 3 * cn, 8))
    parts = (0.75, 0.25) if x3 else (1.0,)   # stand-ins for hi and lo: any split with hi + lo = w
    for ci in range(cin):
        half, cc = divmod(ci, cl)
        chunk = cc // 8
        for kd in range(3):
            for kh in range(3):
                for kw in range(3):
                    q_rows = (half * 9 + kd * 3 + kw) * (kcl * 3 * cn)
                    row = q_rows + kcl * 2 * cn + chunk * cn if kh == 1 else q_rows + chunk * 2 * cn + (0 if kh == 2 else cn)
                    for p, frac in enumerate(parts):
                        rows[row + p * cout: row + (p + 1) * cout, cc % 8] = frac * w[:, ci, kd, kh, kw]
    return rows


def b_operand(rows, start_row, lbo_rows, n, k_chunks):
    """The N x K matrix an UMMA descriptor (start, LBO) with instruction N describes: chunk c of K = rows
    [start + c * lbo, start + c * lbo + n)."""
    return np.concatenate([rows[start_row + c * lbo_rows: start_row + c * lbo_rows + n] for c in range(k_chunks)], axis=1)


@pytest.mark.parametrize("x3", [False, True], ids=["bf16", "split-fp16"])
def test_schedule_and_layout_reproduce_the_convolution(x3):
    rng = np.random.default_rng(3)
    cout, cin = 32, 32
    cl = 16 if x3 else 32
    kc = cl // 8                              # chunks per launch
    cn = (2 if x3 else 1) * cout
    w = rng.standard_normal((cout, cin, 3, 3, 3))
    x = rng.standard_normal((3, ROWS, W, cin))            # three staged planes (kd), rows h0-1 .. h0+TH, no halo columns
    rows = pack_layout4(w, x3)
    q_rows = kc * 3 * cn                                  # QB / 16
    p_lbo, s_lbo = 2 * cn, cn                             # in 16-byte rows
    acc = np.zeros((NT, 128, cn))                         # accumulator blocks of one plane buffer: tile t at columns t*CN
    n_mma = 0
    for launch in range(cin // cl):
        w_base = launch * 9 * q_rows
        xa = x[..., launch * cl: (launch + 1) * cl].reshape(3, ROWS * W, cl)
        for kd in range(3):
            for kw in (1, 0, 2):
                bp = w_base + (kd * 3 + kw) * q_rows       # pair images [kh2 | kh0]
                bs = bp + kc * p_lbo                       # kh = 1 images behind them

                def a_tile(s):                             # 128 positions from staged row s, shifted by kw - 1, masked lanes zero
                    pos = s * W + (kw - 1) + np.arange(128)
                    a = xa[kd][np.clip(pos, 0, ROWS * W - 1)].copy()
                    wcol = np.arange(128) % W
                    a[(wcol == 0) & (kw == 0)] = 0.0       # disable-output-lane masks = the zero padding along w
                    a[(wcol == W - 1) & (kw == 2)] = 0.0
                    return a
                for t in range(NT):                        # odd rows: kh = 1 -> tile t, N = CN
                    acc[t] += a_tile(2 * t + 1) @ b_operand(rows, bs, s_lbo, cn, kc).T
                    n_mma += 1
                for t in range(NT + 1):                    # even rows: kh = 2 -> tile t-1, kh = 0 -> tile t
                    a = a_tile(2 * t)
                    if t == 0:
                        acc[0] += a @ b_operand(rows, bp + cn, p_lbo, cn, kc).T
                    elif t == NT:
                        acc[NT - 1] += a @ b_operand(rows, bp, p_lbo, cn, kc).T
                    else:
                        d = a @ b_operand(rows, bp, p_lbo, 2 * cn, kc).T        # one MMA, two adjacent accumulator blocks
                        acc[t - 1] += d[:, :cn]
                        acc[t] += d[:, cn:]
                    n_mma += 1
    assert n_mma == (cin // cl) * 9 * (2 * NT + 1)         # 9 instead of 12 MMAs per (kd, kw) and operand half
    out = acc[..., :cout] + acc[..., cout:] if x3 else acc  # split mode: D1 + D2 (epilogue)
    # direct convolution of the TH output rows (output row r <-> staged rows r .. r+2), zero padding along w
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (0, 0)))
    want = np.zeros((TH, W, cout))
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                want += np.einsum("hwc,oc->hwo", xp[kd, kh: kh + TH, kw: kw + W], w[:, :, kd, kh, kw])
    np.testing.assert_allclose(out.reshape(TH, W, cout), want, rtol=1e-10, atol=1e-10)


def test_operand_read_model_of_the_pairing():
    """clk per MMA = max(40, (4096 B of A + N * 32 B of B) / 128) up to N = 128: per (kd, kw) and operand
    half the plane costs 480 clk paired against 576 clk unpaired in split-fp16 (CN = 64), 384 against 480 in bf16 (CN = 32)."""
    def clk(n):
        return max(40.0, (4096 + n * 32) / 128)
    for cn, paired, plain in ((64, 480, 576), (32, 384, 480)):
        assert NT * clk(cn) + 2 * clk(cn) + (NT - 1) * clk(2 * cn) == paired
        assert 3 * NT * clk(cn) == plain
