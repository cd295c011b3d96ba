"""CPU model of the staging layout of the rolling stride-2 kernel (csrc/conv_s2.cu, S2Cfg): the producers
de-interleave an input plane into four (w parity) x (row parity) blocks so that every tap of an M = 128 tile is one
contiguous run of 128 positions.  This test re-derives the constants of S2Cfg in Python and checks, for the three
instantiated geometries, that position `run_start + m` of tap (kh, kw) is exactly input (2*oh + kh - 1, 2*ow + kw - 1)
for every unmasked output lane m, and that the masked lanes are the zero-padding ones."""
import numpy as np
import pytest


def cfg(ci, co, gi, nt, nslot):
    go = gi // 2
    rpt = 128 // go
    th = rpt * nt
    n_odd, n_even = th + 1, th
    b_oo = 8
    b_oe = b_oo + n_odd * go
    b_eo = b_oe + n_even * go
    b_ee = b_eo + n_odd * go
    npos = b_ee + n_even * go
    return dict(ci=ci, co=co, gi=gi, go=go, rpt=rpt, th=th, nt=nt, nslot=nslot, rows=2 * th + 1, b_oo=b_oo, b_oe=b_oe,
                b_eo=b_eo, b_ee=b_ee, npos=npos, slot_bytes=ci // 8 * npos * 16, w_bytes=27 * ci * co * 2)


CONFIGS = [cfg(16, 32, 128, 2, 5), cfg(32, 64, 64, 1, 3), cfg(32, 32, 64, 1, 4)]


@pytest.mark.parametrize("c", CONFIGS, ids=lambda c: f"{c['ci']}to{c['co']}at{c['gi']}")
def test_staged_runs_address_the_right_input_voxels(c):
    gi, go, th, rpt = c["gi"], c["go"], c["th"], c["rpt"]
    for oh0 in (0, th, go - th):                                  # first, an inner and the last row tile
        # producer (one channel chunk): staged[pos] = (input row, input column), -1 = zero padding / pad positions
        staged = np.full((c["npos"], 2), -9, dtype=np.int64)
        staged[:8] = -1
        for r in range(c["rows"]):
            h_in = 2 * oh0 - 1 + r
            for w in range(gi):
                base = ((c["b_oe"] if r & 1 else c["b_oo"]) if w & 1 else (c["b_ee"] if r & 1 else c["b_eo"]))
                staged[base + (r >> 1) * go + (w >> 1)] = (h_in, w) if h_in >= 0 else (-1, -1)
        assert (staged[8:, 0] != -9).all()                        # every position of the four blocks is written
        for t in range(c["nt"]):
            for kh in range(3):
                ridx = rpt * t + (1 if kh == 2 else 0)
                for kw in range(3):
                    base = ((c["b_ee"] if kh == 1 else c["b_eo"]) if kw == 1 else (c["b_oe"] if kh == 1 else c["b_oo"]))
                    start = base + ridx * go - (1 if kw == 0 else 0)
                    for m in range(128):
                        oh, ow = oh0 + rpt * t + m // go, m % go
                        want = (2 * oh + kh - 1, 2 * ow + kw - 1)
                        masked = kw == 0 and m % go == 0            # the disable-output-lane mask of the kw = 0 taps
                        if masked:
                            assert want[1] == -1                    # exactly the left zero-padding column
                            continue
                        got = tuple(staged[start + m])
                        if want[0] < 0:
                            assert got == (-1, -1)                  # top zero-padding row was staged as zeros
                        else:
                            assert got == want, (oh0, t, kh, kw, m)
                            assert 0 <= want[0] < gi and 0 <= want[1] < gi


@pytest.mark.parametrize("c", CONFIGS, ids=lambda c: f"{c['ci']}to{c['co']}at{c['gi']}")
def test_shared_memory_budget_and_alignment(c):
    total = c["nslot"] * c["slot_bytes"] + c["w_bytes"] + c["co"] * 4 + 4 * c["co"] * 2 * 4 + (2 * c["nslot"] + 4) * 8 + 16
    assert total <= 232448                                          # 227 KB opt-in limit per CTA
    assert c["slot_bytes"] % 128 == 0 and (c["nslot"] * c["slot_bytes"]) % 128 == 0
    for b in ("b_oo", "b_oe", "b_eo", "b_ee"):
        assert c[b] % 8 == 0                                        # blocks start on a 128-byte (8-position) boundary
    assert 2 * c["nt"] * c["co"] <= 512                             # two TMEM accumulator buffers
    assert (total + 16) >> 4 < 1 << 14                              # every operand address fits the descriptor field
