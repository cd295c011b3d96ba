"""The 8-flip TTA algebra of the oracle against the reference expression (predict_cls.py:180-203), CPU only.
The reference block is inline in its validate_softmax; it is restated here verbatim on top of a stand-in
``tailor_and_concat`` (any volume -> volume map that does not commute with flips), which is all the block depends on."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import stitch_oracle as S


def _fake_tailor_and_concat(x, missing_modal=None, model=None, target=None):
    g = torch.Generator().manual_seed(5)
    w = torch.randn(x.shape[1:], generator=g)
    return torch.tanh(x * w + 0.3 * x.roll(1, 2) - 0.2 * x.roll(2, 4))          # position dependent: flips matter


def _reference_block(x, missing_modal, model, tailor_and_concat):
    x = x[..., :155]                                                                                    # :180
    logit = F.softmax(tailor_and_concat(x, missing_modal, model, None), 1)                               # :182
    logit += F.softmax(tailor_and_concat(x.flip(dims=(2,)), missing_modal, model).flip(dims=(2,)), 1)     # flip H
    logit += F.softmax(tailor_and_concat(x.flip(dims=(3,)), missing_modal, model).flip(dims=(3,)), 1)     # flip W
    logit += F.softmax(tailor_and_concat(x.flip(dims=(4,)), missing_modal, model).flip(dims=(4,)), 1)     # flip D
    logit += F.softmax(tailor_and_concat(x.flip(dims=(2, 3)), missing_modal, model).flip(dims=(2, 3)), 1)
    logit += F.softmax(tailor_and_concat(x.flip(dims=(2, 4)), missing_modal, model).flip(dims=(2, 4)), 1)
    logit += F.softmax(tailor_and_concat(x.flip(dims=(3, 4)), missing_modal, model).flip(dims=(3, 4)), 1)
    logit += F.softmax(tailor_and_concat(x.flip(dims=(2, 3, 4)), missing_modal, model).flip(dims=(2, 3, 4)), 1)
    return logit / 8.0                                                                                  # :203


def test_tta_average_matches_reference_expression():
    torch.manual_seed(11)
    x = torch.randn(1, 4, 12, 10, 160)
    want = _reference_block(x, None, None, _fake_tailor_and_concat)[0].numpy()
    xs = x[..., :155]
    stitched = [_fake_tailor_and_concat(xs.flip(dims=d) if d else xs)[0].numpy() for d in S.TTA_FLIPS]
    got = S.tta_average_from_stitched(stitched)
    assert got.shape == want.shape == (4, 12, 10, 155)
    assert np.abs(got - want).max() < 5e-7
    assert np.abs(got.sum(0) - 1).max() < 1e-6
    assert S.TTA_FLIPS == [(), (2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]
