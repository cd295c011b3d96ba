"""ORACLE (test infrastructure, not product code).

CPU restatement of the sliding-window driver around the model: patch tiling, stitching,
arg-max labelling and WT/TC/ET Dice.  See ``clswiseformer_oracle.py`` for who may import
this and how it is pinned to the reference (``tests/golden/make_golden.py``).

Reference mode follows ``predict_overlap.py:31-58`` literally, *including* its 5-voxel
z shift (the z-tail copies patch-local 96:123, i.e. global z 123:150, into z 128:155).
The weighted modes (``uniform`` / ``gaussian``) are the extension BASELINE.json configs
2/4/5 ask for; the reference has no code for them, their convention (sum then divide)
follows the TTA average at ``predict_cls.py:184-203``.
"""
from __future__ import annotations

import numpy as np
import torch

PATCH = 128


def axis_starts(length, stride, patch=PATCH):
    """sorted(set(range(0, L-128, stride)) | {L-128})."""
    return sorted(set(range(0, length - patch, stride)) | {length - patch})


def patch_starts(shape, stride):
    """All patch origins, z-major then x then y ... order = (z, x, y) nested so that a
    contiguous chunk of the list is a z-slab."""
    xs, ys, zs = (axis_starts(n, stride) for n in shape)
    return [(x, y, z) for z in zs for x in xs for y in ys]


# The 8 fixed corners of predict_overlap.py:34-41, in the reference's order.
REFERENCE_STARTS = [(x, y, z) for z in (0, 27) for x in (0, 112) for y in (0, 112)]


def reference_plan():
    """Crop table equivalent to predict_overlap.py:49-56: for every patch the destination
    box (global) and the source offset (patch-local) of the block that is copied."""
    plan = []
    for (x0, y0, z0) in REFERENCE_STARTS:
        dx = (0, 128, 0) if x0 == 0 else (128, 240, 16)      # dst lo, dst hi, src lo
        dy = (0, 128, 0) if y0 == 0 else (128, 240, 16)
        dz = (0, 128, 0) if z0 == 0 else (128, 155, 96)      # 96, not 101: the reference's shift
        plan.append(((x0, y0, z0), (dx[0], dy[0], dz[0]), (dx[1], dy[1], dz[1]), (dx[2], dy[2], dz[2])))
    return plan


@torch.no_grad()
def tailor_and_concat(x, missing_modal, model, target=None):
    """predict_overlap.py:31-58.  ``model(patch, missing_modal)[0]`` must return 4 channels
    (the reference writes the probabilities into a clone of the 4-modality input)."""
    y = x.clone()
    for (start, lo, hi, src) in reference_plan():
        sx, sy, sz = start
        p = model(x[..., sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH], missing_modal)[0]
        y[..., lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = p[
            ..., src[0]:src[0] + hi[0] - lo[0], src[1]:src[1] + hi[1] - lo[1], src[2]:src[2] + hi[2] - lo[2]]
    return y[..., :155]


def stitch_reference_from_probs(probs_list, shape=(240, 240, 155)):
    """Same crop-overwrite applied to already computed per-patch probabilities
    (list of (C,128,128,128) arrays in REFERENCE_STARTS order)."""
    c = probs_list[0].shape[0]
    out = np.zeros((c,) + tuple(shape), dtype=probs_list[0].dtype)
    for p, (start, lo, hi, src) in zip(probs_list, reference_plan()):
        out[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = p[
            :, src[0]:src[0] + hi[0] - lo[0], src[1]:src[1] + hi[1] - lo[1], src[2]:src[2] + hi[2] - lo[2]]
    return out


def patch_weight(mode, patch=PATCH):
    """Per-voxel blending weight of one patch (separable).  uniform: 1.  gaussian:
    exp(-0.5*((i-c)/(patch/8))^2) per axis, c=(patch-1)/2, in fp32."""
    if mode == "uniform":
        w1 = np.ones(patch, dtype=np.float32)
    elif mode == "gaussian":
        i = np.arange(patch, dtype=np.float32)
        w1 = np.exp(-0.5 * ((i - np.float32((patch - 1) / 2)) / np.float32(patch / 8)) ** 2).astype(np.float32)
    else:
        raise ValueError(mode)
    return w1


def accumulate_from_probs(probs_list, starts, mode="uniform", shape=(240, 240, 155)):
    """Extension mode: acc += w*p, wsum += w over all covering patches, out = acc / wsum.
    Patches are visited in list order and summed in fp32 in that order."""
    c = probs_list[0].shape[0]
    acc = np.zeros((c,) + tuple(shape), dtype=np.float32)
    wsum = np.zeros(tuple(shape), dtype=np.float32)
    w1 = patch_weight(mode)
    w3 = (w1[:, None, None] * w1[None, :, None]) * w1[None, None, :]
    for p, (sx, sy, sz) in zip(probs_list, starts):
        acc[:, sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH] += w3[None] * p.astype(np.float32)
        wsum[sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH] += w3
    return acc / wsum[None]


# ---- 8-flip test-time augmentation around the tiling (predict_cls.py:180-203) ------------
# flipped dims of the (N,C,X,Y,Z) tensor in the reference's order: none, H(2), W(3), D(4), HW, HD, WD, HWD
TTA_FLIPS = [(), (2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]


def softmax4(a):
    """F.softmax(., 1) on a (C, ...) fp32 array: exp(x - max) / sum, all in fp32."""
    a = np.asarray(a, dtype=np.float32)
    e = np.exp(a - a.max(axis=0, keepdims=True), dtype=np.float32)
    return e / e.sum(axis=0, keepdims=True, dtype=np.float32)


def tta_average_from_stitched(stitched):
    """predict_cls.py:182-203 given the 8 stitched outputs T(flip_f(x)) (each (C,X,Y,155), flip order TTA_FLIPS):
    logit = softmax(T(x));  logit += softmax(T(flip_f(x)).flip(f)) for the 7 flips;  output = logit / 8.
    (The softmax is applied to what already are probabilities - as the reference does.)"""
    logit = softmax4(stitched[0])
    for y, dims in zip(stitched[1:], TTA_FLIPS[1:]):
        logit = logit + softmax4(np.flip(y, axis=tuple(d - 1 for d in dims)))
    return logit / np.float32(8.0)


@torch.no_grad()
def tta_tailor_and_concat(x, missing_modal, model):
    """predict_cls.py:180-203 as a function: x (1,4,240,240,>=155) -> (1,4,240,240,155) averaged probabilities."""
    import torch.nn.functional as F
    x = x[..., :155]
    logit = F.softmax(tailor_and_concat(x, missing_modal, model), 1)
    for dims in TTA_FLIPS[1:]:
        logit += F.softmax(tailor_and_concat(x.flip(dims=dims), missing_modal, model).flip(dims=dims), 1)
    return logit / 8.0


def labels_from_probs(probs):
    """predict_overlap.py:141-143: numpy argmax over the class axis (first maximum wins)."""
    return np.asarray(probs).argmax(0)


def label_histogram(labels):
    """predict_overlap.py:144-148."""
    return [int(np.sum(labels == k)) for k in range(4)]


def dice_score(o, t, eps=1e-8):
    """utils/tools.py:44-47."""
    num = 2 * (o * t).sum() + eps
    den = o.sum() + t.sum() + eps
    return num / den


def softmax_output_dice(output, target):
    """utils/tools.py:89-109: WT = >0, TC = {1,3}, ET = 3 (target label 4 already mapped to 3,
    predict_overlap.py:152)."""
    return [
        dice_score(output > 0, target > 0),
        dice_score((output == 1) | (output == 3), (target == 1) | (target == 3)),
        dice_score(output == 3, target == 3),
    ]


def region_counts(output, target):
    """The 9 integers the Dice needs: per region (|o|, |t|, |o&t|)."""
    regs = [
        (output > 0, target > 0),
        ((output == 1) | (output == 3), (target == 1) | (target == 3)),
        (output == 3, target == 3),
    ]
    return [[int(o.sum()), int(t.sum()), int((o & t).sum())] for o, t in regs]


def tta_flips():
    """predict_cls.py:184-203: identity + the 7 axis-flip subsets of (2,3,4), averaged /8."""
    return [(), (2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]
