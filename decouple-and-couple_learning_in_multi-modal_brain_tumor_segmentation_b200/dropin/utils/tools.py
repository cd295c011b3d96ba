"""Drop-in for the inference-path part of the reference ``utils/tools.py`` (:44-47, :89-109).
Host-side integer arithmetic on label maps; the device path computes the same counters in the
label kernel (csrc/stitch.cu) -- these functions exist so reference scripts keep importing."""
import numpy as np


def dice_score(o, t, eps=1e-8):
    num = 2 * (o * t).sum() + eps
    den = o.sum() + t.sum() + eps
    return num / den


def softmax_output_dice(output, target):
    output, target = np.asarray(output), np.asarray(target)
    regions = (
        (output > 0, target > 0),                                             # whole tumour
        ((output == 1) | (output == 3), (target == 1) | (target == 3)),       # tumour core
        (output == 3, target == 3),                                           # enhancing
    )
    return [dice_score(o, t) for o, t in regions]


def mIOU(o, t, eps=1e-8):
    """predict_simple.py:66-69"""
    num = (o * t).sum() + eps
    den = (o | t).sum() + eps
    return num / den


def softmax_output_mIou(output, target):
    """predict_simple.py:100-118"""
    output, target = np.asarray(output), np.asarray(target)
    return [mIOU(output > 0, target > 0),
            mIOU((output == 1) | (output == 3), (target == 1) | (target == 3)),
            mIOU(output == 3, target == 3)]


def cal_hausdorff(output, target):
    """predict_simple.py:121-144: HD95 of WT / TC / ET on the device (label maps may be numpy or CUDA tensors)."""
    import torch
    from dcl_b200 import volio
    lab = output if isinstance(output, torch.Tensor) else torch.from_numpy(np.asarray(output).astype(np.uint8))
    tgt = target if isinstance(target, torch.Tensor) else torch.from_numpy(np.asarray(target).astype(np.uint8))
    return volio.cal_hausdorff(lab.to("cuda", torch.uint8), tgt.to("cuda", torch.uint8))
