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
