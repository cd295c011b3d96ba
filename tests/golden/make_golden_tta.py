"""Golden fixture for the 8-flip test-time augmentation around the tiling.

Runs the UNMODIFIED reference model and `tailor_and_concat` from /root/reference on CPU and composes them exactly as
predict_cls.py:180-203 does (that block is inline in the reference's validate_softmax, so it is restated here line by
line; everything it calls is the reference's own code).  64 forwards, ~4 min on 8 cores.  Build container only.

    python tests/golden/make_golden_tta.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import N_SAMPLE, load_reference, put  # noqa: E402

FLIPS = [(), (2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)]     # predict_cls.py:182-201, in order


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, predict_overlap, tools = load_reference()
    tailor_and_concat = predict_overlap.tailor_and_concat
    torch.manual_seed(0)
    model = get_model("brats", True, "fixed", 0).eval()
    i = 0
    torch.manual_seed(1000 + i)
    x = torch.randn(1, 4, 240, 240, 155)
    missing_modal = None
    torch.manual_seed(3000 + i)
    keeps = [torch.empty(1, 16, 1, 1, 1).bernoulli_(0.8).div_(0.8).reshape(16).numpy() for _ in range(64)]
    torch.manual_seed(3000 + i)
    with torch.no_grad():
        x = x[..., :155]                                                                               # :180
        logit = F.softmax(tailor_and_concat(x, missing_modal, model, None), 1)                         # :182
        for dims in FLIPS[1:]:                                                                         # :184-201
            logit += F.softmax(tailor_and_concat(x.flip(dims=dims), missing_modal, model).flip(dims=dims), 1)
        output = logit / 8.0                                                                           # :203
    out = {}
    put(out, "tta", output)
    o = output[0, :, :240, :240, :155].numpy()
    labels = o.argmax(0)
    target = np.random.RandomState(i).randint(0, 4, (240, 240, 155))
    out["keep_scale"] = np.stack(keeps)
    out["labels_hist"] = np.array([np.sum(labels == k) for k in range(4)], dtype=np.int64)
    out["labels_sample"] = labels.ravel()[:: labels.size // N_SAMPLE][:N_SAMPLE].astype(np.uint8)
    out["dice"] = np.array(tools.softmax_output_dice(labels, target), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "volume_tta_seed1000.npz"), **out)
    print("tta golden written; dice", out["dice"], "hist", out["labels_hist"])


if __name__ == "__main__":
    main()
