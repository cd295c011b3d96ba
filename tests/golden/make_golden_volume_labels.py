"""Full label map of the reference-tiling golden (tests/golden/volume_seed1000.npz holds only its digest): runs the
UNMODIFIED reference `predict_overlap.tailor_and_concat` (predict_overlap.py:31-58) + `argmax(0)` (:141-143) on CPU with the
same seeds as make_golden.py and stores the 240x240x155 labels packed at 2 bits per voxel, so that the GPU tests can count
label flips over the WHOLE volume against the <= 1e-4 budget.

    python tests/golden/make_golden_volume_labels.py          # 8 reference forwards, ~30 s
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402
from make_golden_overlap50 import pack2  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, predict_overlap, _tools = load_reference()
    torch.manual_seed(0)
    model = get_model("brats", True, "fixed", 0).eval()
    torch.manual_seed(1000)
    xv = torch.randn(1, 4, 240, 240, 155)
    torch.manual_seed(2000)
    with torch.no_grad():
        out = predict_overlap.tailor_and_concat(xv, None, model)
    labels = out[0, :, :240, :240, :160].numpy().argmax(0).astype(np.uint8)
    g = np.load(os.path.join(HERE, "volume_seed1000.npz"))
    assert hashlib.sha256(labels.tobytes()).digest() == g["labels_sha256"].tobytes(), "not the label map of volume_seed1000.npz"
    np.savez_compressed(os.path.join(HERE, "volume_seed1000_labels.npz"), labels_packed=pack2(labels.ravel()),
                        shape=np.array(labels.shape, dtype=np.int64))
    print("written; histogram", np.bincount(labels.ravel(), minlength=4).tolist())


if __name__ == "__main__":
    main()
