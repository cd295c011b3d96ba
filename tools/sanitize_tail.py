"""Small ragged cases through every kernel either side of the forward -- the command run under
`compute-sanitizer --tool memcheck` (sanitizers).  Shapes are chosen so that every tile has a ragged
edge: (37, 41, 29), (33, 1, 65), and a (130, 129, 131) volume for both stitch forms."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_b200 import _native as N  # noqa: E402
from dcl_b200 import volio as V  # noqa: E402


def main():
    lib = N.load_library()
    rng = np.random.RandomState(0)
    for shape in ((37, 41, 29), (33, 1, 65), (1, 1, 1)):
        lab = torch.from_numpy(rng.randint(0, 4, shape).astype(np.uint8)).cuda()
        tgt = torch.from_numpy(rng.randint(0, 4, shape).astype(np.uint8)).cuda()
        V.export_labels(lab)
        V.snapshot_frames(lab)
        V.slice_counts(lab, tgt)
        print(shape, V.hausdorff(lab, tgt)["hd95"])
        mri = torch.from_numpy(rng.rand(4, shape[2], shape[1], shape[0]).astype(np.float32)).cuda()
        V.preprocess_volume(mri, shape[2] + 5)
        V.reorder_labels(lab.permute(2, 1, 0).contiguous(), shape[2] + 3, True)
    sh = (C.c_int32 * 3)(130, 129, 131)
    n = C.c_int32()
    for form in (0, 1):
        for gaussian in (0, 1):
            print("stitch form", form, "gaussian", gaussian, lib.dcl_bench_stitch(sh, 64, gaussian, form, 1, C.byref(n)), n.value)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
