"""One launch of every kernel either side of the forward (stitch, export, input side, surface distances) on a
240x240x155 case -- the command the ncu captures under profiles/ profile (tools/volio_time.py and
tools/stitch_time.py take the timings; numbers printed under a profiler are not measurements)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dcl_b200 import _native as N  # noqa: E402
from dcl_b200 import volio as V  # noqa: E402


def main():
    from test_gpu_volio import blob_labels, random_labels, synthetic_mri
    shape = (240, 240, 155)
    lib = N.load_library()
    n = C.c_int32()
    sh = (C.c_int32 * 3)(*shape)
    for form in (0, 1):
        us = lib.dcl_bench_stitch(sh, 64, 0, form, 1, C.byref(n))
        print("stitch form", form, "patches", n.value, "us", us)
    lab, tgt = torch.from_numpy(blob_labels(shape, 40)).cuda(), torch.from_numpy(blob_labels(shape, 41)).cuda()
    rnd = torch.from_numpy(random_labels(shape, 0)).cuda()
    print(V.hausdorff(lab, tgt))
    V.export_labels(rnd)
    V.snapshot_frames(rnd)
    V.slice_counts(rnd, tgt)
    mri = torch.from_numpy(np.ascontiguousarray(synthetic_mri(shape, 50).transpose(3, 2, 1, 0))).cuda()
    V.preprocess_volume(mri, 160)
    V.reorder_labels(torch.from_numpy(np.ascontiguousarray(blob_labels(shape, 42).transpose(2, 1, 0))).cuda(), 160)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
