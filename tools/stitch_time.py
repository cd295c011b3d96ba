"""Isolated device timing of the overlap stitch (csrc/stitch.cu) on one 240x240x155 volume: the gather form
(gather_finalize_kernel) against the accumulate form (accumulate_vec_kernel x P + finalize_labels_kernel), in
algorithmic bytes against the measured copy bandwidth.  One JSON line per (stride, weights, form)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcl_b200 import _native as N  # noqa: E402

P3 = 128 ** 3


def main():
    import torch
    torch.cuda.init()
    lib = N.load_library()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6461.2
    shape = (C.c_int32 * 3)(240, 240, 155)
    V = 240 * 240 * 155
    for stride in (64, 32):
        for gaussian in (0, 1):
            for form in (0, 1):
                n = C.c_int32()
                us = lib.dcl_bench_stitch(shape, stride, gaussian, form, 20, C.byref(n))
                P = n.value
                # algorithmic bytes: every patch probability read once + one label byte per voxel; the
                # accumulate form additionally round-trips the fp32 accumulators
                algo = P * 4 * P3 * 4 + V
                moved = algo if form == 0 else 5 * V * 4 + P * P3 * (16 + 2 * 20) + V * 21
                print(json.dumps({"form": "gather" if form == 0 else "accumulate", "stride": stride, "patches": P,
                                  "weights": "gaussian" if gaussian else "uniform", "us_per_volume": round(us, 1),
                                  "algorithmic_bytes": algo, "algorithmic_GB/s": round(algo / us / 1e3, 1),
                                  "frac_of_copy_peak": round(algo / us / 1e3 / peak, 3),
                                  "bytes_moved_by_this_form": moved, "moved_GB/s": round(moved / us / 1e3, 1), "peak": peak}))


if __name__ == "__main__":
    main()
