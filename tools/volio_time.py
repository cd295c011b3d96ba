"""Device timing of the kernels either side of the sliding window (csrc/volio.cu, csrc/hausdorff.cu) on one
240x240x155 case: CUDA events on the launching stream, inputs rotated through buffers larger than L2 where the
kernel is a pure byte stream.  Prints one JSON line per kernel (algorithmic bytes / time vs the measured copy peak)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dcl_b200 import volio as V  # noqa: E402


def peak_gbs():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_copy_gbs", "hbm_gbs", "copy_gbs"):
            if k in p:
                return float(p[k]), "measured copy"
        for v in p.values():
            if isinstance(v, dict):
                for k, x in v.items():
                    if "gb" in k.lower() and isinstance(x, (int, float)):
                        return float(x), "measured copy"
    except Exception:
        pass
    return 6461.2, "measured copy (DESIGN.md)"


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def main():
    from test_gpu_volio import blob_labels, random_labels, synthetic_mri
    shape = (240, 240, 155)
    V_ = shape[0] * shape[1] * shape[2]
    peak, src = peak_gbs()
    nbuf = 24                                   # 24 x 8.9 MB label maps > 126 MB L2
    labs = [torch.from_numpy(random_labels(shape, i % 3)).cuda() for i in range(nbuf)]
    tgts = [torch.from_numpy(blob_labels(shape, 40 + i)).cuda() for i in range(2)]
    rows = []

    def row(name, us, nbytes, note=""):
        rows.append({"kernel": name, "us": round(us, 2), "algorithmic_bytes": nbytes, "GB/s": round(nbytes / us / 1e3, 1),
                     "frac_of_copy_peak": round(nbytes / us / 1e3 / peak, 3), "peak": peak, "peak_source": src, "note": note})

    # the C entry points are called directly on preallocated outputs (the torch allocations of the Python wrappers
    # cost more host time than these kernels take on the device)
    import ctypes as C
    from dcl_b200 import _native as N
    lib = N.load_library()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sh = (C.c_int32 * 3)(*shape)
    P = lambda t: C.c_void_p(t.data_ptr())
    seg, turned = torch.empty_like(labs[0]), torch.empty((155, 240, 240), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
    row("export_labels_kernel", timed(lambda i: lib.dcl_export_labels(P(labs[i % nbuf]), sh, P(seg), P(turned), P(cnt), st)),
        V_ * 3, "1 B read + 2 x 1 B written per voxel")
    frames = torch.empty((155, 240, 240, 3), dtype=torch.uint8, device="cuda")
    pal = (C.c_uint8 * 12)(*V.PALETTE_PREDICT)
    row("snapshot_kernel", timed(lambda i: lib.dcl_snapshot_frames(P(labs[i % nbuf]), sh, pal, P(frames), st)), V_ * 4,
        "1 B read + 3 B written per voxel")
    sc = torch.zeros((155, 9), dtype=torch.int64, device="cuda")
    row("slice_counts_kernel", timed(lambda i: lib.dcl_slice_counts(P(labs[i % nbuf]), P(tgts[i % 2]), sh, P(sc), st)), V_ * 2,
        "2 B read per voxel")
    mri = [torch.from_numpy(np.ascontiguousarray(synthetic_mri(shape, 50 + i).transpose(3, 2, 1, 0))).cuda() for i in range(2)]
    vol = torch.empty((4, 240, 240, 160), dtype=torch.float32, device="cuda")
    stats = torch.zeros(17, dtype=torch.float64, device="cuda")
    row("mask_stats + normalise_reorder", timed(lambda i: lib.dcl_preprocess_volume(P(mri[i % 2]), sh, 160, P(vol), P(stats), st)),
        V_ * 16 * 2 + 240 * 240 * 160 * 16, "two 16 B/voxel reads + one 16 B/voxel write (2 x 143 MB inputs > L2)")
    for name, lab, tgt in (("blobs", tgts[0], tgts[1]), ("random vs blobs", labs[0], tgts[0])):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = V.hausdorff(lab, tgt)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        rows.append({"kernel": "dcl_hausdorff (3 regions x 4 kernels + histogram read-back)", "case": name, "ms_wall": round(ms, 3),
                     "hd95": out["hd95"], "surface_voxels": out["surface_voxels"]})
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
