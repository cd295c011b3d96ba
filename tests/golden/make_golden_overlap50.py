"""Golden fixtures for BASELINE.json configs[1] - the BENCHMARKED workload: one 4x240x240x155 volume, 128^3 patches at
50 % overlap (stride 64 -> 18 patches), uniform blend - from the UNMODIFIED reference model in /root/reference on CPU
(fp32).  Runs only in the build container; the produced ``overlap50_seed*.npz`` files are committed.

    python tests/golden/make_golden_overlap50.py            # 54 reference forwards, ~3 min on 8 cores

The reference has no overlap-weighted blend: the per-patch forward is the reference's
(``ClsWiseFormer.forward``, cls_wise_former.py:585-592, through ``model(x, None)[0]`` exactly as
predict_overlap.py:46-47 calls it) and the blend follows the reference's only accumulate-and-normalise path, the
sum-then-divide of predict_cls.py:184-203: acc += p, count += 1 over the covering patches in list order (fp32), then
acc / count; labels = argmax(0) as predict_overlap.py:141-143; Dice = utils.tools.softmax_output_dice (:153).
Volumes 0..2 (seed 1000 + i), dropout draws replayed from seed 2000 + i (Unet_skipconnection.py:31).
Stored per volume: the 18 keep-scale vectors, a digest of the blended probabilities, the label map packed at 2 bits per
voxel (volume 0: every voxel; volumes 1-2: every 4th voxel of the flattened map), histogram, Dice, and the quantiles
of the top-1 / top-2 probability margin that size the label-flip budget.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import N_SAMPLE, load_reference, put  # noqa: E402

PATCH = 128
SHAPE = (240, 240, 155)
STRIDE = 64


def axis_starts(n, stride):
    return sorted(set(range(0, n - PATCH, stride)) | {n - PATCH})


def pack2(labels_flat):
    """uint8 labels in {0..3} -> 4 per byte (little end first), zero padded."""
    a = np.asarray(labels_flat, dtype=np.uint8)
    pad = (-a.size) % 4
    a = np.concatenate([a, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (a[:, 0] | (a[:, 1] << 2) | (a[:, 2] << 4) | (a[:, 3] << 6)).astype(np.uint8)


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, _predict_overlap, tools = load_reference()
    torch.manual_seed(0)
    model = get_model("brats", True, "fixed", 0).eval()
    xs, ys, zs = (axis_starts(n, STRIDE) for n in SHAPE)
    starts = [(x, y, z) for z in zs for x in xs for y in ys]          # z-major: a contiguous chunk is a z-slab
    assert len(starts) == 18
    for i in range(3):
        torch.manual_seed(1000 + i)
        xv = torch.randn(1, 4, *SHAPE)
        torch.manual_seed(2000 + i)
        keeps = [torch.empty(1, 16, 1, 1, 1).bernoulli_(0.8).div_(0.8).reshape(16).numpy() for _ in starts]
        torch.manual_seed(2000 + i)                                     # the forwards below draw exactly these masks
        acc = np.zeros((4,) + SHAPE, dtype=np.float32)
        cnt = np.zeros(SHAPE, dtype=np.float32)
        with torch.no_grad():
            for (sx, sy, sz) in starts:
                p = model(xv[..., sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH], None)[0][0].numpy()
                acc[:, sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH] += p
                cnt[sx:sx + PATCH, sy:sy + PATCH, sz:sz + PATCH] += 1.0
        out = acc / cnt[None]
        labels = out.argmax(0)
        target = np.random.RandomState(i).randint(0, 4, SHAPE)
        dice = tools.softmax_output_dice(labels, target)
        g = {"starts": np.array(starts, dtype=np.int32), "keep_scale": np.stack(keeps)}
        put(g, "blend", torch.from_numpy(out)[None])
        flat = labels.astype(np.uint8).ravel()
        g["labels_sha256"] = np.frombuffer(hashlib.sha256(flat.tobytes()).digest(), dtype=np.uint8)
        g["labels_step"] = np.int64(1 if i == 0 else 4)
        g["labels_packed"] = pack2(flat[:: int(g["labels_step"])])
        g["labels_hist"] = np.bincount(flat, minlength=4).astype(np.int64)
        g["dice"] = np.array(dice, dtype=np.float64)
        srt = np.sort(out, axis=0)
        margin = srt[-1] - srt[-2]
        g["margin_quantiles"] = np.quantile(margin, [1e-5, 1e-4, 1e-3, 1e-2]).astype(np.float64)
        g["margin_below"] = np.array([(margin < t).mean() for t in (1e-6, 1e-5, 1e-4, 1e-3)], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"overlap50_seed{1000 + i}.npz"), **g)
        print(f"volume {i}: dice {dice} hist {g['labels_hist'].tolist()} margin<1e-5: {g['margin_below'][1]:.2e}", flush=True)
    _ = N_SAMPLE


if __name__ == "__main__":
    main()
