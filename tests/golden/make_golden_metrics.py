"""Golden values for the metric row from the UNMODIFIED reference glue:
``predict_simple.cal_hausdorff`` (predict_simple.py:121-144) -> ``utils/hausdorff.py:86-123``,
``predict_simple.softmax_output_mIou`` (:100-118) and ``utils.tools.softmax_output_dice`` (utils/tools.py:89-109), run
on seeded label maps in the build container.

    python tests/golden/make_golden_metrics.py

The reference hands the distance computation itself to ``medpy.metric.hd95`` / ``hd``; medpy is absent from this image
(and from /opt/wheelhouse), so ``medpy.metric`` is substituted by the scipy restatement of medpy 0.4.0 in
``oracle/volio_oracle.py`` -- everything around it (region definitions, the bool ``+`` of predict_simple.py:131-132, the
empty / full guards that return 0) is the reference's own code.  The inputs are stored with the values, so the tests
need neither the reference nor this script.
"""
import json
import os
import sys
import types

import numpy as np
from scipy.ndimage import gaussian_filter

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import volio_oracle as O  # noqa: E402


def load_reference():
    for name in ("nibabel", "imageio", "medpy", "medpy.metric", "setproctitle", "SimpleITK"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    metric = sys.modules["medpy.metric"]
    metric.hd95 = lambda result, reference, voxelspacing=None, connectivity=1: O.medpy_hd95(result, reference)
    metric.hd = lambda result, reference, voxelspacing=None, connectivity=1: O.medpy_hd(result, reference)
    sys.modules["medpy"].metric = metric
    sys.path.insert(0, REF)
    import predict_simple
    import utils.hausdorff as hausdorff
    import utils.tools as tools
    return predict_simple, hausdorff, tools


def blobs(shape, seed, shift):
    rng = np.random.RandomState(seed)
    f = gaussian_filter(rng.randn(*shape), sigma=min(shape) / 9.0)
    f = (f - f.mean()) / f.std() + shift
    lab = np.zeros(shape, np.uint8)
    lab[f > 1.0] = 2
    lab[f > 1.5] = 1
    lab[f > 2.0] = 3
    return lab


def main():
    predict_simple, hausdorff, tools = load_reference()
    shape = (40, 36, 28)
    cases = {}
    a, b = blobs(shape, 1, 0.8), blobs(shape, 2, 0.8)
    cases["blobs"] = (a, b)
    cases["shifted"] = (np.roll(b, (2, -3, 1), axis=(0, 1, 2)), b)
    rng = np.random.RandomState(3)
    cases["dense_random"] = (rng.randint(0, 4, shape).astype(np.uint8), rng.randint(0, 4, shape).astype(np.uint8))
    cases["empty_prediction"] = (np.zeros(shape, np.uint8), b)
    cases["full_prediction"] = (np.full(shape, 3, np.uint8), b)
    p, q = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
    p[3, 4, 5] = 3
    q[30, 33, 27] = 3
    cases["single_voxels"] = (p, q)
    p, q = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
    p[5:15, 5:15, 5:15] = 2
    q[8:20, 5:15, 0:10] = 2
    cases["wt_only"] = (p, q)
    out, arrays = {}, {}
    for name, (o, t) in cases.items():
        arrays[name + "/output"], arrays[name + "/target"] = o, t
        out[name] = {
            "cal_hausdorff": [float(v) for v in predict_simple.cal_hausdorff(o, t)],
            "hausdorff_distance": [float(hausdorff.hausdorff_distance(x, y)) for x, y in zip(O.regions(o), O.regions(t))],
            "softmax_output_mIou": [float(v) for v in predict_simple.softmax_output_mIou(o, t)],
            "softmax_output_dice": [float(v) for v in tools.softmax_output_dice(o, t)],
        }
        print(name, out[name])
    np.savez_compressed(os.path.join(HERE, "metrics_cases.npz"), **arrays)
    with open(os.path.join(HERE, "metrics_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
