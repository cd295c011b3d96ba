"""GPU parity of the kernels either side of the sliding window against the numpy / scipy
oracle: integer and byte work is bit-exact, the z-score is within 1e-5 relative (fp32 mean / std estimates)."""
import os

import numpy as np
import pytest
import torch
from scipy.ndimage import gaussian_filter

from dcl_b200 import volio as V
from oracle import volio_oracle as O

pytestmark = pytest.mark.gpu

FULL = (240, 240, 155)
RAGGED = (37, 41, 29)


def random_labels(shape, seed):
    return np.random.RandomState(seed).randint(0, 4, shape).astype(np.uint8)


def blob_labels(shape, seed, shift=0.0):
    """nested smooth regions like a tumour: WT (label 2) contains TC (label 1) contains ET (label 3)"""
    rng = np.random.RandomState(seed)
    f = gaussian_filter(rng.randn(*shape), sigma=min(shape) / 14.0)
    f = (f - f.mean()) / f.std() + shift
    lab = np.zeros(shape, np.uint8)
    lab[f > 1.0] = 2
    lab[f > 1.5] = 1
    lab[f > 2.0] = 3
    return lab


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [FULL, RAGGED, (1, 1, 1), (33, 1, 65)])
def test_export_labels_bit_exact(shape):
    lab = random_labels(shape, 10)
    out = V.export_labels(dev(lab))
    seg = O.export_seg(lab)
    assert np.array_equal(out["seg"].cpu().numpy(), seg)
    assert np.array_equal(out["seg_nifti"].cpu().numpy(), seg.transpose(2, 1, 0))       # x fastest = seg.tobytes('F')
    assert out["seg_nifti"].cpu().numpy().tobytes() == seg.tobytes(order="F")
    assert np.array_equal(out["counts"].cpu().numpy(), O.export_counts(seg))


@pytest.mark.parametrize("shape", [FULL, RAGGED])
def test_snapshot_frames_bit_exact(shape):
    lab = random_labels(shape, 11)
    got = V.snapshot_frames(dev(lab), V.PALETTE_PREDICT).cpu().numpy()
    want = O.snapshot_predict(lab)                                                      # (H, W, 3, T)
    assert np.array_equal(got, want.transpose(3, 0, 1, 2))
    got = V.snapshot_frames(dev(lab), V.PALETTE_SIMPLE).cpu().numpy()
    for z in (0, shape[2] // 2, shape[2] - 1):
        assert np.array_equal(got[z], O.snapshot_simple(lab[:, :, z]))


@pytest.mark.parametrize("shape", [FULL, RAGGED])
def test_slice_counts_and_rows_equal_reference(shape):
    lab, tgt = random_labels(shape, 12), blob_labels(shape, 13)
    counts = V.slice_counts(dev(lab), dev(tgt)).cpu().numpy()
    for z in range(shape[2]):
        for r, (a, b) in enumerate(zip(O.regions(lab[:, :, z]), O.regions(tgt[:, :, z]))):
            assert list(counts[z, 3 * r:3 * r + 3]) == [a.sum(), b.sum(), (a & b).sum()], (z, r)
    assert V.slice_dice_rows("s", counts) == O.slice_rows("s", lab, tgt)


def synthetic_mri(shape, seed):
    """integer intensities inside an ellipsoid 'brain', zero background, one modality with a hole"""
    rng = np.random.RandomState(seed)
    g = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    brain = (g[0] ** 2 + g[1] ** 2 / 0.8 + g[2] ** 2 / 0.9) < 0.7
    img = rng.randint(1, 3000, shape + (4,)).astype(np.float32) * brain[..., None]
    img[..., 2] *= (g[0] > -0.2)
    return img


@pytest.mark.parametrize("shape", [FULL, RAGGED])
def test_preprocess_volume_matches_recipe(shape):
    img = synthetic_mri(shape, 20)                                                      # (X, Y, Z, 4)
    zp = shape[2] + 5
    want = O.preprocess(img, zp)                                                        # (4, X, Y, zp)
    storage = np.ascontiguousarray(img.transpose(3, 2, 1, 0))                           # (4, Z, Y, X) as NIfTI stores
    got, stats = V.preprocess_volume(dev(storage), zp)
    got, stats = got.cpu().numpy()[0], stats.cpu().numpy()
    mask = img.sum(-1) > 0
    assert stats[8] == mask.sum()
    for k in range(4):
        y = img[..., k][mask].astype(np.float64)
        assert abs(stats[9 + k] - y.mean()) <= 1e-6 * abs(y.mean())
        assert abs(stats[13 + k] - y.std()) <= 1e-6 * y.std()
    assert got.shape == want.shape
    assert np.all(got[:, :, :, shape[2]:] == 0)
    Z = shape[2]
    assert np.array_equal(got[..., :Z][:, ~mask], want[..., :Z][:, ~mask])              # untouched outside the mask
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err <= 1e-5, err                                                             # tolerance: fp32 mean / std estimates


def test_reorder_labels_and_load_case(tmp_path):
    shape = RAGGED
    img = synthetic_mri(shape, 21)
    seg = O.export_seg(blob_labels(shape, 22, shift=1.0))                               # labels {0,1,2,4}
    case = tmp_path / "BraTS_X"
    case.mkdir()
    for k, m in enumerate(V.MODALITIES):
        O.write_nifti_numpy(str(case / f"BraTS_X_{m}.nii.gz"), img[..., k].astype(np.int16))
    O.write_nifti_numpy(str(case / "BraTS_X_seg.nii.gz"), seg)
    t = V.reorder_labels(dev(seg.transpose(2, 1, 0)), shape[2] + 3, map4to3=True).cpu().numpy()
    want = np.pad(np.where(seg == 4, 3, seg), ((0, 0), (0, 0), (0, 3)))
    assert np.array_equal(t, want)
    x, target = V.load_case(str(case), z_pad=shape[2] + 3)
    ref = O.preprocess(img, shape[2] + 3)
    assert tuple(x.shape) == (1, 4) + (shape[0], shape[1], shape[2] + 3)
    assert np.abs(x.cpu().numpy()[0] - ref).max() / np.abs(ref).max() <= 1e-5
    assert target.dtype == torch.int64 and np.array_equal(target.cpu().numpy()[0], np.pad(seg, ((0, 0), (0, 0), (0, 3))))


def test_save_prediction_files(tmp_path):
    lab = random_labels(RAGGED, 23)
    out = V.save_prediction(dev(lab), str(tmp_path / "sub"), "case1", "nii", snapshot=True, visual=str(tmp_path / "vis"))
    h = O.parse_nifti(out["path"])
    assert out["path"].endswith("case1.nii.gz") and np.array_equal(h["data"], O.export_seg(lab))
    assert np.array_equal(out["counts"], O.export_counts(O.export_seg(lab)))
    from PIL import Image
    snap = O.snapshot_predict(lab)
    assert out["frames"] == RAGGED[2]
    for z in (0, RAGGED[2] - 1):
        assert np.array_equal(np.asarray(Image.open(tmp_path / "vis" / "case1" / f"{z}.png")), snap[:, :, :, z])
    out = V.save_prediction(dev(lab), str(tmp_path / "sub"), "case1", "npy")
    got = np.load(out["path"])
    assert got.dtype == np.int64 and np.array_equal(got, lab)


def check_hausdorff(lab, tgt):
    got = V.hausdorff(dev(lab), dev(tgt))
    want95, wanthd = O.cal_hausdorff(lab, tgt), O.cal_hd(lab, tgt)
    assert got["hd95"] == want95, (got, want95)                                         # bit-exact doubles
    assert got["hd"] == wanthd, (got, wanthd)
    return got


def test_hausdorff_blobs_full_volume():
    got = check_hausdorff(blob_labels(FULL, 30), blob_labels(FULL, 31))
    assert got["hd"][0] > 5 and min(got["surface_voxels"]) > 0


def test_hausdorff_shifted_prediction_and_small_shapes():
    tgt = blob_labels((96, 80, 72), 32, shift=0.6)
    lab = np.roll(tgt, (3, -2, 1), axis=(0, 1, 2))
    check_hausdorff(lab, tgt)
    check_hausdorff(blob_labels(RAGGED, 33, shift=1.0), blob_labels(RAGGED, 34, shift=1.0))
    check_hausdorff(random_labels(RAGGED, 35), random_labels(RAGGED, 36))               # every voxel is a border voxel
    check_hausdorff(random_labels((64, 64, 155), 37), blob_labels((64, 64, 155), 38, shift=0.8))


def test_hausdorff_edge_cases():
    shape = (24, 30, 40)
    a, b = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
    a[2, 3, 4] = 3
    b[20, 25, 39] = 3                                                                   # far single voxels, one on the wall
    got = check_hausdorff(a, b)
    assert got["hd"][2] == float(np.sqrt(18 ** 2 + 22 ** 2 + 35 ** 2))
    check_hausdorff(np.zeros(shape, np.uint8), b)                                       # empty prediction -> 0
    assert V.cal_hausdorff(dev(np.zeros(shape, np.uint8)), dev(b)) == [0.0, 0.0, 0.0]
    full = np.full(shape, 3, np.uint8)
    check_hausdorff(full, b)                                                            # full prediction -> 0
    a[:] = 0
    a[5:15, 5:15, 5:15] = 2                                                             # WT only: TC / ET empty -> 0
    b[:] = 0
    b[8:20, 5:15, 0:10] = 2
    got = check_hausdorff(a, b)
    assert got["hd95"][1:] == [0.0, 0.0] and got["hd95"][0] > 0


def test_dropin_validate_softmax_exports_and_metrics(seed0_state_dict, tmp_path):
    """validate_softmax(..., savepath, save_format='nii', snapshot=True) through the drop-in modules writes what the
    reference's export block describes (predict.py:310-350); utils.hausdorff / utils.tools keep their signatures."""
    import predict_overlap
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    from tests.util import volume_input, volume_target
    from utils import hausdorff as H
    from utils.tools import cal_hausdorff, softmax_output_mIou
    model = get_cls_wise_former("brats", True, "fixed", 0)
    model.load_state_dict(seed0_state_dict)
    model.precision = __import__("dcl_b200").Precision.BF16
    model.compute_aux = False
    model.deterministic = True
    x = torch.cat([volume_input(0), torch.zeros(1, 4, 240, 240, 5)], -1)              # the loader pads 155 -> 160
    target = torch.from_numpy(np.pad(volume_target(0), ((0, 0), (0, 0), (0, 5))))[None]
    loader = [(x, target)]
    with torch.no_grad():
        wt, tc, et = predict_overlap.validate_softmax(loader, model, None, False, savepath=str(tmp_path / "out"),
                                                      names=["case0"], save_format="nii", snapshot=True,
                                                      visual=str(tmp_path / "vis"), valid_in_train=True)
    seg = O.parse_nifti(str(tmp_path / "out" / "case0.nii.gz"))["data"]
    assert seg.shape == (240, 240, 155) and set(np.unique(seg)) <= {0, 1, 2, 4}
    lab = np.where(seg == 4, 3, seg)
    d = O.softmax_output_dice(lab, volume_target(0))
    assert np.allclose([wt, tc, et], d, atol=1e-12)
    assert len(os.listdir(tmp_path / "vis" / "case0")) == 155
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(tmp_path / "vis" / "case0" / "77.png")), O.snapshot_predict(lab)[:, :, :, 77])
    # metrics drop-ins on the exported map
    small_o, small_t = lab[60:140, 60:140, 40:100], volume_target(0)[60:140, 60:140, 40:100]
    assert cal_hausdorff(small_o, small_t) == O.cal_hausdorff(small_o, small_t)
    assert softmax_output_mIou(small_o, small_t) == [float(v) for v in O.softmax_output_mIou(small_o, small_t)]
    a, b = blob_labels(RAGGED, 60, 1.0) > 0, blob_labels(RAGGED, 61, 1.0) > 0
    assert H.hausdorff_distance_95(a, b) == float(O.medpy_hd95(a, b)) and H.hausdorff_distance(a, b) == float(O.medpy_hd(a, b))
    assert H.hausdorff_distance_95(np.zeros_like(a), b) == 0


def test_device_metrics_match_reference_goldens():
    """HD95 / HD on the device against the values the reference's own cal_hausdorff / hausdorff_distance glue produced in
    the build container (tests/golden/make_golden_metrics.py; only medpy itself substituted): bit-exact doubles."""
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(here, "metrics_golden.json")))
    arrays = np.load(os.path.join(here, "metrics_cases.npz"))
    for name, g in gold.items():
        got = V.hausdorff(dev(arrays[name + "/output"]), dev(arrays[name + "/target"]))
        assert got["hd95"] == g["cal_hausdorff"], (name, got)
        assert got["hd"] == g["hausdorff_distance"], (name, got)
