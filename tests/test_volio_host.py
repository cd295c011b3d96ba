"""CPU-only checks of the host side of SURVEY 8f ranks 2-4: the file writers / reader of the C library
(pure host code), numpy's percentile rule restated in C, and the metrics oracle on known answers."""
import gzip
import os

import numpy as np
import pytest

from dcl_b200 import volio as V
from dcl_b200 import DclError
from oracle import volio_oracle as O


def test_npy_writer_is_byte_identical_to_numpy_save(tmp_path):
    rng = np.random.RandomState(0)
    for shape in ((240, 240, 155), (7, 5, 3), (1000, 2, 1)):
        lab = rng.randint(0, 4, shape).astype(np.uint8)
        ours, ref = tmp_path / "ours.npy", tmp_path / "ref.npy"
        V.write_npy_labels(str(ours), lab)
        np.save(ref, lab.astype(np.int64))                          # predict.py:314: output is the int64 argmax
        assert ours.read_bytes() == ref.read_bytes(), shape
        assert np.array_equal(np.load(ours), lab)


@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_nifti_writer_round_trip_and_header(tmp_path, ext):
    rng = np.random.RandomState(1)
    seg = O.export_seg(rng.randint(0, 4, (24, 20, 15)))
    path = str(tmp_path / ("seg" + ext))
    V.write_nifti(path, seg)
    h = O.parse_nifti(path)
    assert h["sizeof_hdr"] == 348 and h["magic"] == b"n+1\0" and h["vox_offset"] == 352.0
    assert list(h["dim"]) == [3, 24, 20, 15, 1, 1, 1, 1] and h["datatype"] == 2 and h["bitpix"] == 8
    assert list(h["pixdim"][1:4]) == [1.0, 1.0, 1.0] and h["qform_code"] == 0 and h["sform_code"] == 0
    assert np.isnan(h["scl_slope"])                                 # "not scaled", as nibabel writes uint8 data
    assert np.array_equal(h["data"], seg)
    if ext == ".nii.gz":
        assert open(path, "rb").read(2) == b"\x1f\x8b"
    # our reader on our writer
    back, shape, pix = V.read_nifti(path)
    assert shape == (24, 20, 15) and pix == (1.0, 1.0, 1.0)
    assert np.array_equal(back.transpose(2, 1, 0), seg.astype(np.float32))


def test_nifti_reader_datatypes_scaling_and_errors(tmp_path):
    rng = np.random.RandomState(2)
    cases = [(rng.randint(-500, 4000, (9, 8, 7)).astype(np.int16), None, None),
             (rng.randint(0, 4000, (9, 8, 7)).astype(np.int16), 0.5, 10.0),
             (rng.randn(5, 6, 7).astype(np.float32), None, None),
             (rng.randint(0, 5, (5, 6, 7)).astype(np.uint8), 1.0, 0.0),
             (rng.randint(0, 60000, (5, 6, 7)).astype(np.uint16), 0.0, 0.0)]
    for i, (a, slope, inter) in enumerate(cases):
        path = str(tmp_path / f"c{i}.nii.gz")
        O.write_nifti_numpy(path, a, slope, inter)
        got, shape, _ = V.read_nifti(path)
        want = a.astype(np.float64)
        if slope not in (None, 0.0):
            want = want * slope + (inter or 0.0)
        assert shape == a.shape
        assert np.array_equal(got.transpose(2, 1, 0), want.astype(np.float32)), i
    bad = tmp_path / "bad.nii"
    bad.write_bytes(b"\0" * 400)
    with pytest.raises(DclError):
        V.read_nifti(str(bad))
    with pytest.raises(DclError):
        V.read_nifti(str(tmp_path / "missing.nii.gz"))
    trunc = tmp_path / "trunc.nii"
    O.write_nifti_numpy(str(trunc), cases[0][0])
    trunc.write_bytes(trunc.read_bytes()[:600])
    with pytest.raises(DclError):
        V.read_nifti(str(trunc))


def test_png_writer_decodes_to_the_same_pixels(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(3)
    lab = rng.randint(0, 4, (33, 47))
    frame = O.snapshot_simple(lab)
    path = str(tmp_path / "f.png")
    V.write_png(path, frame)
    img = np.asarray(Image.open(path))
    assert img.shape == (33, 47, 3) and np.array_equal(img, frame)


def test_percentile_from_hist_equals_numpy_percentile():
    rng = np.random.RandomState(4)
    for n in (1, 2, 3, 19, 20, 21, 40, 41, 1000, 54321):
        for trial in range(4):
            d2 = rng.randint(0, 6 if trial % 2 else 5000, n)
            hist = np.bincount(d2).astype(np.uint32)
            a = np.sqrt(d2.astype(np.float64))
            for q in (95.0, 50.0, 0.0, 100.0, 99.9):
                got, want = V.percentile_from_hist(hist, q), float(np.percentile(a, q))
                assert got == want, (n, trial, q, got, want)
    assert np.isnan(V.percentile_from_hist(np.zeros(4, np.uint32), 95.0))


def test_metrics_oracle_known_answers():
    a = np.zeros((20, 20, 20), np.uint8)
    b = np.zeros((20, 20, 20), np.uint8)
    a[5, 5, 5] = 3
    b[5, 9, 8] = 3                                                  # single voxels: distance 5 both ways
    assert O.cal_hausdorff(a, b) == [5.0, 5.0, 5.0]
    assert O.cal_hd(a, b) == [5.0, 5.0, 5.0]
    a[:] = 0; b[:] = 0
    a[4:10, 4:10, 4:10] = 1
    b[6:12, 4:10, 4:10] = 1                                         # the same cube shifted by 2 along x
    assert O.cal_hd(a, b)[:2] == [2.0, 2.0] and O.cal_hd(a, b)[2] == 0      # ET empty on both sides -> 0
    assert 0 < O.cal_hausdorff(a, b)[0] <= 2.0
    assert O.cal_hausdorff(np.zeros_like(a), b) == [0.0, 0.0, 0.0]  # empty prediction -> 0 (utils/hausdorff.py:112-120)
    assert O.cal_hausdorff(np.ones_like(a), b)[0] == 0.0            # full prediction -> 0


def test_miou_from_counts_equals_reference_expression():
    from dcl_b200.engine import dice_from_counts
    from oracle import stitch_oracle as S
    rng = np.random.RandomState(5)
    o, t = rng.randint(0, 4, (30, 31, 17)), rng.randint(0, 4, (30, 31, 17))
    counts = [int(np.sum(o == k)) for k in range(4)] + [v for c in S.region_counts(o, t) for v in c]
    assert V.miou_from_counts(counts) == [float(v) for v in O.softmax_output_mIou(o, t)]
    assert np.allclose(dice_from_counts(counts), O.softmax_output_dice(o, t), atol=0)


def test_slice_rows_from_counts_equal_reference_rows():
    rng = np.random.RandomState(6)
    o, t = rng.randint(0, 4, (12, 13, 9)), rng.randint(0, 4, (12, 13, 9))
    t[:, :, 4] = 0                                                  # an empty label slice is skipped
    counts = np.zeros((9, 9), np.int64)
    for z in range(9):
        for r, (a, b) in enumerate(zip(O.regions(o[:, :, z]), O.regions(t[:, :, z]))):
            counts[z, 3 * r:3 * r + 3] = [a.sum(), b.sum(), (a & b).sum()]
    got, want = V.slice_dice_rows("case", counts), O.slice_rows("case", o, t)
    assert len(got) == 8 and got == want


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    with pytest.raises(DclError):
        V.export_labels(torch.zeros(4, 4, 4, dtype=torch.uint8))
    with pytest.raises(DclError):
        V.hausdorff(torch.zeros(4, 4, 4, dtype=torch.uint8), torch.zeros(4, 4, 4, dtype=torch.uint8))


def _metric_goldens():
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return json.load(open(os.path.join(here, "metrics_golden.json"))), np.load(os.path.join(here, "metrics_cases.npz"))


def test_metrics_oracle_and_counter_formulas_match_reference_goldens():
    """tests/golden/make_golden_metrics.py ran the reference's own cal_hausdorff / softmax_output_mIou /
    softmax_output_dice (only medpy itself substituted): the oracle and the counter-based host formulas reproduce them."""
    from dcl_b200.engine import dice_from_counts
    from oracle import stitch_oracle as S
    gold, arrays = _metric_goldens()
    assert len(gold) == 7
    for name, g in gold.items():
        o, t = arrays[name + "/output"], arrays[name + "/target"]
        assert O.cal_hausdorff(o, t) == g["cal_hausdorff"], name
        assert O.cal_hd(o, t) == g["hausdorff_distance"], name
        counts = [int(np.sum(o == k)) for k in range(4)] + [v for c in S.region_counts(o, t) for v in c]
        assert V.miou_from_counts(counts) == g["softmax_output_mIou"], name
        assert dice_from_counts(counts) == g["softmax_output_dice"], name


def test_host_entry_points_reject_bad_arguments(tmp_path):
    import ctypes as C
    from dcl_b200 import _native as N
    lib = N.load_library()
    bad = (C.c_int32 * 3)(0, 4, 4)
    ok = (C.c_int32 * 3)(240, 240, 155)
    assert lib.dcl_hausdorff_workspace_bytes(bad) < 0
    ws = lib.dcl_hausdorff_workspace_bytes(ok)
    assert 100e6 < ws < 200e6                                      # border 9 MB + 16-bit and 32-bit transforms of two sets
    a = np.zeros((4, 4, 4), np.float32)
    assert lib.dcl_write_nifti(str(tmp_path / "x.nii").encode(), a.ctypes.data_as(C.c_void_p), 3, (C.c_int32 * 3)(4, 4, 4)) < 0
    assert b"datatype" in lib.dcl_last_error()
    assert lib.dcl_write_nifti(b"/nonexistent_dir/x.nii", a.ctypes.data_as(C.c_void_p), 16, (C.c_int32 * 3)(4, 4, 4)) < 0
    assert lib.dcl_write_png_rgb(str(tmp_path / "x.png").encode(), a.ctypes.data_as(C.c_void_p), 0, 4) < 0
    with pytest.raises(DclError):
        V.write_png(str(tmp_path / "y.png"), np.zeros((4, 4), np.uint8))
    with pytest.raises(DclError):
        V.write_nifti_storage(str(tmp_path / "z.nii"), np.zeros((2, 2, 2), np.complex64))
