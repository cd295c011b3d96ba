"""TEST INFRASTRUCTURE -- CPU restatement (numpy / scipy) of the reference code either side of the sliding
window: label export, the input recipe and the surface-distance metrics.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.

Pins
  * export_seg / export_counts / snapshot / slice_rows follow the reference expressions literally
    (predict.py:317-350, predict_simple.py:186-243) -- they ARE the reference's numpy, re-typed as functions.
  * hd95 / hd: the reference calls medpy (utils/hausdorff.py:86-123, utils/tools.py:64-86), a third-party dependency
    that is neither vendored nor version-pinned (no requirements file) and is absent from this image.  medpy 0.4.0's
    published algorithm (medpy/metric/binary.py: __surface_distances, hd, hd95) is restated on the same scipy.ndimage
    primitives medpy itself calls.  PARITY UNPINNED beyond that: no medpy run is available to check against.
  * preprocess: data/ClsWiseBraTS128Test.py is imported by the reference (test_overlap.py:14) but NOT shipped; the recipe
    below is the TransBTS one the predict scripts descend from (mask = sum over modalities > 0, per-modality z-score over
    the mask, pad z 155 -> 160).  PARITY UNPINNED.
  * parse_nifti is an independent pure-numpy reader used to check the C writer (nibabel is absent).
"""
import gzip

import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


# ---- export (predict.py:317-350) --------------------------------------------------------------
def export_seg(output):
    """seg_img of predict.py:320-324"""
    H, W, T = output.shape
    seg_img = np.zeros(shape=(H, W, T), dtype=np.uint8)
    seg_img[np.where(output == 1)] = 1
    seg_img[np.where(output == 2)] = 2
    seg_img[np.where(output == 3)] = 4
    return seg_img


def export_counts(seg_img):
    """the verbose print of predict.py:325-328: n(1), n(2), n(4), WT, TC, ET"""
    return np.array([np.sum(seg_img == 1), np.sum(seg_img == 2), np.sum(seg_img == 4),
                     np.sum((seg_img == 1) | (seg_img == 2) | (seg_img == 4)),
                     np.sum((seg_img == 1) | (seg_img == 4)), np.sum(seg_img == 4)], dtype=np.int64)


def snapshot_predict(output):
    """Snapshot_img of predict.py:338-344, (H, W, 3, T)"""
    H, W, T = output.shape
    Snapshot_img = np.zeros(shape=(H, W, 3, T), dtype=np.uint8)
    Snapshot_img[:, :, 0, :][np.where(output == 0)] = 255
    Snapshot_img[:, :, 1, :][np.where(output == 2)] = 255
    Snapshot_img[:, :, 2, :][np.where(output == 3)] = 255
    return Snapshot_img


def snapshot_simple(item):
    """one frame of output_pic (predict_simple.py:192-197)"""
    img = np.zeros(shape=item.shape + (3,), dtype=np.uint8)
    img[:, :][np.where(item == 1)] = [250, 250, 149]
    img[:, :][np.where(item == 2)] = [244, 130, 128]
    img[:, :][np.where(item == 3)] = [97, 136, 200]
    return img


def dice_score(o, t, eps=1e-8):
    """utils/tools.py:44-47"""
    num = 2 * (o * t).sum() + eps
    den = o.sum() + t.sum() + eps
    return num / den


def softmax_output_dice(output, target):
    """utils/tools.py:89-109"""
    return [dice_score(output > 0, target > 0),
            dice_score((output == 1) | (output == 3), (target == 1) | (target == 3)),
            dice_score(output == 3, target == 3)]


def mIOU(o, t, eps=1e-8):
    """predict_simple.py:66-69"""
    num = (o * t).sum() + eps
    den = (o | t).sum() + eps
    return num / den


def softmax_output_mIou(output, target):
    """predict_simple.py:100-118"""
    return [mIOU(output > 0, target > 0),
            mIOU((output == 1) | (output == 3), (target == 1) | (target == 3)),
            mIOU(output == 3, target == 3)]


def slice_rows(name, output, label):
    """the rows output_excel collects (predict_simple.py:232-243), over every frame of the arrays"""
    rows = []
    for frame in range(output.shape[2]):
        item, label_item = output[:, :, frame], label[:, :, frame]
        if label_item.max() > 0:
            dice = softmax_output_dice(item, label_item)
            rows.append({"name": name + "_" + str(frame), "wt": dice[0], "tc": dice[1], "et": dice[2],
                         "sum": dice[0] * dice[1] * dice[2]})
    return rows


# ---- medpy.metric.binary (0.4.0) restated -----------------------------------------------------
def surface_distances(result, reference, connectivity=1):
    """medpy/metric/binary.py::__surface_distances with voxelspacing=None"""
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if 0 == np.count_nonzero(result):
        raise RuntimeError('The first supplied array does not contain any binary object.')
    if 0 == np.count_nonzero(reference):
        raise RuntimeError('The second supplied array does not contain any binary object.')
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=None)
    return dt[result_border]


def medpy_hd(result, reference):
    return max(surface_distances(result, reference).max(), surface_distances(reference, result).max())


def medpy_hd95(result, reference):
    hd1 = surface_distances(result, reference)
    hd2 = surface_distances(reference, result)
    return np.percentile(np.hstack((hd1, hd2)), 95)


def _guarded(fn, test, reference):
    """utils/hausdorff.py:86-123 with nan_for_nonexisting=False: 0 for an empty or full mask"""
    if (not np.any(test)) or np.all(test) or (not np.any(reference)) or np.all(reference):
        return 0
    return fn(test, reference)


def regions(lab):
    return [lab > 0, (lab == 1) | (lab == 3), lab == 3]


def cal_hausdorff(output, target):
    """predict_simple.py:121-144 -> [wt, tc, et] HD95"""
    return [float(_guarded(medpy_hd95, o, t)) for o, t in zip(regions(output), regions(target))]


def cal_hd(output, target):
    return [float(_guarded(medpy_hd, o, t)) for o, t in zip(regions(output), regions(target))]


# ---- input recipe (TransBTS lineage; the reference's loader is not shipped) --------------------
def preprocess(images_xyzc, z_pad=160):
    """images (X, Y, Z, 4) float32 -> (4, X, Y, z_pad): mask = images.sum(-1) > 0; x[mask] -= mean; x[mask] /= std"""
    images = np.array(images_xyzc, dtype=np.float32, copy=True)
    mask = images.sum(-1) > 0
    for k in range(4):
        x = images[..., k]
        y = x[mask]
        x[mask] -= y.mean()
        x[mask] /= y.std()
        images[..., k] = x
    images = np.pad(images, ((0, 0), (0, 0), (0, z_pad - images.shape[2]), (0, 0)), mode='constant')
    return np.ascontiguousarray(images.transpose(3, 0, 1, 2))


# ---- independent NIfTI-1 parser -----------------------------------------------------------------
_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32}


def parse_nifti(path):
    """-> dict(header fields, data as an (X,Y,Z) array like nibabel's dataobj)"""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    hdr = raw[:348]
    out = {
        "sizeof_hdr": int(np.frombuffer(hdr, "<i4", 1, 0)[0]),
        "dim": np.frombuffer(hdr, "<i2", 8, 40).copy(),
        "datatype": int(np.frombuffer(hdr, "<i2", 1, 70)[0]),
        "bitpix": int(np.frombuffer(hdr, "<i2", 1, 72)[0]),
        "pixdim": np.frombuffer(hdr, "<f4", 8, 76).copy(),
        "vox_offset": float(np.frombuffer(hdr, "<f4", 1, 108)[0]),
        "scl_slope": float(np.frombuffer(hdr, "<f4", 1, 112)[0]),
        "scl_inter": float(np.frombuffer(hdr, "<f4", 1, 116)[0]),
        "qform_code": int(np.frombuffer(hdr, "<i2", 1, 252)[0]),
        "sform_code": int(np.frombuffer(hdr, "<i2", 1, 254)[0]),
        "magic": hdr[344:348],
    }
    X, Y, Z = (int(v) for v in out["dim"][1:4])
    dt = np.dtype(_NIFTI_DTYPES[out["datatype"]]).newbyteorder("<")
    off = int(out["vox_offset"])
    out["data"] = np.frombuffer(raw, dt, X * Y * Z, off).reshape((X, Y, Z), order="F")
    return out


def write_nifti_numpy(path, array_xyz, slope=None, inter=None):
    """test fixture writer, independent of the C writer: a minimal NIfTI-1 single file"""
    a = np.asarray(array_xyz)
    code = {v: k for k, v in _NIFTI_DTYPES.items()}[a.dtype.type]
    hdr = bytearray(348)
    hdr[0:4] = np.int32(348).tobytes()
    dim = np.array([3, a.shape[0], a.shape[1], a.shape[2], 1, 1, 1, 1], dtype="<i2")
    hdr[40:56] = dim.tobytes()
    hdr[70:72] = np.int16(code).tobytes()
    hdr[72:74] = np.int16(a.dtype.itemsize * 8).tobytes()
    hdr[76:108] = np.ones(8, dtype="<f4").tobytes()
    hdr[108:112] = np.float32(352).tobytes()
    hdr[112:116] = np.float32(np.nan if slope is None else slope).tobytes()
    hdr[116:120] = np.float32(np.nan if inter is None else inter).tobytes()
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + a.tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(payload)
