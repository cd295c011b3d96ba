"""Host side of the two ends of the sliding-window path: label export, the input
side and the surface-distance metrics.  Like ``engine.py`` this module only marshals pointers into the C ABI
(``include/dcl_b200.h``); the arithmetic is in ``csrc/volio.cu`` and ``csrc/hausdorff.cu``.

Reference call sites
  * export: ``predict.py:310-350`` (savepath / save_format / snapshot arguments of ``validate_softmax``),
    ``predict_simple.py:186-278`` (per-slice pictures and Dice tables)
  * input: ``data/ClsWiseBraTS128Test.BraDataSet128`` (``test_overlap.py:14,94-97``; not shipped by the reference)
  * metrics: ``cal_hausdorff`` ``predict_simple.py:121-144`` -> ``utils/hausdorff.py:86-123``;
    ``softmax_output_mIou`` ``predict_simple.py:100-118``
"""
from __future__ import annotations

import csv
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N
from ._native import DclError

PALETTE_PREDICT = (255, 0, 0, 0, 0, 0, 0, 255, 0, 0, 0, 255)                  # predict.py:342-344
PALETTE_SIMPLE = (0, 0, 0, 250, 250, 149, 244, 130, 128, 97, 136, 200)        # predict_simple.py:193-197
MODALITIES = ("flair", "t1ce", "t1", "t2")                                     # TransBTS order of the four channels
NIFTI_CODES = {np.dtype(np.uint8): 2, np.dtype(np.int16): 4, np.dtype(np.int32): 8, np.dtype(np.float32): 16,
               np.dtype(np.float64): 64, np.dtype(np.int8): 256, np.dtype(np.uint16): 512, np.dtype(np.uint32): 768}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _shape3(t):
    return (C.c_int32 * 3)(*[int(v) for v in t.shape[-3:]])


def _need_cuda_u8(name, t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3):
        raise DclError(f"{name} must be a CUDA uint8 (X,Y,Z) tensor (dcl_b200 has no CPU path)")
    return t.contiguous()


# ---- export -----------------------------------------------------------------------------------
def export_labels(labels, want_seg=True, want_nifti_order=True):
    """labels (X,Y,Z) uint8 in {0..3} -> dict(seg (X,Y,Z) with 3->4, seg_nifti (Z,Y,X) = NIfTI storage order,
    counts int64[6] = n(1), n(2), n(4), WT, TC, ET)."""
    lib = N.load_library()
    labels = _need_cuda_u8("labels", labels)
    seg = torch.empty_like(labels) if want_seg else None
    X, Y, Z = labels.shape
    turned = torch.empty((Z, Y, X), dtype=torch.uint8, device=labels.device) if want_nifti_order else None
    counts = torch.zeros(6, dtype=torch.int64, device=labels.device)
    N.check(lib.dcl_export_labels(_ptr(labels), _shape3(labels), _ptr(seg), _ptr(turned), _ptr(counts), _stream()))
    return {"seg": seg, "seg_nifti": turned, "counts": counts}


def snapshot_frames(labels, palette=PALETTE_PREDICT):
    """(Z, X, Y, 3) uint8 frames: frame z = Snapshot_img[:, :, :, z] (predict.py:338-350)."""
    lib = N.load_library()
    labels = _need_cuda_u8("labels", labels)
    X, Y, Z = labels.shape
    pal = (C.c_uint8 * 12)(*[int(v) for v in palette])
    out = torch.empty((Z, X, Y, 3), dtype=torch.uint8, device=labels.device)
    N.check(lib.dcl_snapshot_frames(_ptr(labels), _shape3(labels), pal, _ptr(out), _stream()))
    return out


def slice_counts(labels, target):
    """(Z, 9) int64: per z slice (|o|,|t|,|o&t|) for WT, TC, ET."""
    lib = N.load_library()
    labels, target = _need_cuda_u8("labels", labels), _need_cuda_u8("target", target)
    out = torch.zeros((labels.shape[2], 9), dtype=torch.int64, device=labels.device)
    N.check(lib.dcl_slice_counts(_ptr(labels), _ptr(target), _shape3(labels), _ptr(out), _stream()))
    return out


def slice_dice_rows(name, counts):
    """The rows ``output_excel`` collects (predict_simple.py:224-243): one per frame whose label slice is not empty,
    dict(name, wt, tc, et, sum = wt*tc*et), Dice with the reference's eps (utils/tools.py:44-47)."""
    rows = []
    for z, c in enumerate(np.asarray(counts).astype(np.int64)):
        if c[1] == 0:                                   # label_item.max() > 0  <=>  the slice has a WT target voxel
            continue
        d = [(2 * int(c[3 * r + 2]) + 1e-8) / (int(c[3 * r]) + int(c[3 * r + 1]) + 1e-8) for r in range(3)]
        rows.append({"name": f"{name}_{z}", "wt": d[0], "tc": d[1], "et": d[2], "sum": d[0] * d[1] * d[2]})
    return rows


def write_slice_tables(directory, modal, name, rows):
    """export_item_excel (predict_simple.py:266-278): three tables sorted by wt / tc / et."""
    os.makedirs(directory, exist_ok=True)
    paths = []
    for region in ("wt", "tc", "et"):
        path = os.path.join(directory, f"{modal}_{name}_{region}.csv")
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=["name", "wt", "tc", "et", "sum"])
            w.writeheader()
            for r in sorted(rows, key=lambda r: r[region]):
                w.writerow(r)
        paths.append(path)
    return paths


def write_nifti(path, array_xyz):
    """nib.save(nib.Nifti1Image(array, None), path) for a C-ordered (X,Y,Z) numpy array or a (Z,Y,X) storage-order
    tensor from ``export_labels``: pass ``storage=`` arrays through ``write_nifti_storage``."""
    a = np.asarray(array_xyz)
    return write_nifti_storage(path, np.ascontiguousarray(a.transpose(2, 1, 0)))


def write_nifti_storage(path, storage_zyx):
    """storage_zyx: (Z,Y,X) C-contiguous numpy array = NIfTI order (x fastest)."""
    lib = N.load_library()
    a = np.ascontiguousarray(storage_zyx)
    if a.dtype not in NIFTI_CODES:
        raise DclError(f"unsupported NIfTI dtype {a.dtype}")
    Z, Y, X = a.shape
    shape = (C.c_int32 * 3)(X, Y, Z)
    N.check(lib.dcl_write_nifti(os.fsencode(path), a.ctypes.data_as(C.c_void_p), NIFTI_CODES[a.dtype], shape))
    return path


def write_npy_labels(path, labels_host):
    """np.save(path, output) of the int64 arg-max map (predict.py:313-314) from uint8 labels."""
    lib = N.load_library()
    a = np.ascontiguousarray(np.asarray(labels_host, dtype=np.uint8))
    shape = (C.c_int32 * 3)(*a.shape)
    N.check(lib.dcl_write_npy_labels(os.fsencode(path), a.ctypes.data_as(C.c_void_p), shape))
    return path


def write_png(path, rgb_host):
    lib = N.load_library()
    a = np.ascontiguousarray(np.asarray(rgb_host, dtype=np.uint8))
    if a.ndim != 3 or a.shape[2] != 3:
        raise DclError("write_png takes an (H, W, 3) uint8 array")
    N.check(lib.dcl_write_png_rgb(os.fsencode(path), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]))
    return path


def save_prediction(labels, savepath, name, save_format="nii", snapshot=False, visual="", verbose=False):
    """The export block of validate_softmax (predict.py:310-350) on device labels: '<name>_preds.npy' or
    '<name>.nii.gz' (3 -> 4), optional per-frame PNG snapshots under visual/<name>/<frame>.png."""
    assert save_format in ("npy", "nii")
    os.makedirs(savepath, exist_ok=True)
    out = {}
    if save_format == "npy":
        out["path"] = write_npy_labels(os.path.join(savepath, name + "_preds.npy"), labels.cpu().numpy())
    else:
        ex = export_labels(labels, want_seg=False, want_nifti_order=True)
        c = ex["counts"].cpu().numpy()
        if verbose:
            print('1:', c[0], ' | 2:', c[1], ' | 4:', c[2])
            print('WT:', c[3], ' | TC:', c[4], ' | ET:', c[5])
        out["path"] = write_nifti_storage(os.path.join(savepath, name + ".nii.gz"), ex["seg_nifti"].cpu().numpy())
        out["counts"] = c
        print('Successfully save {}'.format(out["path"]))
    if snapshot:
        frames = snapshot_frames(labels).cpu().numpy()
        os.makedirs(os.path.join(visual, name), exist_ok=True)
        for z in range(frames.shape[0]):
            write_png(os.path.join(visual, name, str(z) + ".png"), frames[z])
        out["frames"] = frames.shape[0]
    return out


# ---- input side ---------------------------------------------------------------------------------
def read_nifti(path):
    """(array (Z,Y,X) float32 in storage order, (X,Y,Z), pixdim) -- a C-ordered view of nibabel's get_fdata().T"""
    lib = N.load_library()
    shape, dt, pix = (C.c_int32 * 3)(), C.c_int32(), (C.c_float * 3)()
    N.check(lib.dcl_read_nifti_header(os.fsencode(path), shape, C.byref(dt), pix))
    X, Y, Z = (int(v) for v in shape)
    out = np.empty((Z, Y, X), dtype=np.float32)
    N.check(lib.dcl_read_nifti_f32(os.fsencode(path), out.ctypes.data_as(C.c_void_p), out.size))
    return out, (X, Y, Z), tuple(float(v) for v in pix)


def preprocess_volume(modalities, z_pad=160):
    """modalities: CUDA float32 (4, Z, Y, X) (four NIfTI arrays as stored) -> (1, 4, X, Y, z_pad) z-scored over the brain
    mask + stats dict(count, mean[4], std[4])."""
    lib = N.load_library()
    if not (modalities.is_cuda and modalities.dtype == torch.float32 and modalities.dim() == 4 and modalities.shape[0] == 4):
        raise DclError("modalities must be a CUDA fp32 (4,Z,Y,X) tensor")
    modalities = modalities.contiguous()
    Z, Y, X = (int(v) for v in modalities.shape[1:])
    out = torch.empty((1, 4, X, Y, int(z_pad)), dtype=torch.float32, device=modalities.device)
    stats = torch.zeros(17, dtype=torch.float64, device=modalities.device)
    shape = (C.c_int32 * 3)(X, Y, Z)
    N.check(lib.dcl_preprocess_volume(_ptr(modalities), shape, int(z_pad), _ptr(out), _ptr(stats), _stream()))
    return out, stats


def reorder_labels(seg_storage, z_pad=160, map4to3=False):
    """seg_storage: CUDA uint8 (Z,Y,X) -> (X,Y,z_pad) target, optionally 4 -> 3."""
    lib = N.load_library()
    if not (seg_storage.is_cuda and seg_storage.dtype == torch.uint8 and seg_storage.dim() == 3):
        raise DclError("seg_storage must be a CUDA uint8 (Z,Y,X) tensor")
    seg_storage = seg_storage.contiguous()
    Z, Y, X = (int(v) for v in seg_storage.shape)
    out = torch.empty((X, Y, int(z_pad)), dtype=torch.uint8, device=seg_storage.device)
    shape = (C.c_int32 * 3)(X, Y, Z)
    N.check(lib.dcl_reorder_labels(_ptr(seg_storage), shape, int(z_pad), int(bool(map4to3)), _ptr(out), _stream()))
    return out


def load_case(case_dir, device="cuda", z_pad=160, with_seg=True):
    """One BraTS subject directory (<id>_flair.nii.gz ... <id>_seg.nii.gz) -> (x (1,4,X,Y,z_pad), target (1,X,Y,z_pad)
    int64 | None), the tuple head BraDataSet128 yields (predict_overlap.py:132-135)."""
    case_dir = os.path.normpath(case_dir)
    cid = os.path.basename(case_dir)

    def find(suffix):
        for ext in (".nii.gz", ".nii"):
            p = os.path.join(case_dir, f"{cid}_{suffix}{ext}")
            if os.path.exists(p):
                return p
        raise DclError(f"{case_dir}: no {cid}_{suffix}.nii[.gz]")

    arrays = [read_nifti(find(m))[0] for m in MODALITIES]
    host = torch.from_numpy(np.stack(arrays, 0)).pin_memory()
    x, _ = preprocess_volume(host.to(device, non_blocking=True), z_pad)
    target = None
    if with_seg:
        seg = torch.from_numpy(read_nifti(find("seg"))[0].astype(np.uint8)).to(device)
        target = reorder_labels(seg, z_pad)[None].long()
    return x, target


# ---- metrics ----------------------------------------------------------------------------------
def hausdorff(labels, target):
    """cal_hausdorff (predict_simple.py:121-144) + the plain Hausdorff distance of utils/tools.py:64-86 on device label
    maps: dict(hd95 [wt, tc, et], hd [..], surface_voxels [..])."""
    lib = N.load_library()
    labels, target = _need_cuda_u8("labels", labels), _need_cuda_u8("target", target)
    if labels.shape != target.shape:
        raise DclError("labels and target must have the same shape")            # utils/hausdorff.py:4-7
    shape = _shape3(labels)
    nbytes = int(N.check(lib.dcl_hausdorff_workspace_bytes(shape)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=labels.device)
    hd95, hd, cnt = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_uint64 * 3)()
    N.check(lib.dcl_hausdorff(_ptr(labels), _ptr(target), shape, _ptr(ws), nbytes, hd95, hd, cnt, _stream()))
    return {"hd95": [float(v) for v in hd95], "hd": [float(v) for v in hd], "surface_voxels": [int(v) for v in cnt]}


def cal_hausdorff(labels, target):
    return hausdorff(labels, target)["hd95"]


def miou_from_counts(counts):
    """softmax_output_mIou (predict_simple.py:100-118) on the 13 counters of the label kernel:
    (|o&t| + eps) / (|o|t| + eps) with |o|t| = |o| + |t| - |o&t|."""
    c = [int(v) for v in counts]
    return [(c[6 + 3 * r] + 1e-8) / (c[4 + 3 * r] + c[5 + 3 * r] - c[6 + 3 * r] + 1e-8) for r in range(3)]


def percentile_from_hist(hist, q):
    lib = N.load_library()
    h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint32))
    return float(lib.dcl_percentile_from_hist(h.ctypes.data_as(C.c_void_p), h.size, float(q)))
