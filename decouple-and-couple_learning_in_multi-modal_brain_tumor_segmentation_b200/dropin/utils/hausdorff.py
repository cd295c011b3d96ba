"""Drop-in for the reference ``utils/hausdorff.py`` (:86-123): same function names and arguments, distances computed
on the device by ``dcl_hausdorff`` (exact integer distance transform + histogram, csrc/hausdorff.cu) instead of medpy.
Only unit voxel spacing and connectivity 1 (the values every reference call site uses) are supported."""
import numpy as np
import torch

from dcl_b200 import volio as _V


def _run(test, reference, voxel_spacing, connectivity, key, nan_for_nonexisting):
    if voxel_spacing is not None or connectivity != 1:
        raise NotImplementedError("dcl_b200: unit voxel spacing and connectivity 1 only")
    test, reference = np.asarray(test), np.asarray(reference)
    assert test.shape == reference.shape, "Shape mismatch: {} and {}".format(test.shape, reference.shape)   # :4-7
    if (not np.any(test)) or np.all(test) or (not np.any(reference)) or np.all(reference):                   # :95-101
        return float("NaN") if nan_for_nonexisting else 0
    lab = torch.from_numpy((test != 0).astype(np.uint8) * 3).cuda()          # label 3 = the mask in every region
    tgt = torch.from_numpy((reference != 0).astype(np.uint8) * 3).cuda()
    return _V.hausdorff(lab, tgt)[key][2]


def hausdorff_distance(test=None, reference=None, confusion_matrix=None, nan_for_nonexisting=False, voxel_spacing=None,
                       connectivity=1, **kwargs):
    return _run(test, reference, voxel_spacing, connectivity, "hd", nan_for_nonexisting)


def hausdorff_distance_95(test=None, reference=None, confusion_matrix=None, nan_for_nonexisting=False,
                          voxel_spacing=None, connectivity=1, **kwargs):
    return _run(test, reference, voxel_spacing, connectivity, "hd95", nan_for_nonexisting)
