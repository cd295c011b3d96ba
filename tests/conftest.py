import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200")
for p in (ROOT, os.path.join(PKG, "dropin")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def seed0_state_dict():
    """Reference-identical random-init weights (torch.manual_seed(0)) from the drop-in module."""
    import torch
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    torch.manual_seed(0)
    return {k: v.detach().clone() for k, v in get_cls_wise_former("brats", True, "fixed", 0).state_dict().items()}


@pytest.fixture(scope="session")
def golden_patch():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "patch_seed1.npz"))


@pytest.fixture(scope="session")
def golden_volume():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "volume_seed1000.npz"))
