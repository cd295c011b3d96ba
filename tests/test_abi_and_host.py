"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol of
include/dcl_b200.h, the weight catalogue equals the reference state_dict schema, the product fails
loudly without a GPU, and the host-side planning logic agrees with the oracle."""
import json
import os
import re

import numpy as np
import pytest
import torch

import dcl_b200
from dcl_b200 import _native as N
from dcl_b200 import engine as E
from oracle import stitch_oracle as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dcl_b200.h")).read()
    declared = set(re.findall(r"DCL_API\s+[\w\s\*]+?\b(dcl_\w+)\s*\(", header))
    assert len(declared) >= 19
    assert declared == set(N.EXPORTED_SYMBOLS), declared ^ set(N.EXPORTED_SYMBOLS)
    lib = N.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert dcl_b200.abi_version() == 1


def test_weight_catalogue_is_the_reference_state_dict():
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    cat = dcl_b200.weight_catalogue()
    assert len(cat) == 222 and {n for n, _, _ in cat} == set(ref)
    for name, numel, aux in cat:
        assert numel == int(np.prod(ref[name])), name
        assert aux == ("supervise_label" in name.split(".")[0]), name


def test_dropin_module_matches_reference_schema_and_init(seed0_state_dict):
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    dig = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_seed0_digest.json")))
    assert list(seed0_state_dict) == list(ref)                      # same keys, same order
    for k, v in seed0_state_dict.items():
        assert list(v.shape) == ref[k], k
        assert abs(float(v.double().sum()) - dig[k][0]) < 1e-9, k   # same RNG stream as the reference
        assert abs(float(v.double().abs().sum()) - dig[k][1]) < 1e-9, k


def test_workspace_size_is_reported():
    assert 1 << 30 < dcl_b200.workspace_bytes() < 8 << 30


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    with pytest.raises(dcl_b200.DclError):
        dcl_b200.Engine()
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    model = get_cls_wise_former("brats", True, "fixed", 0)
    with pytest.raises(dcl_b200.DclError):
        model(torch.zeros(1, 4, 128, 128, 128), None)


def test_learned_positional_encoding_is_rejected_like_the_reference_fails():
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    with pytest.raises(NotImplementedError):
        get_cls_wise_former("brats", True, "learned", 0)


def test_patch_plans_agree_with_oracle():
    assert E.reference_starts() == S.REFERENCE_STARTS
    for stride in (32, 64, 96):
        assert E.patch_starts((240, 240, 155), stride) == S.patch_starts((240, 240, 155), stride)
    assert E.patch_starts((128, 128, 128), 64) == [(0, 0, 0)]


def test_dice_from_counts_agrees_with_oracle():
    rng = np.random.RandomState(3)
    o, t = rng.randint(0, 4, (40, 40, 31)), rng.randint(0, 4, (40, 40, 31))
    counts = [int(np.sum(o == k)) for k in range(4)] + [v for c in S.region_counts(o, t) for v in c]
    assert np.allclose(E.dice_from_counts(counts), S.softmax_output_dice(o, t), atol=1e-12)
    from utils.tools import softmax_output_dice
    assert np.allclose(softmax_output_dice(o, t), S.softmax_output_dice(o, t), atol=0)


def test_bench_helpers_roundtrip_and_configs_match():
    """bench.py host logic: the 2-bit label packing of the overlap50 goldens round-trips, both arms print the same
    `config` dict (the driver compares them), and the committed goldens decode to a full label map."""
    import argparse
    import os
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    from make_golden_overlap50 import pack2
    lab = np.random.RandomState(0).randint(0, 4, 1001).astype(np.uint8)
    assert np.array_equal(bench.unpack2(pack2(lab), lab.size), lab)
    args = argparse.Namespace(workload="overlap50", sharding="volume", gpus=1)
    assert bench.config_dict(args, 18, False) == bench.config_dict(args, 18, args.sharding == "patch" and args.gpus > 1)
    g = np.load(os.path.join(root, "tests", "golden", "overlap50_seed1000.npz"))
    full = bench.unpack2(g["labels_packed"], 240 * 240 * 155)
    assert int(g["labels_step"]) == 1 and np.array_equal(np.bincount(full, minlength=4), g["labels_hist"])
    assert g["keep_scale"].shape == (18, 16) and g["starts"].shape == (18, 3)
