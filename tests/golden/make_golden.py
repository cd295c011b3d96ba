"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
from /root/reference on CPU (fp32).  Runs only in the build container (the GPU box has
no /root/reference); the produced ``*.npz`` / ``*.json`` files are committed.

    python tests/golden/make_golden.py            # ~1 min on 8 cores

Recipe: stub the unused imports (nibabel, imageio, medpy), create the
missing ``fix_index.txt`` in a temp cwd, seed 0 for the weights, seed 1 / 1000+i for the
inputs, seed 2000+i immediately before a forward so the always-on dropout3d draw
(Unet_skipconnection.py:31) is replayable.
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
N_SAMPLE = 4096


def load_reference():
    for name in ("nibabel", "imageio", "medpy", "medpy.metric", "setproctitle"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["medpy"].metric = sys.modules["medpy.metric"]
    sys.path.insert(0, REF)
    tmp = tempfile.mkdtemp(prefix="dcl_ref_")
    d = os.path.join(tmp, "2-MICCAI_BraTS_2018", "MICCAI_BraTS_2018_Data_Training")
    os.makedirs(d)
    with open(os.path.join(d, "fix_index.txt"), "w") as f:
        f.write(repr({str(k): [k] * 512 for k in range(2048)}))
    os.chdir(tmp)
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    import predict_overlap
    import utils.tools as tools
    return get_cls_wise_former, predict_overlap, tools


def digest(t):
    """Shape, moments and a fixed strided sample of a tensor."""
    a = t.detach().to(torch.float64).flatten()
    n = a.numel()
    step = max(1, n // N_SAMPLE)
    return {
        "shape": np.array(t.shape, dtype=np.int64),
        "mean": np.float64(a.mean()), "std": np.float64(a.std(unbiased=False)),
        "absmax": np.float64(a.abs().max()), "l2": np.float64(a.norm()),
        "sample": t.detach().flatten()[::step][:N_SAMPLE].to(torch.float32).numpy().copy(),
    }


def put(store, name, t):
    for k, v in digest(t).items():
        store[f"{name}/{k}"] = v


def main():
    torch.set_num_threads(os.cpu_count())
    get_model, predict_overlap, tools = load_reference()
    torch.manual_seed(0)
    model = get_model("brats", True, "fixed", 0).eval()
    sd = model.state_dict()
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in sd.items()}, f, indent=0)
    # seed-0 initialisation digest: lets the drop-in module prove it consumes the RNG like the reference
    with open(os.path.join(HERE, "state_dict_seed0_digest.json"), "w") as f:
        json.dump({k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}, f, indent=0)
    if "--weights-only" in sys.argv:
        return

    # ------------------------------------------------------------------ config 1: one patch
    store = {}
    stages = {}

    def hook(name):
        def fn(_m, _i, out):
            stages[name] = (out if torch.is_tensor(out) else out[0]).detach().clone()
        return fn

    def unet_hook(_m, _i, out):
        for n, t in zip(("x1_1", "x2_1", "x3_1", "x4"), out):
            stages[n] = t.detach().clone()

    hs = [model.Unet_list.register_forward_hook(unet_hook),
          model.Unet_list.InitConv.register_forward_hook(hook("init"))]
    for r, e, s in (("1", "relu_edge", "relu_list"), ("2", "relu_edge_2", "relu_list_2"),
                    ("4", "relu_edge_4", "relu_list_4")):
        hs.append(getattr(model, e).register_forward_hook(hook("edge_" + r)))
        hs.append(getattr(model, s).register_forward_hook(hook("sem_" + r)))
    for key in ("01", "02", "04"):
        hs.append(getattr(model, "transformer_" + key).register_forward_hook(hook("coupler_" + key)))
    hs.append(model.fusion_transformer_1_2_4.register_forward_hook(hook("coupler_fusion")))
    hs.append(model.sum_fusion.register_forward_hook(hook("enc_out")))
    for n, m in (("dec8", "Enblock8_2"), ("dec4", "DeBlock4_1"), ("dec3", "DeBlock3_1"), ("dec2", "DeBlock2_1")):
        hs.append(getattr(model.decoder, m).register_forward_hook(hook(n)))

    topk_log = []
    orig_topk = torch.Tensor.topk

    def logged_topk(self, *a, **k):
        out = orig_topk(self, *a, **k)
        topk_log.append(out[1][0, 0].clone())
        return out

    torch.manual_seed(1)
    x = torch.randn(1, 4, 128, 128, 128)
    torch.manual_seed(2000)
    keep = torch.empty(1, 16, 1, 1, 1).bernoulli_(0.8).div_(0.8).reshape(1, 16)
    torch.manual_seed(2000)
    torch.Tensor.topk = logged_topk
    try:
        with torch.no_grad():
            probs, sup, edge, mid_sem, mid_edge = model(x, None)
    finally:
        torch.Tensor.topk = orig_topk
    for h in hs:
        h.remove()
    assert len(topk_log) == 13
    tags = [f"{k}_{s}" for k in ("01", "02", "04") for s in ("ee", "es", "ss", "se")] + ["fusion"]
    store["keep_scale"] = keep.numpy()
    for tag, idx in zip(tags, topk_log):
        store["topk_" + tag] = idx.numpy().astype(np.int32)
    for n, t in stages.items():
        put(store, n, t)
    put(store, "probs", probs)
    for nm, dct in (("sup", sup), ("edgeout", edge), ("mid_sem", mid_sem), ("mid_edge", mid_edge)):
        for key, t in dct.items():
            put(store, f"{nm}_{key}", t)
    lab = probs[0].numpy().argmax(0).astype(np.uint8)
    store["labels_sha256"] = np.frombuffer(hashlib.sha256(lab.tobytes()).digest(), dtype=np.uint8)
    store["labels_hist"] = np.bincount(lab.ravel(), minlength=4).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "patch_seed1.npz"), **store)
    print("patch golden written:", len(store), "arrays")

    # ------------------------------------------------------------------ config 2 (reference tiling): one volume
    vol = {}
    i = 0
    torch.manual_seed(1000 + i)
    xv = torch.randn(1, 4, 240, 240, 155)
    torch.manual_seed(2000 + i)
    keeps = [torch.empty(1, 16, 1, 1, 1).bernoulli_(0.8).div_(0.8).reshape(16).numpy() for _ in range(8)]
    torch.manual_seed(2000 + i)
    with torch.no_grad():
        out = predict_overlap.tailor_and_concat(xv, None, model)
    output = out[0, :, :240, :240, :160].numpy()
    labels = output.argmax(0)
    target = np.random.RandomState(i).randint(0, 4, (240, 240, 155))
    dice = tools.softmax_output_dice(labels, target)
    vol["keep_scale"] = np.stack(keeps)
    put(vol, "stitched", out)
    vol["labels_sha256"] = np.frombuffer(hashlib.sha256(labels.astype(np.uint8).tobytes()).digest(), dtype=np.uint8)
    vol["labels_hist"] = np.array([np.sum(labels == k) for k in range(4)], dtype=np.int64)
    vol["labels_sample"] = labels.ravel()[:: labels.size // N_SAMPLE][:N_SAMPLE].astype(np.uint8)
    vol["dice"] = np.array(dice, dtype=np.float64)
    # margin between the two largest class probabilities, to size the label-flip budget
    srt = np.sort(output, axis=0)
    margin = srt[-1] - srt[-2]
    vol["margin_quantiles"] = np.quantile(margin, [1e-5, 1e-4, 1e-3, 1e-2]).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "volume_seed1000.npz"), **vol)
    print("volume golden written; dice", dice, "hist", vol["labels_hist"])


if __name__ == "__main__":
    main()
