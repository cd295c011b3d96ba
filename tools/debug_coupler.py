import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200/dropin")
import numpy as np, torch
import dcl_b200
from models.clswiseformer.cls_wise_former import get_cls_wise_former
torch.manual_seed(0)
sd = get_cls_wise_former("brats", True, "fixed", 0).state_dict()
torch.manual_seed(1)
x = torch.randn(1, 4, 128, 128, 128).cuda()
g = np.load("/root/repo/tests/golden/patch_seed1.npz")
outs = {}
for name in ("FP32", "BF16X3", "BF16"):
    eng = dcl_b200.Engine(dcl_b200.Precision[name], want_aux=True, keep_stages=True)
    eng.load_state_dict(sd)
    eng.forward(x, g["keep_scale"], want_aux=True)
    torch.cuda.synchronize()
    outs[name] = {k: eng.read_stage(k).cpu().numpy() for k in ("coupler_01", "coupler_02", "coupler_04", "coupler_fusion", "enc_out")}
    eng.close()
for k in outs["FP32"]:
    a = outs["FP32"][k]
    for m in ("BF16X3", "BF16"):
        b = outs[m][k]
        if k.startswith("coupler"):
            A = a.reshape(-1, 512); B = b.reshape(-1, 512)
            rowerr = np.abs(A - B).max(1) / np.abs(A).max()
            bad = np.nonzero(rowerr > 1e-3)[0]
            print(k, m, "max", rowerr.max(), "bad rows", len(bad), bad[:10], bad[-5:])
        else:
            print(k, m, np.abs(a - b).max() / np.abs(a).max())
