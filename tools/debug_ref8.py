import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200/dropin")
import numpy as np, torch
import dcl_b200
from dcl_b200.engine import reference_starts, TOPK_TAGS
from models.clswiseformer.cls_wise_former import get_cls_wise_former
torch.manual_seed(0)
sd = get_cls_wise_former("brats", True, "fixed", 0).state_dict()
g = np.load("/root/repo/tests/golden/volume_seed1000.npz")
torch.manual_seed(1000)
vol = torch.randn(1, 4, 240, 240, 155).cuda()
a = dcl_b200.Engine(dcl_b200.Precision.F16X3, keep_stages=True); a.load_state_dict(sd)
b = dcl_b200.Engine(dcl_b200.Precision.FP32, keep_stages=True); b.load_state_dict(sd)
for i, (sx, sy, sz) in enumerate(reference_starts()):
    x = vol[..., sx:sx + 128, sy:sy + 128, sz:sz + 128]
    k = g["keep_scale"][i]
    pa = a.forward(x, k); ta = a.read_topk()
    pb = b.forward(x, k); tb = b.read_topk()
    torch.cuda.synchronize()
    nd = [t for t in TOPK_TAGS if set(ta[t].tolist()) != set(tb[t].tolist())]
    err = float((pa - pb).abs().max() / pb.abs().max())
    st = {}
    for name in ("init", "x1_1", "x2_1", "x3_1", "x4", "enc_out", "dec2"):
        sa, sb = a.read_stage(name), b.read_stage(name)
        st[name] = float((sa - sb).abs().max() / sb.abs().max())
    print(i, "keep", k.tolist().count(0.0), "zeros; err %.2e" % err, "topk diffs", nd, {n: "%.1e" % v for n, v in st.items()})
