"""Runs N forwards of one 128^3 patch through dcl_forward (for ncu launch lists / quick timing).
usage: python tools/one_patch.py [precision: fp32|f16x3|bf16x3|bf16] [n_forwards]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200", "dropin"))

import torch  # noqa: E402
import dcl_b200  # noqa: E402
from models.clswiseformer.cls_wise_former import get_cls_wise_former  # noqa: E402


def main():
    prec = {"fp32": dcl_b200.Precision.FP32, "bf16x3": dcl_b200.Precision.F16X3, "f16x3": dcl_b200.Precision.F16X3, "bf16": dcl_b200.Precision.BF16}[
        sys.argv[1] if len(sys.argv) > 1 else "fp32"]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.manual_seed(0)
    sd = get_cls_wise_former("brats", True, "fixed", 0).state_dict()
    torch.manual_seed(1)
    x = torch.randn(1, 4, 128, 128, 128).cuda()
    eng = dcl_b200.Engine(prec)
    eng.load_state_dict(sd)
    eng.forward(x, None)
    torch.cuda.synchronize()
    l0 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(n):
        p = eng.forward(x, None)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"precision={prec.name} forwards={n} ms_per_patch={dt * 1e3:.3f} launches_per_patch={(eng.launch_count - l0) // n} "
          f"checksum={float(p.double().sum()):.6f}")
    if os.environ.get("DCL_STAMPS"):
        import numpy as np
        import ctypes as C
        from dcl_b200 import _native as N
        st = np.zeros(16, dtype=np.uint64)
        N.check(N.load_library().dcl_debug_stamps(eng._h, st.ctypes.data_as(C.c_void_p)))
        names = ["encoder", "decoupler+tokenise", "region couplers", "cross-region coupler + sum_fusion...", "decoder 16^3",
                 "decoder 32^3", "decoder 64^3", "decoder 128^3", "endconv"]
        for i, nm in enumerate(names):
            print(f"  stage {nm:38s} {(int(st[i + 1]) - int(st[i])) / 1e3:8.1f} us")
        print(f"  total {(int(st[9]) - int(st[0])) / 1e3:8.1f} us")
        if st[10]:
            for a, b, nm in ((2, 10, "region 0 lane A: 2 selects"), (10, 11, "A1 block"), (11, 12, "wait A2 + A3 block"), (12, 13, "wait A4 + FFN")):
                print(f"  coupler {nm:30s} {(int(st[b]) - int(st[a])) / 1e3:8.1f} us")
    if len(sys.argv) > 3 and sys.argv[3] == "prof":
        names = {0: "all k3 convs", 2: "roll16@128", 3: "roll32@64", 4: "slab", 5: "gemm conv", 6: "deup", 7: "norm_act_b",
                 8: "token path", 9: "endconv", 10: "tokenise", 11: "1x1 conv"}
        eng.profile(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            eng.forward(x, None)
        ev1.record()
        torch.cuda.synchronize()
        eng.profile(False)
        print(f"profiled pass: {ev0.elapsed_time(ev1) / n * 1e3:.0f} us per patch")
        for cls, name in names.items():
            ms, cnt, work = eng.profile_read(cls)
            if cnt:
                extra = f"  {work / (ms * 1e-3) / 1e12:7.1f} TFLOP/s" if work > 0 else ""
                print(f"  {name:14s} {ms / n * 1e3:8.1f} us per patch  {cnt // n:3d} launches  {ms / cnt * 1e3:7.1f} us each{extra}")
    eng.close()


if __name__ == "__main__":
    main()
