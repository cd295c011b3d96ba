"""In-kernel timeline of the LAST slab-kernel launch of a full bf16 forward (tags 10-15 survive later kernels)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200", "dropin"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dcl_b200  # noqa: E402
from dcl_b200 import _native as N  # noqa: E402
from models.clswiseformer.cls_wise_former import get_cls_wise_former  # noqa: E402

lib = N.load_library()
torch.manual_seed(0)
sd = get_cls_wise_former("brats", True, "fixed", 0).state_dict()
torch.manual_seed(1)
x = torch.randn(1, 4, 128, 128, 128).cuda()
eng = dcl_b200.Engine(dcl_b200.Precision.BF16)
eng.load_state_dict(sd)
for _ in range(3):
    eng.forward(x, None)
torch.cuda.synchronize()
N.check(lib.dcl_trace_enable(1))
eng.forward(x, None)
torch.cuda.synchronize()
buf = np.zeros(2 * 8192, dtype=np.int64)
n = lib.dcl_trace_read(buf.ctypes.data_as(C.c_void_p), 8192)
recs = [(int(buf[2 * i]) >> 32, int(buf[2 * i]) & 0xffffffff, int(buf[2 * i + 1])) for i in range(n)]
names = {10: "wgt: stage free", 11: "slab: copies issued", 12: "slab: landed", 13: "slab: published", 14: "mma: weights landed",
         15: "epi: acc complete", 16: "all roles done"}
sl = sorted([r for r in recs if r[0] >= 10], key=lambda r: r[2])
t0 = sl[0][2]
for tag, step, t in sl:
    print(f"{t - t0:9d}  {names.get(tag, tag):22s} {step}")
N.check(lib.dcl_trace_enable(0))
