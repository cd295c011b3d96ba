"""In-kernel timelines (SM clocks of CTA 0) of the tcgen05 kernels for single-operator cases.
usage: python tools/trace_kernel.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dcl_b200  # noqa: E402
from dcl_b200 import _native as N  # noqa: E402
from dcl_b200.engine import op_conv3d_k3  # noqa: E402

lib = N.load_library()


def read_trace():
    buf = np.zeros(2 * 8192, dtype=np.int64)
    n = lib.dcl_trace_read(buf.ctypes.data_as(C.c_void_p), 8192)
    recs = [(int(buf[2 * i]) >> 32, int(buf[2 * i]) & 0xffffffff, int(buf[2 * i + 1])) for i in range(n)]
    return recs


def show(title, recs, names):
    if not recs:
        print(title, ": no records")
        return
    ext = {r[0]: r[2] for r in recs if r[0] >= 30}
    recs = [r for r in recs if r[0] < 30]
    if 30 in ext and 31 in ext:
        print(f"{title}: grid extent (first CTA entry -> last CTA exit) {(ext[31] + ext[30] - (1 << 62)) / 1e3:.2f} us")
    t0 = min(r[2] for r in recs)
    print(f"== {title}: {len(recs)} records, span {max(r[2] for r in recs) - t0} clk")
    for tag, step, t in sorted(recs, key=lambda r: r[2]):
        print(f"  {t - t0:9d}  {str(names.get(tag, tag)):28s} {step}")


IMPL = 1 if os.environ.get("TRACE_X3") else 2      # TRACE_X3=1: the split-fp16 kernels


def conv_case(c, g, cout=None, stride=1, norm=True):
    cout = cout or c
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(c, g, g, g, generator=gen).cuda()
    w = (torch.randn(cout, c, 3, 3, 3, generator=gen) / (c * 27) ** 0.5).cuda()
    b = torch.randn(cout, generator=gen).cuda()
    mean = torch.zeros(c).cuda()
    rstd = torch.ones(c).cuda()
    for _ in range(2):
        op_conv3d_k3(x, w, b, None, stride, (mean, rstd) if norm else None, 1 if norm else 0, None, impl=IMPL)
    torch.cuda.synchronize()
    read_trace()
    op_conv3d_k3(x, w, b, None, stride, (mean, rstd) if norm else None, 1 if norm else 0, None, impl=IMPL)
    torch.cuda.synchronize()
    return read_trace()


K1 = {0: "entry", 1: "setup done", 2: "prod: slot free, stage", 3: "prod: staged", 4: "mma: issue step", 5: "mma: issued",
      6: "epi: acc complete", 7: "epi: all stored", 8: "mma: weights landed", 9: "prod: plane landed",
      10: "prod: burst issued", 11: "epi: item stored"}
K2 = {0: "entry", 1: "setup done", 2: "prod: stage issued", 3: "mma: stage landed", 4: "epi: acc complete", 5: "epi: done",
      10: "wgt: stage free, issue tap", 11: "slab: cp.async issued", 12: "slab: landed", 13: "slab: published",
      14: "mma: tap weights landed", 15: "epi: acc complete", 16: "all roles done"}

if __name__ == "__main__":
    cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    N.check(lib.dcl_trace_enable(1 + cta))
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "roll"):
        show("roll 16->16 128^3", conv_case(16, 128, norm=True), K1)
        show("roll 32->32 64^3", conv_case(32, 64, norm=True), K1)
    if which == "roll":
        sys.exit(0)
    show("slab 64->64 32^3", conv_case(64, 32, norm=True), K2)
    show("slab 128->128 16^3", conv_case(128, 16, norm=False), K2)
    N.check(lib.dcl_trace_enable(0))
