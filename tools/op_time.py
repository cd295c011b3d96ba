"""CUDA-event timing of single tcgen05 convolution kernels launched back to back (no other kernels in between)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dcl_b200  # noqa: E402
from dcl_b200 import _native as N  # noqa: E402

lib = N.load_library()
lib.dcl_bench_conv.restype = C.c_double
lib.dcl_bench_conv.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
for (cin, cout, g, stride) in [(16, 16, 128, 1), (32, 32, 64, 1), (64, 64, 32, 1), (128, 128, 16, 1), (96, 96, 32, 1), (256, 384, 16, 1),
                               (128, 256, 16, 1), (16, 32, 128, 2), (32, 64, 64, 2), (64, 128, 32, 2), (32, 32, 64, 2)]:
    x3 = 8 if len(sys.argv) > 1 and sys.argv[1] == "x3" else 0      # python tools/op_time.py x3: the split-fp16 kernels
    for mode in ((0, 1, 2, 4, 7) if stride == 1 and g >= 64 and not x3 else (0, 7)):
        mode |= x3
        us = lib.dcl_bench_conv(cin, cout, g, stride, mode, 20)
        og = (g - 1) // stride + 1
        fl = 2.0 * 27 * cin * cout * og ** 3
        print(f"conv {cin:3d}->{cout:3d} @{g:3d}^3 s{stride} {'+'.join(n for b, n in ((1, 'norm'), (2, 'res'), (4, 'stats')) if mode & b) or 'plain':14s}{' x3' if x3 else ''}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
