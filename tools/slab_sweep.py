"""Sweep of the slab kernel's configuration (tiles per CTA, output channels per CTA, input-channel passes)
for one layer shape: DCL_SLAB_DEBUG=1 python tools/slab_sweep.py [x3]"""
import ctypes as C
import os
import sys

os.environ["DCL_SLAB_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dcl_b200  # noqa: E402,F401
from dcl_b200 import _native as N  # noqa: E402

lib = N.load_library()
lib.dcl_bench_conv.restype = C.c_double
lib.dcl_bench_conv.argtypes = [C.c_int32] * 6
x3 = 8 if len(sys.argv) > 1 and sys.argv[1] == "x3" else 0
for (cin, cout, g) in [(64, 64, 32), (128, 128, 16), (96, 96, 32), (128, 256, 16), (256, 384, 16)]:
    res = []
    for mt in (1, 2, 4):
        for nt in (16, 32, 64):
            for npass in (1, 2, 4):
                os.environ["DCL_SLAB_FORCE"] = f"{mt},{nt},{npass}"
                us = lib.dcl_bench_conv(cin, cout, g, 1, 7 | x3, 10)
                if us > 0:
                    res.append((round(us, 1), mt, nt, npass))
                else:
                    err = lib.dcl_last_error().decode()
                    if "does not fit" not in err:
                        print("   FAILED", (mt, nt, npass), err, flush=True)
                        try:
                            torch.cuda.synchronize()
                        except Exception as ex:  # sticky device error: nothing further can run in this process
                            print("   device error:", str(ex)[:200], flush=True)
                            sys.exit(1)
    res.sort()
    print(f"== {cin}->{cout} @{g}^3 {'x3' if x3 else 'bf16'}: best (us, mt, nt, npass):", res[:8], flush=True)
    os.environ.pop("DCL_SLAB_FORCE")
    print("   default:", lib.dcl_bench_conv(cin, cout, g, 1, 7 | x3, 10), flush=True)
