"""ctypes binding of include/dcl_b200.h (one Python function per exported symbol)."""
from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libdcl_b200.so"
_lib = None


class DclError(RuntimeError):
    """A dcl_* call returned a negative status (message from dcl_last_error)."""


class Precision(enum.IntEnum):
    FP32 = 0
    F16X3 = 1       # split operands: fp16 hi + fp16 lo, 3-4 tcgen05 MMAs per product (the parity-grade tensor-core mode)
    BF16X3 = 1      # the name the mode was specified under; same value
    BF16 = 2


class StitchMode(enum.IntEnum):
    REFERENCE = 0
    ALIGNED = 1
    UNIFORM = 2
    GAUSSIAN = 3


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32), ("want_aux", C.c_int32),
                ("keep_stages", C.c_int32), ("reserved", C.c_int32 * 12)]


ABI_VERSION = 1


def library_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


def build_library(verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise DclError("building libdcl_b200.so failed")
    return library_path()


_SIGNATURES = {
    "dcl_last_error": (C.c_char_p, []),
    "dcl_abi_version": (C.c_int, []),
    "dcl_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "dcl_destroy": (C.c_int, [C.c_void_p]),
    "dcl_workspace_bytes": (C.c_int64, [C.POINTER(Config)]),
    "dcl_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "dcl_weight_count": (C.c_int, []),
    "dcl_weight_spec": (C.c_int, [C.c_int, C.c_char_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "dcl_missing_weights": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "dcl_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_void_p), C.c_void_p]),
    "dcl_predict_volume": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_predict_volume_aux": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "dcl_predict_volume_host": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "dcl_predict_volume_tta": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_accumulate_patches": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "dcl_slots_ensure": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "dcl_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcl_ipc_import": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dcl_ipc_release": (C.c_int, [C.c_void_p]),
    "dcl_forward_patches_to_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "dcl_gather_finalize_range": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p),
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_finalize_labels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_export_labels": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_snapshot_frames": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_slice_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    "dcl_write_nifti": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "dcl_write_npy_labels": (C.c_int, [C.c_char_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "dcl_write_png_rgb": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]),
    "dcl_read_nifti_header": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "dcl_read_nifti_f32": (C.c_int64, [C.c_char_p, C.c_void_p, C.c_int64]),
    "dcl_preprocess_volume": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_reorder_labels": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "dcl_hausdorff_workspace_bytes": (C.c_int64, [C.POINTER(C.c_int32)]),
    "dcl_hausdorff": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_int64, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_void_p]),
    "dcl_percentile_from_hist": (C.c_double, [C.c_void_p, C.c_int64, C.c_double]),
    "dcl_read_stage": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dcl_read_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_launch_count": (C.c_int64, [C.c_void_p]),
    "dcl_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "dcl_profile_read": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_double)]),
    "dcl_op_conv3d_k3": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dcl_op_instnorm_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dcl_bench_conv": (C.c_double, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dcl_bench_stitch": (C.c_double, [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "dcl_debug_stamps": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dcl_trace_enable": (C.c_int, [C.c_int32]),
    "dcl_trace_read": (C.c_int64, [C.c_void_p, C.c_int64]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library():
    """Loads libdcl_b200.so; raises DclError when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise DclError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(dcl_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dcl_abi_version() != ABI_VERSION:
        raise DclError("libdcl_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        raise DclError(load_library().dcl_last_error().decode() or f"dcl status {rc}")
    return rc


def abi_version() -> int:
    return load_library().dcl_abi_version()


def make_config(precision=Precision.FP32, want_aux=False, keep_stages=False) -> Config:
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.precision = int(precision)
    cfg.want_aux = int(bool(want_aux))
    cfg.keep_stages = int(bool(keep_stages))
    return cfg


def workspace_bytes(precision=Precision.FP32, want_aux=False) -> int:
    cfg = make_config(precision, want_aux)
    return int(load_library().dcl_workspace_bytes(C.byref(cfg)))


def weight_catalogue():
    """[(state_dict key, numel, aux_only)] in the reference's registration order."""
    lib = load_library()
    out = []
    buf = C.create_string_buffer(256)
    for i in range(lib.dcl_weight_count()):
        numel, aux = C.c_int64(), C.c_int32()
        check(lib.dcl_weight_spec(i, buf, 256, C.byref(numel), C.byref(aux)))
        out.append((buf.value.decode(), int(numel.value), bool(aux.value)))
    return out
