"""Host-side engine: owns one ``dcl_handle`` and marshals torch CUDA tensors (device memory and
streams only -- no torch arithmetic) into the C ABI of ``include/dcl_b200.h``."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from ._native import DclError, Precision, StitchMode

PATCH = 128
AUX_ORDER = [(head, key) for head in ("supervise", "edge", "mid_semantic", "mid_edge") for key in ("01", "02", "04")]
TOPK_TAGS = [f"{k}_{s}" for k in ("01", "02", "04") for s in ("ee", "es", "ss", "se")] + ["fusion"]


def _single(t, what):
    """(1,4,...) -> (4,...).  The C ABI takes ONE volume / patch per call; the reference's drivers use batch_size=1
    (test_overlap.py:94-97).  A larger batch is refused rather than silently truncated."""
    if t.dim() == 5:
        if t.shape[0] != 1:
            raise DclError(f"{what}: batch of {int(t.shape[0])} given, one volume per call (loop over the batch)")
        t = t[0]
    return t


def reference_starts():
    """The 8 fixed corners of predict_overlap.py:34-41 in the reference's order."""
    return [(x, y, z) for z in (0, 27) for x in (0, 112) for y in (0, 112)]


def patch_starts(shape, stride):
    """Sliding-window origins for the weighted (extension) modes: per axis every `stride` voxels plus
    the last position that still fits; z-major so a contiguous chunk of the list is a z-slab."""
    def axis(n):
        return sorted(set(range(0, n - PATCH, stride)) | {n - PATCH})
    xs, ys, zs = (axis(int(n)) for n in shape)
    return [(x, y, z) for z in zs for x in xs for y in ys]


def dice_from_counts(counts):
    """utils/tools.py:44-47 on the 13 integer counters: (2|o&t| + eps) / (|o| + |t| + eps) for WT, TC, ET."""
    c = [int(v) for v in counts]
    return [(2 * c[6 + 3 * r] + 1e-8) / (c[4 + 3 * r] + c[5 + 3 * r] + 1e-8) for r in range(3)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One handle = workspace for one 128^3 patch on the current CUDA device."""

    def __init__(self, precision=Precision.FP32, want_aux=False, keep_stages=False):
        self._lib = N.load_library()
        if not torch.cuda.is_available():
            raise DclError("dcl_b200 needs a CUDA device (there is no CPU path)")
        self.precision = Precision(precision)
        self.want_aux = bool(want_aux)
        self.keep_stages = bool(keep_stages)
        self.device = torch.device("cuda", torch.cuda.current_device())
        cfg = N.make_config(self.precision, self.want_aux, self.keep_stages)
        h = C.c_void_p()
        N.check(self._lib.dcl_create(C.byref(cfg), C.byref(h)))
        self._h = h

    def close(self):
        for ctx in getattr(self, "_shard_ctx", {}).values():      # peers' slot buffers mapped through CUDA IPC (sharded.py)
            ctx.close()
        self.__dict__.pop("_shard_ctx", None)
        if getattr(self, "_h", None):
            self._lib.dcl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        """Feeds reference state_dict tensors (CPU or CUDA, optional 'module.' prefix)."""
        known = {n for n, _, _ in N.weight_catalogue()}
        for name, t in state_dict.items():
            key = name[7:] if name.startswith("module.") else name
            if key not in known:
                if strict:
                    raise DclError(f"unexpected state_dict key {name!r}")
                continue
            t = t.detach().to(torch.float32).contiguous()
            N.check(self._lib.dcl_set_weight(self._h, key.encode(), _ptr(t), t.numel()))
        if strict:
            buf = C.create_string_buffer(256)
            missing = self._lib.dcl_missing_weights(self._h, buf, 256)
            if missing:
                raise DclError(f"{missing} state_dict tensors missing, first: {buf.value.decode()}")

    # ---- one patch -----------------------------------------------------------------------
    def forward(self, x, keep_scale=None, want_aux=False):
        """ClsWiseFormer.forward for one (1,4,128,128,128) or (4,128,128,128) CUDA fp32 view.
        Returns probs (1,4,128,128,128) and, if want_aux, the 4 dicts of the reference."""
        if x.dim() == 5:
            if x.shape[0] != 1:
                raise DclError("forward takes one patch; loop over the batch (SURVEY H6)")
            x = x[0]
        if tuple(x.shape) != (4, PATCH, PATCH, PATCH) or x.dtype != torch.float32 or not x.is_cuda:
            raise DclError("x must be a CUDA fp32 (4,128,128,128) tensor")
        if x.stride(3) != 1:
            x = x.contiguous()
        strides = (C.c_int64 * 4)(*x.stride())
        keep = None
        if keep_scale is not None:
            keep = np.ascontiguousarray(np.asarray(keep_scale, dtype=np.float32).reshape(16))
        probs = torch.empty((1, 4, PATCH, PATCH, PATCH), dtype=torch.float32, device=x.device)
        aux_t, aux_p = None, None
        if want_aux:
            if not self.want_aux:
                raise DclError("engine was created without want_aux")
            aux_t = [torch.empty((1, 2, PATCH, PATCH, PATCH), dtype=torch.float32, device=x.device) for _ in range(12)]
            aux_p = (C.c_void_p * 12)(*[t.data_ptr() for t in aux_t])
        N.check(self._lib.dcl_forward(self._h, _ptr(x), strides, keep.ctypes.data_as(C.c_void_p) if keep is not None
                                      else C.c_void_p(0), _ptr(probs), aux_p, _stream()))
        if not want_aux:
            return probs
        dicts = {head: {} for head, _ in AUX_ORDER}
        for (head, key), t in zip(AUX_ORDER, aux_t):
            dicts[head][key] = t
        return probs, dicts["supervise"], dicts["edge"], dicts["mid_semantic"], dicts["mid_edge"]

    # ---- one volume ----------------------------------------------------------------------
    @staticmethod
    def _plan_args(mode, starts, keep_scales, n_expected=None):
        mode = StitchMode(mode)
        s_arr, n = None, 0
        if mode in (StitchMode.UNIFORM, StitchMode.GAUSSIAN):
            if not starts:
                raise DclError("weighted stitch modes need a patch list (see patch_starts)")
            s_arr = np.ascontiguousarray(np.asarray(starts, dtype=np.int32).reshape(-1, 3))
            n = s_arr.shape[0]
        else:
            n = 8
        k_arr = None
        if keep_scales is not None:
            k_arr = np.ascontiguousarray(np.asarray(keep_scales, dtype=np.float32).reshape(n, 16))
        return mode, n, s_arr, k_arr

    def predict_volume(self, vol, mode=StitchMode.REFERENCE, starts=None, keep_scales=None, target=None,
                       want_probs=True, want_labels=True):
        """tailor_and_concat + arg-max + Dice counters on a CUDA (4,X,Y,Z) / (1,4,X,Y,Z) fp32 volume.
        Returns dict(probs (1,4,X,Y,Zout) | None, labels uint8 (X,Y,Zout) | None, counts int64[13] tensor)."""
        vol = _single(vol, "vol")
        if vol.dim() != 4 or vol.shape[0] != 4 or vol.dtype != torch.float32 or not vol.is_cuda:
            raise DclError("vol must be a CUDA fp32 (4,X,Y,Z) tensor")
        vol = vol.contiguous()
        mode, n, s_arr, k_arr = self._plan_args(mode, starts, keep_scales)
        X, Y, Z = (int(v) for v in vol.shape[1:])
        zout = 155 if mode in (StitchMode.REFERENCE, StitchMode.ALIGNED) else Z
        shape = (C.c_int32 * 3)(X, Y, Z)
        probs = torch.empty((1, 4, X, Y, zout), dtype=torch.float32, device=vol.device) if want_probs else None
        labels = torch.empty((X, Y, zout), dtype=torch.uint8, device=vol.device) if want_labels else None
        counts = torch.zeros(13, dtype=torch.int64, device=vol.device)
        if target is not None:
            target = target.to(device=vol.device, dtype=torch.uint8).contiguous()
            if tuple(target.shape) != (X, Y, zout):
                raise DclError(f"target must have shape {(X, Y, zout)}")
        N.check(self._lib.dcl_predict_volume(
            self._h, _ptr(vol), shape, int(mode), n,
            s_arr.ctypes.data_as(C.c_void_p) if s_arr is not None else C.c_void_p(0),
            k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0),
            _ptr(probs), _ptr(labels), _ptr(target), _ptr(counts), _stream()))
        return {"probs": probs, "labels": labels, "counts": counts}

    def predict_volume_aux(self, vol, mode=StitchMode.UNIFORM, starts=None, keep_scales=None, target=None,
                           want_probs=True, want_labels=True):
        """BASELINE config 4: the weighted sliding window with the six final auxiliary heads blended alongside
        (forward()[1] and [2]: supervise / edge x {'01','02','04'}).  Returns predict_volume's dict plus
        'supervise' and 'edge': dicts of (1,2,X,Y,Z) tensors keyed like the reference's outputs."""
        if not self.want_aux:
            raise DclError("engine was created without want_aux")
        vol = _single(vol, "vol")
        if vol.dim() != 4 or vol.shape[0] != 4 or vol.dtype != torch.float32 or not vol.is_cuda:
            raise DclError("vol must be a CUDA fp32 (4,X,Y,Z) tensor")
        vol = vol.contiguous()
        mode, n, s_arr, k_arr = self._plan_args(mode, starts, keep_scales)
        if s_arr is None:
            raise DclError("predict_volume_aux needs a weighted stitch mode and a patch list")
        X, Y, Z = (int(v) for v in vol.shape[1:])
        shape = (C.c_int32 * 3)(X, Y, Z)
        probs = torch.empty((1, 4, X, Y, Z), dtype=torch.float32, device=vol.device) if want_probs else None
        aux = torch.empty((6, 2, X, Y, Z), dtype=torch.float32, device=vol.device)
        labels = torch.empty((X, Y, Z), dtype=torch.uint8, device=vol.device) if want_labels else None
        counts = torch.zeros(13, dtype=torch.int64, device=vol.device)
        if target is not None:
            target = target.to(device=vol.device, dtype=torch.uint8).contiguous()
            if tuple(target.shape) != (X, Y, Z):
                raise DclError(f"target must have shape {(X, Y, Z)}")
        N.check(self._lib.dcl_predict_volume_aux(
            self._h, _ptr(vol), shape, int(mode), n, s_arr.ctypes.data_as(C.c_void_p),
            k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0),
            _ptr(probs), _ptr(aux), _ptr(labels), _ptr(target), _ptr(counts), _stream()))
        out = {"probs": probs, "labels": labels, "counts": counts, "supervise": {}, "edge": {}}
        for j, (head, key) in enumerate(AUX_ORDER[:6]):
            out[head][key] = aux[j][None]
        return out

    def predict_volume_tta(self, vol, keep_scales=None, target=None, want_probs=True, want_labels=True):
        """8-flip test-time augmentation around the reference tiling (predict_cls.py:180-203) on a CUDA
        (4,240,240,>=155) / (1,4,...) fp32 volume.  keep_scales: None or (8 flips, 8 patches, 16).
        Returns dict(probs (1,4,240,240,155) | None, labels uint8 (240,240,155) | None, counts int64[13])."""
        vol = _single(vol, "vol")
        if vol.dim() != 4 or vol.shape[0] != 4 or vol.dtype != torch.float32 or not vol.is_cuda:
            raise DclError("vol must be a CUDA fp32 (4,X,Y,Z) tensor")
        vol = vol.contiguous()
        X, Y, Z = (int(v) for v in vol.shape[1:])
        shape = (C.c_int32 * 3)(X, Y, Z)
        k_arr = None
        if keep_scales is not None:
            k_arr = np.ascontiguousarray(np.asarray(keep_scales, dtype=np.float32).reshape(8, 8, 16))
        probs = torch.empty((1, 4, X, Y, 155), dtype=torch.float32, device=vol.device) if want_probs else None
        labels = torch.empty((X, Y, 155), dtype=torch.uint8, device=vol.device) if want_labels else None
        counts = torch.zeros(13, dtype=torch.int64, device=vol.device)
        if target is not None:
            target = target.to(device=vol.device, dtype=torch.uint8).contiguous()
            if tuple(target.shape) != (X, Y, 155):
                raise DclError(f"target must have shape {(X, Y, 155)}")
        N.check(self._lib.dcl_predict_volume_tta(
            self._h, _ptr(vol), shape, k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0),
            _ptr(probs), _ptr(labels), _ptr(target), _ptr(counts), _stream()))
        return {"probs": probs, "labels": labels, "counts": counts}

    def predict_volume_host(self, vol_host, mode=StitchMode.REFERENCE, starts=None, keep_scales=None,
                            target_host=None, labels_out=None, probs_out=None):
        """The end-to-end call: HOST (ideally pinned) fp32 volume in, HOST uint8 labels + 13 counters out;
        both copies happen inside the library on the current stream, which is synchronised on return."""
        vol_host = _single(vol_host, "vol_host")
        if vol_host.is_cuda or vol_host.dtype != torch.float32 or vol_host.shape[0] != 4:
            raise DclError("vol_host must be a CPU fp32 (4,X,Y,Z) tensor")
        vol_host = vol_host.contiguous()
        mode, n, s_arr, k_arr = self._plan_args(mode, starts, keep_scales)
        X, Y, Z = (int(v) for v in vol_host.shape[1:])
        zout = 155 if mode in (StitchMode.REFERENCE, StitchMode.ALIGNED) else Z
        shape = (C.c_int32 * 3)(X, Y, Z)
        if labels_out is None:
            labels_out = torch.empty((X, Y, zout), dtype=torch.uint8).pin_memory()
        counts = np.zeros(13, dtype=np.uint64)
        if target_host is not None:
            target_host = target_host.to(torch.uint8).contiguous()
            if tuple(target_host.shape) != (X, Y, zout):
                raise DclError(f"target_host must have shape {(X, Y, zout)}")
        if tuple(labels_out.shape) != (X, Y, zout) or labels_out.dtype != torch.uint8 or not labels_out.is_contiguous():
            raise DclError(f"labels_out must be a contiguous uint8 tensor of shape {(X, Y, zout)}")
        N.check(self._lib.dcl_predict_volume_host(
            self._h, _ptr(vol_host), shape, int(mode), n,
            s_arr.ctypes.data_as(C.c_void_p) if s_arr is not None else C.c_void_p(0),
            k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0),
            _ptr(probs_out), _ptr(labels_out), _ptr(target_host), counts.ctypes.data_as(C.c_void_p), _stream()))
        return {"labels": labels_out, "counts": counts.astype(np.int64), "probs": probs_out}

    # ---- multi-GPU building blocks -----------------------------------------------
    def accumulate_patches(self, vol, mode, starts, keep_scales, first, count, acc, wsum):
        vol = _single(vol, "vol")
        vol = vol.contiguous()
        mode, n, s_arr, k_arr = self._plan_args(mode, starts, keep_scales)
        shape = (C.c_int32 * 3)(*[int(v) for v in vol.shape[1:]])
        N.check(self._lib.dcl_accumulate_patches(
            self._h, _ptr(vol), shape, int(mode), n,
            s_arr.ctypes.data_as(C.c_void_p) if s_arr is not None else C.c_void_p(0),
            k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0),
            int(first), int(count), _ptr(acc), _ptr(wsum), _stream()))

    def finalize_labels(self, acc, wsum, v0, nvox, labels, target=None, counts=None, probs_out=None):
        total = acc.numel() // 4
        N.check(self._lib.dcl_finalize_labels(_ptr(acc), _ptr(wsum), total, int(v0), int(nvox), _ptr(probs_out),
                                              _ptr(labels), _ptr(target), _ptr(counts), _stream()))

    # ---- multi-GPU, owner-computes form (sharded.py) ----------------------------------------------
    def slots_ensure(self, n_slots):
        """Device address of this handle's exportable slot buffer, sized for n_slots patches (4 x 128^3 fp32 each)."""
        p = C.c_void_p()
        N.check(self._lib.dcl_slots_ensure(self._h, int(n_slots), C.byref(p)))
        return int(p.value)

    def ipc_export(self, dev_ptr):
        buf = C.create_string_buffer(64)
        N.check(self._lib.dcl_ipc_export(C.c_void_p(int(dev_ptr)), buf))
        return buf.raw

    def ipc_import(self, handle):
        p = C.c_void_p()
        N.check(self._lib.dcl_ipc_import(C.create_string_buffer(bytes(handle), 64), C.byref(p)))
        return int(p.value)

    def ipc_release(self, dev_ptr):
        N.check(self._lib.dcl_ipc_release(C.c_void_p(int(dev_ptr))))

    def forward_patches_to_slots(self, vol, mode, starts, keep_scales, first, count):
        vol = _single(vol, "vol").contiguous()
        mode, n, s_arr, k_arr = self._plan_args(mode, starts, keep_scales)
        shape = (C.c_int32 * 3)(*[int(v) for v in vol.shape[1:]])
        N.check(self._lib.dcl_forward_patches_to_slots(
            self._h, _ptr(vol), shape, int(mode), n, s_arr.ctypes.data_as(C.c_void_p),
            k_arr.ctypes.data_as(C.c_void_p) if k_arr is not None else C.c_void_p(0), int(first), int(count), _stream()))

    def gather_finalize_range(self, shape, mode, starts, slot_ptrs, x0, x1, labels, target=None, counts=None, probs_out=None):
        """Blend + normalise + arg-max + counters of the rows x0 <= x < x1 from one slot pointer per patch (local or
        IPC-mapped peer memory).  labels / target / probs_out are whole-volume tensors."""
        mode, n, s_arr, _ = self._plan_args(mode, starts, None)
        ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in slot_ptrs])
        shp = (C.c_int32 * 3)(*[int(v) for v in shape])
        N.check(self._lib.dcl_gather_finalize_range(shp, int(mode), n, s_arr.ctypes.data_as(C.c_void_p), ptrs, int(x0), int(x1),
                                                    _ptr(probs_out), _ptr(labels), _ptr(target), _ptr(counts), _stream()))

    # ---- introspection -------------------------------------------------------------------
    def read_stage(self, name):
        n = N.check(self._lib.dcl_read_stage(self._h, name.encode(), C.c_void_p(0), 0, _stream()))
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        N.check(self._lib.dcl_read_stage(self._h, name.encode(), _ptr(out), n, _stream()))
        return out

    def read_topk(self):
        out = np.zeros((13, 128), dtype=np.int32)
        N.check(self._lib.dcl_read_topk(self._h, out.ctypes.data_as(C.c_void_p), _stream()))
        return dict(zip(TOPK_TAGS, out))

    def profile(self, on=True):
        N.check(self._lib.dcl_profile_enable(self._h, int(bool(on))))

    def profile_read(self, cls):
        """(total ms, launches, total work) of kernel class `cls` (0: 3x3x3 convs / flops, 1: stitch tail / bytes)
        since the last read, from CUDA events recorded around every launch on the launching stream."""
        ms, n, work = C.c_double(), C.c_int64(), C.c_double()
        N.check(self._lib.dcl_profile_read(self._h, int(cls), C.byref(ms), C.byref(n), C.byref(work)))
        return ms.value, int(n.value), work.value

    @property
    def launch_count(self):
        return int(self._lib.dcl_launch_count(self._h))


def op_conv3d_k3(x0, weight, bias=None, x1=None, stride=1, norm=None, act=0, residual=None, impl=0, stats=None):
    """Single-operator entry (tests): y = conv3d(act(norm(cat(x0, x1))), k=3, p=1) + residual."""
    lib = N.load_library()
    c0, d, h, w = (int(v) for v in x0.shape)
    c1 = int(x1.shape[0]) if x1 is not None else 0
    cout = int(weight.shape[0])
    od, oh, ow = ((v - 1) // stride + 1 for v in (d, h, w))
    y = torch.empty((cout, od, oh, ow), dtype=torch.float32, device=x0.device)
    dims = (C.c_int32 * 3)(d, h, w)
    mean, rstd = (norm if norm is not None else (None, None))
    N.check(lib.dcl_op_conv3d_k3(_ptr(x0.contiguous()), c0, _ptr(x1), c1, dims, _ptr(weight.contiguous()), _ptr(bias),
                                 cout, int(stride), _ptr(mean), _ptr(rstd), int(act), _ptr(residual), _ptr(y),
                                 int(impl), _ptr(stats), _stream()))
    return y


def op_instnorm_stats(x):
    lib = N.load_library()
    c = int(x.shape[0])
    spatial = x.numel() // c
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(c, dtype=torch.float32, device=x.device)
    N.check(lib.dcl_op_instnorm_stats(_ptr(x.contiguous()), c, spatial, _ptr(mean), _ptr(rstd), _stream()))
    return mean, rstd
