#!/usr/bin/env python
"""Benchmark of the sliding-window ClsWiseFormer inference path (BASELINE.json metric: BraTS
4x240x240x155 volumes/sec).

    python bench.py [--gpus N --steps K --warmup W] [--precision fp32|bf16x3|bf16] [--workload overlap50|reference8|overlap75]
    python bench.py --impl reference ...      # the CPU oracle port of the reference path on the host cores

A step = one full volume: patch gather -> clswiseformer forward per 128^3 patch -> overlap
accumulate / stitch -> normalise + arg-max + label histogram + Dice counters.  Default workload is
BASELINE.json configs[1]: 50 % overlap (stride 64 -> 18 patches, uniform blend).  Prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # see dcl_b200/__init__.py; before the CUDA context exists
PKG = os.path.join(ROOT, "decouple-and-couple_learning_in_multi-modal_brain_tumor_segmentation_b200")
for p in (ROOT, os.path.join(PKG, "dropin")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SHAPE = (240, 240, 155)
VOXELS = SHAPE[0] * SHAPE[1] * SHAPE[2]
FLOPS_PER_PATCH = 523.35e9 + 0.44e9      # conv+linear without aux heads + attention
WORKLOADS = {"overlap50": ("UNIFORM", 64), "overlap75": ("UNIFORM", 32), "reference8": ("REFERENCE", None),
             "tta8": ("TTA", None)}


def seed0_weights():
    from models.clswiseformer.cls_wise_former import get_cls_wise_former
    torch.manual_seed(0)
    return get_cls_wise_former("brats", True, "fixed", 0).state_dict()


def synth_volume(i):
    torch.manual_seed(1000 + i)
    return torch.randn(1, 4, *SHAPE)


def synth_target(i):
    return torch.from_numpy(np.random.RandomState(i).randint(0, 4, SHAPE).astype(np.uint8))


def keep_scales(i, n):
    """Replay of the reference's always-on dropout3d draw, one per patch forward (seed 2000+i)."""
    g = torch.Generator().manual_seed(2000 + i)
    return torch.stack([torch.empty(1, 16, 1, 1, 1).bernoulli_(0.8, generator=g).div_(0.8).reshape(16)
                        for _ in range(n)]).numpy()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms DURING the timed region, in-process through NVML
    (nvidia-ml-py; the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons...` prints).  A polling
    `nvidia-smi -lms` child process was used first: roughly one run in four one of its queries stalled for ~80 ms and the
    device-resident loop with it (33-37 instead of 25.4 ms per volume; tools/host_enqueue.py, which has no sampler, never
    showed it).  NVML is initialised before the warm-up; only samples taken inside the timed region are reported.
    Falls back to one `nvidia-smi` query right after the region when the NVML module is unavailable."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.nvml, self.handle = index, None, None
        self.samples, self.thread, self.running = [], None, False

    def start(self):
        try:
            if os.environ.get("BENCH_NO_CLOCKS"):       # diagnostic switch: no sampling at all
                raise RuntimeError("sampling disabled")
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self._read()          # first call pays the lazy initialisation
        except Exception:
            self.nvml = None

    def _read(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        return float(sm), float(mx), int(mask)

    def _loop(self):
        while self.running:
            try:
                self.samples.append(self._read())
            except Exception:
                pass
            time.sleep(0.02)

    def region_begin(self):
        if self.nvml is not None:
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def region_end(self):
        if self.thread is not None:
            self.running = False
            self.thread.join()
            try:
                self.samples.append(self._read())
            except Exception:
                pass

    def stop(self):
        if self.nvml is None:
            return self._nvidia_smi_once()
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[2]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(name for bit, name in self.REASONS if mask & bit), "samples": len(sm),
                "source": "NVML, 20 ms period, inside the timed region"}

    def _nvidia_smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=20).stdout
            f = [v.strip() for v in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi, one query right after the timed region (NVML module unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def workload_plan(name):
    from dcl_b200 import StitchMode, patch_starts
    mode_name, stride = WORKLOADS[name]
    if mode_name == "TTA":
        return "TTA", None, 64
    if stride is None:
        return StitchMode.REFERENCE, None, 8
    starts = patch_starts(SHAPE, stride)
    return StitchMode[mode_name], starts, len(starts)


# ------------------------------------------------------------------------------------------------
def cpu_reference_volumes_per_s(workload, n_sample_patches, threads, steps=1, warmup=0):
    """The oracle port of the reference path on the host cores: `n_sample_patches` real patch forwards
    per step (extrapolated to the workload's patch count) + the full-size stitch / arg-max / Dice tail."""
    from oracle import clswiseformer_oracle as O
    from oracle import stitch_oracle as S
    torch.set_num_threads(threads)
    sd = seed0_weights()
    mode_name, stride = WORKLOADS[workload]
    if mode_name == "TTA":
        raise SystemExit("the CPU arm of the tta8 workload is 8 x the reference8 workload; run --workload reference8")
    starts = S.REFERENCE_STARTS if stride is None else S.patch_starts(SHAPE, stride)
    x = synth_volume(0)
    tgt = synth_target(0).numpy()
    ks = keep_scales(0, len(starts))
    per_step = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        probs = []
        for j in range(n_sample_patches):
            sx, sy, sz = starts[j]
            probs.append(O.forward(sd, x[..., sx:sx + 128, sy:sy + 128, sz:sz + 128],
                                   torch.from_numpy(ks[j:j + 1]), want_aux=False)[0][0].numpy())
        t_patch = (time.perf_counter() - t0) / n_sample_patches
        t1 = time.perf_counter()
        full = [probs[j % n_sample_patches] for j in range(len(starts))]
        if stride is None:
            out = S.stitch_reference_from_probs(full)
        else:
            out = S.accumulate_from_probs(full, starts, "uniform")
        labels = S.labels_from_probs(out)
        S.label_histogram(labels)
        S.softmax_output_dice(labels, tgt)
        t_tail = time.perf_counter() - t1
        if it >= warmup:
            per_step.append(t_patch * len(starts) + t_tail)
    sec = statistics.mean(per_step)
    return 1.0 / sec, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = 2
    t0 = time.perf_counter()
    vps, sec = cpu_reference_volumes_per_s(args.workload, n_sample, threads, steps=args.steps, warmup=args.warmup)
    _, _, n_patches = (None, None, 8) if WORKLOADS[args.workload][1] is None else (None, None, None)
    from oracle import stitch_oracle as S
    stride = WORKLOADS[args.workload][1]
    n_patches = 8 if stride is None else len(S.patch_starts(SHAPE, stride))
    sample = (f"{n_sample} of {n_patches} patch forwards per step on {threads} host threads (torch CPU fp32 oracle "
              f"port of predict_overlap.py + clswiseformer), extrapolated x{n_patches / n_sample:g}, plus the "
              f"full-volume stitch/argmax/Dice tail")
    line = {
        "impl": "reference", "metric": "BraTS 4x240x240x155 volumes/sec", "value": vps, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "patches_per_volume": n_patches,
                   "weights": "random-init seed 0"},
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_name(w):
    return {"overlap50": "predict_overlap sliding window, one 4x240x240x155 volume, 128^3 patches at 50% overlap "
                         "(stride 64, 18 patches, uniform blend)",
            "overlap75": "4x240x240x155 volume, 128^3 patches at 75% overlap (stride 32, 50 patches, uniform blend)",
            "reference8": "predict_overlap.py 8-corner tiling + crop-overwrite stitch, one 4x240x240x155 volume",
            "tta8": "predict_cls.py 8-flip TTA around the 8-corner tiling (64 patch forwards), one 4x240x240x155 volume"}[w]


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import dcl_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle port)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prec = {"fp32": dcl_b200.Precision.FP32, "bf16x3": dcl_b200.Precision.BF16X3, "bf16": dcl_b200.Precision.BF16}[
        args.precision]
    eng = dcl_b200.Engine(prec)
    eng.load_state_dict(seed0_weights())
    mode, starts, n_patches = workload_plan(args.workload)
    n_rot = 3     # rotating inputs: 3 x 143 MB per rank, each larger than the 126 MB L2
    by_patch = args.sharding == "patch" and world > 1     # one volume per step, its patches split over the ranks
    base = 0 if by_patch else rank * n_rot
    vols_h = [synth_volume(base + i)[0].pin_memory() for i in range(n_rot)]
    tgts_h = [synth_target(base + i).pin_memory() for i in range(n_rot)]
    vols_d = [v.cuda() for v in vols_h]
    tgts_d = [t.cuda() for t in tgts_h]
    keeps = [keep_scales(base + i, n_patches) for i in range(n_rot)]
    lab_h = torch.empty(SHAPE, dtype=torch.uint8).pin_memory()

    if by_patch:
        from dcl_b200 import sharded
        if starts is None:
            raise SystemExit("--sharding patch needs a weighted workload (overlap50 / overlap75)")
        stage = torch.empty_like(vols_d[0])

        def step_dev(i):
            j = i % n_rot
            return sharded.predict_volume_sharded(eng, vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j])

        def step_e2e(i):     # every rank uploads the volume over its own PCIe link; rank 0 downloads the label map
            j = i % n_rot
            stage.copy_(vols_h[j], non_blocking=True)
            tgt = tgts_h[j].cuda(non_blocking=True)
            out = sharded.predict_volume_sharded(eng, stage, mode, starts=starts, keep_scales=keeps[j], target=tgt)
            if rank == 0:
                lab_h.copy_(out["labels"], non_blocking=True)
                out["counts"].cpu()
            torch.cuda.synchronize()
            return out
    elif mode == "TTA":
        stage = torch.empty_like(vols_d[0])

        def step_dev(i):
            j = i % n_rot
            return eng.predict_volume_tta(vols_d[j], keep_scales=keeps[j], target=tgts_d[j], want_probs=False)

        def step_e2e(i):
            j = i % n_rot
            stage.copy_(vols_h[j], non_blocking=True)
            out = eng.predict_volume_tta(stage, keep_scales=keeps[j], target=tgts_h[j].cuda(non_blocking=True), want_probs=False)
            lab_h.copy_(out["labels"], non_blocking=True)
            out["counts"].cpu()
            torch.cuda.synchronize()
            return out
    else:
        def step_dev(i):
            j = i % n_rot
            return eng.predict_volume(vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j],
                                      want_probs=False, want_labels=True)

        def step_e2e(i):
            j = i % n_rot
            return eng.predict_volume_host(vols_h[j], mode, starts=starts, keep_scales=keeps[j], target_host=tgts_h[j],
                                           labels_out=lab_h)

    # ---- device-resident throughput (no per-kernel events inside this region) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # before the warm-up: its start-up must not land in the timed region
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.region_begin()
    ev0.record()
    for i in range(args.steps):
        out = step_dev(i)
    ev1.record()
    barrier()
    sampler.region_end()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    counts = out["counts"].cpu().numpy()

    # ---- per-kernel-class device timing: the same steps again with CUDA events recorded around every launch of
    # the class on its launching stream (the events perturb the step, so they stay out of the region above) ----
    prof_steps = min(args.steps, 2)
    eng.profile(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(prof_steps):
        step_dev(i)
    pe1.record()
    torch.cuda.synchronize()
    prof_ms = pe0.elapsed_time(pe1)
    eng.profile(False)
    conv_ms, conv_n, conv_flops = eng.profile_read(0)
    tail_ms, tail_n, tail_bytes = eng.profile_read(1)
    kinds = {2: "conv3d_k3_roll_kernel<16ch @128^3> (tcgen05 rolling implicit GEMM)",
             3: "conv3d_k3_roll_kernel<32ch @64^3>", 4: "conv_slab_kernel", 5: "conv_gemm_kernel (stride 2)",
             12: "conv3d_k3s2_roll_kernel"}
    per_kind = {k: eng.profile_read(k) for k in kinds}
    # the overlap stitch alone (dcl_bench_stitch: 20 back-to-back volumes of per-patch slots, > L2), both forms
    stitch_iso = None
    stride = WORKLOADS[args.workload][1]
    if rank == 0 and mode != "TTA" and stride is not None:
        import ctypes as C
        from dcl_b200 import _native as N
        lib, npat = N.load_library(), C.c_int32()
        shp = (C.c_int32 * 3)(*SHAPE)
        us_g = lib.dcl_bench_stitch(shp, stride, 0, 0, 20, C.byref(npat))
        us_a = lib.dcl_bench_stitch(shp, stride, 0, 1, 20, C.byref(npat))
        algo = npat.value * 4 * 128 ** 3 * 4 + VOXELS        # every patch probability once + one label byte per voxel
        hbm = peaks()["hbm"]
        stitch_iso = {"algorithmic_bytes": algo, "gather_form_us": us_g, "gather_form_GB/s": algo / us_g / 1e3,
                      "gather_form_frac": algo / us_g / 1e3 / hbm, "accumulate_form_us": us_a,
                      "accumulate_form_frac": algo / us_a / 1e3 / hbm}

    # ---- end to end through the host-buffer C-ABI call ----
    # The call is synchronous (upload, compute, download, sync), so a throughput-minded caller keeps two volumes in
    # flight: two worker threads, each with its own handle and stream (ctypes drops the GIL during the call), so the
    # host-to-device copy of one volume overlaps the compute of the other.  --e2e-workers 1 is the plain serial loop.
    workers = 1 if (by_patch or mode == "TTA") else max(1, args.e2e_workers)
    if workers == 1:
        for i in range(max(1, args.warmup // 2)):
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = step_e2e(i)
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        engines = [eng] + [dcl_b200.Engine(prec) for _ in range(workers - 1)]
        for e in engines[1:]:
            e.load_state_dict(seed0_weights())
        streams = [torch.cuda.Stream() for _ in range(workers)]
        labs = [torch.empty(SHAPE, dtype=torch.uint8).pin_memory() for _ in range(workers)]

        def worker(w, lo, hi):
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[w]):
                for i in range(lo + w, hi, workers):
                    j = i % n_rot
                    engines[w].predict_volume_host(vols_h[j], mode, starts=starts, keep_scales=keeps[j],
                                                   target_host=tgts_h[j], labels_out=labs[w])

        def run(lo, hi):
            ths = [threading.Thread(target=worker, args=(w, lo, hi)) for w in range(workers)]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()

        run(0, 2 * workers)          # warm-up: every engine captures its graph
        barrier()
        t0 = time.perf_counter()
        run(0, args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        for e in engines[1:]:
            e.close()

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        vols_per_step = 1 if by_patch else world
        vps = vols_per_step * args.steps / (ms / 1e3)
        e2e_vps = vols_per_step * args.steps / (e2e_ms / 1e3)
        conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        tail_gbs = tail_bytes / (tail_ms * 1e-3) / 1e9 if tail_ms > 0 else 0.0
        line = {
            "metric": "BraTS 4x240x240x155 volumes/sec", "value": vps, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_patch else "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32", "bf16x3": "bf16x3 (split bf16 operands, fp32 accumulate)", "bf16": "bf16"}[
                args.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "patches_per_volume": n_patches,
                       "weights": "random-init seed 0",
                       "sharding": ("one volume per step, patch slabs across ranks, NCCL reduce-scatter of the accumulators"
                                    if by_patch else "volumes across ranks, no collective"),
                       "l2": "inputs larger than L2: 3 rotating 143 MB volumes per rank, >1.8 GB of activations per patch"},
            "e2e": {"value": e2e_vps, "unit": "volumes/s", "h2d_bytes_per_step": 4 * VOXELS * 4 + VOXELS,
                    "d2h_bytes_per_step": VOXELS + 13 * 8, "workers": workers,
                    "call": "dcl_predict_volume_host (pinned host volume + target in, host labels + 13 counters out)"},
            "gpu_launches": launches,
            "model_tflops": vols_per_step * args.steps * n_patches * FLOPS_PER_PATCH / (ms / 1e3) / 1e12,
            "roofline": dominant_roofline(per_kind, kinds, conv_ms, conv_n, conv_flops, prof_ms, pk),
            "roofline_all_k3_convs": {"bound": "tensor", "achieved": conv_tflops, "peak": pk["tensor"], "unit": "TFLOP/s",
                                      "frac": conv_tflops / pk["tensor"], "launches": conv_n,
                                      "kernel": "every 3x3x3 convolution launch of the profiled steps (rolling / slab / stride-2 / "
                                                "im2col kernels in bf16 mode, FFMA kernel in fp32 mode)",
                                      "share_of_step": conv_ms / prof_ms, "peak_source": pk["source"] + " sustained bf16 dense"},
            "roofline_accumulate": {"bound": "hbm", "achieved": tail_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": tail_gbs / pk["hbm"], "launches": tail_n,
                                    "kernel": ("gather_finalize_kernel (overlap blend + normalise + arg-max + counters in one "
                                               "pass over the per-patch probability slots)"
                                               if (stitch_iso and not by_patch and os.environ.get("DCL_GATHER", "1") != "0")
                                               else "accumulate / stitch_copy / finalize_labels"),
                                    "share_of_step": tail_ms / prof_ms, "peak_source": pk["source"] + " copy",
                                    "note": "timed inside the profiled steps, right after the last patch forward",
                                    **({"isolated": stitch_iso} if stitch_iso else {})},
            "clocks": clocks,
            "label_hist": counts[:4].tolist(),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, sec = cpu_reference_volumes_per_s(args.workload, 2, threads, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": v, "unit": "volumes/s", "cores": threads, "kind": "port",
                                    "sample": f"2 of {n_patches} oracle patch forwards (torch CPU fp32) extrapolated to "
                                              f"{n_patches} + full-volume stitch/argmax/Dice tail, one step"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def dominant_roofline(per_kind, kinds, conv_ms, conv_n, conv_flops, prof_ms, pk):
    """Roofline line of the kernel with the largest share of the step (CUDA events around each of its launches on
    the launching stream; algorithmic FLOPs = 2 x MACs of the convolutions it ran).  `traffic` = DRAM bytes per
    launch from the committed ncu --set full capture of that kernel (profiles/), when one exists."""
    best = max(per_kind, key=lambda k: per_kind[k][0]) if any(v[1] for v in per_kind.values()) else None
    if best is None:      # fp32 mode: one FFMA kernel runs every convolution
        ms, n, flops, name = conv_ms, conv_n, conv_flops, "conv3d_k3_kernel (fp32 FFMA)"
    else:
        (ms, n, flops), name = per_kind[best], kinds[best]
    tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_full_dominant.json")
    if best is not None and os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("kind") == best:
            traffic = t.get("dram_bytes_per_launch")
    return {"bound": "tensor", "achieved": tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"],
            "traffic": traffic, "kernel": name, "launches": n, "avg_launch_ms": ms / max(n, 1),
            "flops_per_launch": flops / max(n, 1), "share_of_step": ms / prof_ms,
            "peak_source": pk["source"] + " sustained bf16 dense",
            "note": "narrow-N (16..48) MMAs are bound by the A-operand shared-memory read and the layer by HBM "
                    "(201 MB per launch), see DESIGN.md section 3"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--workload", default="overlap50", choices=sorted(WORKLOADS))
    ap.add_argument("--sharding", default="volume", choices=["volume", "patch"],
                    help="N > 1: 'volume' = one volume per rank per step (no collective, weak scaling); "
                         "'patch' = one volume per step split by patch slab with an NCCL exchange (strong scaling)")
    ap.add_argument("--e2e-workers", type=int, default=2,
                    help="host threads (each with its own handle and stream) feeding the end-to-end loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
