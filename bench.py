#!/usr/bin/env python
"""Benchmark of the sliding-window ClsWiseFormer inference path (BASELINE.json metric: BraTS
4x240x240x155 volumes/sec).

    python bench.py [--gpus N --steps K --warmup W] [--precision f16x3|bf16|fp32] [--workload overlap50|reference8|overlap75]
    python bench.py --impl reference ...      # the CPU oracle port of the reference path on the host cores

A step = one full volume: patch gather -> clswiseformer forward per 128^3 patch -> overlap
accumulate / stitch -> normalise + arg-max + label histogram + Dice counters.  Default workload is
BASELINE.json configs[1]: 50 % overlap (stride 64 -> 18 patches, uniform blend).  Prints ONE JSON line.

The headline (`value`, `e2e`, `roofline`) is the PARITY-GRADE mode DCL_BF16X3 (split-fp16 operands on the tcgen05
kernels: meets the fp32 tolerances of north_star); the plain bf16 mode (2e-2 class) is measured in the same run and
reported beside it under "bf16".  "parity" holds the observed error of each mode on this very workload against the
golden minted from the unmodified reference (tests/golden/make_golden_overlap50.py), computed outside the timed region.
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
             "tta8": ("TTA", None),
             # BASELINE.json configs[3]: the sliding window whose output carries WT/TC/ET class probabilities AND the six
             # auxiliary region / edge heads (cls_wise_former.py:545-546, :585-592), 16 blended channels per voxel
             "overlap50_aux": ("UNIFORM", 64)}


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
def cpu_reference_volumes_per_s(workload, n_sample_patches, threads, steps=1, warmup=0, budget_s=None):
    """The oracle port of the reference path on the host cores: `n_sample_patches` real patch forwards per step
    (None = every patch of the workload; fewer = extrapolated to the workload's patch count) + the full-size
    stitch / arg-max / Dice tail.  budget_s: when the first step shows that all steps would not fit, the remaining
    steps fall back to a bounded sample (returned so that the caller can say so).
    Returns (volumes/s, extrapolated s/volume, measured wall s per step, patches actually run per step)."""
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
    per_step, wall_step, ran = [], [], []
    if n_sample_patches is None:
        n_sample_patches = len(starts)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        probs = []
        for j in range(n_sample_patches):
            sx, sy, sz = starts[j]
            probs.append(O.forward(sd, x[..., sx:sx + 128, sy:sy + 128, sz:sz + 128],
                                   torch.from_numpy(ks[j:j + 1]), want_aux=workload == "overlap50_aux")[0][0].numpy())
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
        if it >= warmup or steps == 0:
            per_step.append(t_patch * len(starts) + t_tail)
            wall_step.append(time.perf_counter() - t0)
            ran.append(n_sample_patches)
        if it == 0 and budget_s is not None:
            full = t_patch * len(starts) + t_tail
            if full * (warmup + steps) > budget_s:
                n_sample_patches = max(1, min(len(starts), int((budget_s / (warmup + steps) - t_tail) / t_patch)))
    sec = statistics.mean(per_step)
    return 1.0 / sec, sec, statistics.mean(wall_step), min(ran)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    # every step runs EVERY patch of the workload (18 forwards, ~17 s on 16 cores) unless the declared steps would
    # not fit 15 minutes - then the later steps run a bounded number of patches and the line says so
    vps, sec, wall_step, n_ran = cpu_reference_volumes_per_s(args.workload, None, threads, steps=args.steps, warmup=args.warmup,
                                                            budget_s=900.0)
    from oracle import stitch_oracle as S
    stride = WORKLOADS[args.workload][1]
    n_patches = 8 if stride is None else len(S.patch_starts(SHAPE, stride))
    sample = (f"all {n_patches} patch forwards of the volume in every step" if n_ran == n_patches else
              f"{n_ran} of {n_patches} patch forwards per step, extrapolated x{n_patches / n_ran:g} (time budget)")
    sample += (f" on {threads} host threads (torch CPU fp32 oracle port of predict_overlap.py + clswiseformer) plus the "
               f"full-volume blend/argmax/Dice tail; ms_per_step is the measured wall time of a step")
    line = {
        "impl": "reference", "metric": "BraTS 4x240x240x155 volumes/sec", "value": vps, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config_dict(args, n_patches, args.sharding == "patch" and args.gpus > 1),
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, n_patches, by_patch=False):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workload_name(args.workload), "patches_per_volume": n_patches, "weights": "random-init seed 0",
            "sharding": ("one volume per step, patch slabs across ranks, NCCL exchange of the overlap halos"
                         if by_patch else "volumes across ranks, no collective"),
            "l2": "inputs larger than L2: 3 rotating 143 MB volumes per rank, >1.8 GB of activations per patch",
            "schedule": "value: one handle, 3 patches in flight; e2e: two host threads with a handle each (upload of one "
                        "volume overlaps the compute of the other)"}


def workload_name(w):
    return {"overlap50": "predict_overlap sliding window, one 4x240x240x155 volume, 128^3 patches at 50% overlap "
                         "(stride 64, 18 patches, uniform blend)",
            "overlap75": "4x240x240x155 volume, 128^3 patches at 75% overlap (stride 32, 50 patches, uniform blend)",
            "reference8": "predict_overlap.py 8-corner tiling + crop-overwrite stitch, one 4x240x240x155 volume",
            "tta8": "predict_cls.py 8-flip TTA around the 8-corner tiling (64 patch forwards), one 4x240x240x155 volume",
            "overlap50_aux": "sliding window with WT/TC/ET + edge outputs (4 class + 6 x 2 auxiliary-head channels blended per "
                             "voxel), one 4x240x240x155 volume, 128^3 patches at 50% overlap (18 patches, uniform blend)"}[w]


def measure_device(eng, step_dev, args, barrier, sampler):
    """W untimed warm-up steps, then exactly K timed steps between CUDA events (inputs resident in HBM)."""
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.region_begin()
    ev0.record()
    out = None
    for i in range(args.steps):
        out = step_dev(i)
    ev1.record()
    barrier()
    if sampler is not None:
        sampler.region_end()
    return {"ms": ev0.elapsed_time(ev1), "launches": eng.launch_count - l0, "out": out}


KINDS = {2: "conv3d_k3_roll_kernel<16ch @128^3> (tcgen05 rolling implicit GEMM)",
         3: "conv3d_k3_roll_kernel<32ch @64^3>", 4: "conv_slab_kernel", 5: "conv_gemm_kernel (stride 2)",
         12: "conv3d_k3s2_roll_kernel"}


def profile_pass(eng, step_dev, prof_steps):
    """Per-kernel-class device timing: the same steps again with CUDA events recorded around every launch of the class
    on its launching stream (the events perturb the step, so they stay out of the timed region)."""
    eng.profile(True)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(prof_steps):
        step_dev(i)
    pe1.record()
    torch.cuda.synchronize()
    prof_ms = pe0.elapsed_time(pe1)
    eng.profile(False)
    return {"prof_ms": prof_ms, "conv": eng.profile_read(0), "tail": eng.profile_read(1),
            "per_kind": {k: eng.profile_read(k) for k in KINDS}}


def measure_patch_sharded(eng, args, mode, starts, n_patches, rank, world, barrier):
    """ONE volume per step, its patches split over the ranks (dcl_b200/sharded.py, owner-computes form): device-resident
    and end-to-end (every rank uploads only the x-slab its patches read, rank 0 downloads the label map), timed like the
    main loop (CUDA events / wall clock between barriers, max over ranks).  All ranks work on the same volumes."""
    import torch.distributed as dist
    from dcl_b200 import sharded
    n_rot = 3
    vols_h = [synth_volume(i)[0].pin_memory() for i in range(n_rot)]
    tgts_h = [synth_target(i).pin_memory() for i in range(n_rot)]
    vols_d = [v.cuda() for v in vols_h]
    tgts_d = [t.cuda() for t in tgts_h]
    keeps = [keep_scales(i, n_patches) for i in range(n_rot)]
    stage = torch.empty_like(vols_d[0])
    lab_h = torch.empty(SHAPE, dtype=torch.uint8).pin_memory()

    def step_dev(i):
        j = i % n_rot
        return sharded.predict_volume_sharded(eng, vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j])

    up_bytes = [0]

    def step_e2e(i):
        j = i % n_rot
        up_bytes[0] = sharded.upload_own_region(vols_h[j], stage, starts)
        out = sharded.predict_volume_sharded(eng, stage, mode, starts=starts, keep_scales=keeps[j], target=tgts_h[j])
        if rank == 0:
            lab_h.copy_(out["labels"], non_blocking=True)
            out["counts"].cpu()
        torch.cuda.synchronize()
        return out

    for i in range(max(2, args.warmup)):
        step_dev(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        out = step_dev(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    # the sharded label map equals the one-GPU label map bit for bit: checked here on the last volume, outside the timing
    j = (args.steps - 1) % n_rot
    one = eng.predict_volume(vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j], want_probs=False)
    identical = bool(torch.equal(one["labels"], out["labels"]) and torch.equal(one["counts"], out["counts"]))
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, e2e_ms, float(up_bytes[0])], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, up_max = t.tolist()
    parts = sharded.partition_patches(n_patches, world)
    slot = 4 * 128 ** 3 * 4
    return {"scaling": "strong", "value": args.steps / (ms / 1e3), "unit": "volumes/s", "ms_per_volume": ms / args.steps,
            "e2e": {"value": args.steps / (e2e_ms / 1e3), "unit": "volumes/s", "ms_per_volume": e2e_ms / args.steps,
                    "h2d_bytes_per_step_max_rank": int(up_max) + -(-SHAPE[0] // world) * SHAPE[1] * SHAPE[2],
                    "d2h_bytes_per_step_rank0": VOXELS + 13 * 8},
            "patches_per_rank": [c for _, c in parts], "load_balance_bound": n_patches / (world * max(c for _, c in parts)),
            "exchange": "owner-computes: gather_finalize_kernel reads the covering patches' probability slots in place - "
                        "local HBM or a peer's through CUDA IPC over NVLink; collectives: 1 x all_reduce(int32) barrier, "
                        "1 x all_gather(uint8 label rows), 1 x all_reduce(13 x int64)",
            "nvlink_bytes_pulled_per_volume": f"<= {n_patches * slot} (every remote probability at most once), about 1/{world} of it per rank",
            "labels_bit_identical_to_one_gpu": identical}


def unpack2(packed, n):
    a = np.asarray(packed, dtype=np.uint8)
    out = np.stack([a & 3, (a >> 2) & 3, (a >> 4) & 3, (a >> 6) & 3], 1).ravel()
    return out[:n]


def parity_block(eng, mode, starts):
    """This engine's result on volume 0 of the workload against tests/golden/overlap50_seed1000.npz (18 CPU forwards of
    the UNMODIFIED reference model + the sum-then-divide blend): sampled probability error, label flips over the WHOLE
    volume, per-region Dice delta.  Outside every timed region."""
    path = os.path.join(ROOT, "tests", "golden", "overlap50_seed1000.npz")
    if not os.path.exists(path):
        return None
    from dcl_b200.engine import dice_from_counts
    g = np.load(path)
    vol = synth_volume(0).cuda()
    tgt = synth_target(0).cuda()
    out = eng.predict_volume(vol, mode, starts=starts, keep_scales=g["keep_scale"], target=tgt)
    torch.cuda.synchronize()
    flat = out["probs"].flatten()
    step = max(1, flat.numel() // 4096)
    samp = flat[::step][:4096].float().cpu().numpy().astype(np.float64)
    want = g["blend/sample"].astype(np.float64)
    labels = out["labels"].cpu().numpy().ravel()
    lstep = int(g["labels_step"])
    ref_labels = unpack2(g["labels_packed"], labels[::lstep].size)
    dice = np.asarray(dice_from_counts(out["counts"].cpu().numpy()), dtype=np.float64)
    return {"against": "unmodified reference model (CPU fp32), tests/golden/overlap50_seed1000.npz",
            "probs_rel": float(np.abs(samp - want).max() / np.abs(want).max()),
            "label_flip_frac": float((labels[::lstep] != ref_labels).mean()), "label_voxels_compared": int(ref_labels.size),
            "dice_delta": float(np.abs(dice - g["dice"]).max()),
            "tolerances": {"probs_rel": 1e-3, "label_flip_frac": 1e-4, "dice_delta": 1e-3}}


def all_convs_roofline(prof, pk):
    conv_ms, conv_n, conv_flops = prof["conv"]
    tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    return {"bound": "tensor", "achieved": tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"],
            "frac_of_nominal_2250": tf / 2250.0, "launches": conv_n,
            "kernel": "every 3x3x3 convolution launch of the profiled steps (rolling / slab / stride-2 / im2col kernels in "
                      "the tcgen05 modes, FFMA kernel in fp32 mode); algorithmic FLOPs = 2 x MACs",
            "share_of_step": conv_ms / prof["prof_ms"], "peak_source": pk["source"] + " sustained bf16 dense"}



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

    if args.precision == "f16x3":
        args.precision = "f16x3"
    prec = {"fp32": dcl_b200.Precision.FP32, "f16x3": dcl_b200.Precision.F16X3, "bf16": dcl_b200.Precision.BF16}[
        args.precision]
    with_aux = args.workload == "overlap50_aux"
    eng = dcl_b200.Engine(prec, want_aux=with_aux)
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

        def step_e2e(i):     # a rank uploads the x-slab its patches read over its own PCIe link; rank 0 downloads the label map
            j = i % n_rot
            sharded.upload_own_region(vols_h[j], stage, starts)
            out = sharded.predict_volume_sharded(eng, stage, mode, starts=starts, keep_scales=keeps[j], target=tgts_h[j])
            if rank == 0:
                lab_h.copy_(out["labels"], non_blocking=True)
                out["counts"].cpu()
            torch.cuda.synchronize()
            return out
    elif with_aux:
        stage = torch.empty_like(vols_d[0])

        def step_dev(i):
            j = i % n_rot
            return eng.predict_volume_aux(vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j], want_probs=False)

        def step_e2e(i):     # host volume + target in, host labels + counters out; the 12 blended auxiliary channels stay on the device
            j = i % n_rot
            stage.copy_(vols_h[j], non_blocking=True)
            out = eng.predict_volume_aux(stage, mode, starts=starts, keep_scales=keeps[j], target=tgts_h[j].cuda(non_blocking=True),
                                         want_probs=False)
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
    m = measure_device(eng, step_dev, args, barrier, sampler)
    ms, launches, out = m["ms"], m["launches"], m["out"]
    clocks = sampler.stop() if rank == 0 else None
    counts = out["counts"].cpu().numpy()
    prof = profile_pass(eng, step_dev, min(args.steps, 2))
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

    # ---- observed parity of this mode on THIS workload against the golden from the unmodified reference ----
    parity = None
    if rank == 0 and args.workload == "overlap50" and not by_patch:
        parity = parity_block(eng, mode, starts)

    # ---- the plain bf16 mode beside the parity-grade headline (same workload, same run) ----
    bf16_rec = None
    if rank == 0 and world == 1 and args.precision == "f16x3" and mode != "TTA" and not args.no_bf16 and not with_aux:
        e16 = dcl_b200.Engine(dcl_b200.Precision.BF16)
        e16.load_state_dict(seed0_weights())

        def step16(i):
            j = i % n_rot
            return e16.predict_volume(vols_d[j], mode, starts=starts, keep_scales=keeps[j], target=tgts_d[j],
                                      want_probs=False, want_labels=True)

        m16 = measure_device(e16, step16, args, barrier, None)
        p16 = profile_pass(e16, step16, min(args.steps, 2))
        pk16 = peaks()
        bf16_rec = {"value": args.steps / (m16["ms"] / 1e3), "unit": "volumes/s", "ms_per_step": m16["ms"] / args.steps,
                    "dtype": "bf16 operands, fp32 accumulate (2e-2 class: NOT within the label budget, see parity)",
                    "gpu_launches": m16["launches"],
                    "roofline": dominant_roofline(p16, pk16, "bf16"),
                    "roofline_all_k3_convs": all_convs_roofline(p16, pk16),
                    "parity": parity_block(e16, mode, starts) if args.workload == "overlap50" else None}
        e16.close()

    # ---- end to end through the host-buffer C-ABI call ----
    # The call is synchronous (upload, compute, download, sync), so a throughput-minded caller keeps two volumes in
    # flight: two worker threads, each with its own handle and stream (ctypes drops the GIL during the call), so the
    # host-to-device copy of one volume overlaps the compute of the other.  --e2e-workers 1 is the plain serial loop.
    workers = 1 if (by_patch or mode == "TTA" or with_aux) else max(1, args.e2e_workers)
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

    # ---- N > 1: the SAME run also measures ONE volume sharded by patch slab over the ranks (BASELINE configs 3 / 5:
    # strong scaling, owner-computes exchange over NVLink), attached as a sub-record ----
    sharded_rec = None
    if world > 1 and not by_patch and mode != "TTA" and starts is not None and not args.no_sharded and not with_aux:
        try:
            sharded_rec = measure_patch_sharded(eng, args, mode, starts, n_patches, rank, world, barrier)
        except Exception as exc:      # e.g. CUDA IPC unavailable in this container: the weak-scaling line must still be printed
            sharded_rec = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        vols_per_step = 1 if by_patch else world
        vps = vols_per_step * args.steps / (ms / 1e3)
        e2e_vps = vols_per_step * args.steps / (e2e_ms / 1e3)
        tail_ms, tail_n, tail_bytes = prof["tail"]
        tail_gbs = tail_bytes / (tail_ms * 1e-3) / 1e9 if tail_ms > 0 else 0.0
        line = {
            "metric": "BraTS 4x240x240x155 volumes/sec", "value": vps, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if by_patch else "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32 (FFMA)",
                      "f16x3": "f16x3: split operands (fp16 hi + fp16 lo), 3-4 tcgen05 MMAs per product, fp32 accumulate - parity-grade",
                      "bf16": "bf16"}[args.precision],
            "data": "synthetic",
            "config": config_dict(args, n_patches, by_patch),
            "e2e": {"value": e2e_vps, "unit": "volumes/s", "h2d_bytes_per_step": 4 * VOXELS * 4 + VOXELS,
                    "d2h_bytes_per_step": VOXELS + 13 * 8, "workers": workers,
                    "call": "dcl_predict_volume_host (pinned host volume + target in, host labels + 13 counters out)"},
            "gpu_launches": launches,
            "model_tflops": vols_per_step * args.steps * n_patches * (FLOPS_PER_PATCH + (8.40e9 if with_aux else 0.0)) / (ms / 1e3) / 1e12,
            "roofline": dominant_roofline(prof, pk, args.precision),
            "roofline_all_k3_convs": all_convs_roofline(prof, pk),
            "roofline_accumulate": {"bound": "hbm", "achieved": tail_gbs, "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": tail_gbs / pk["hbm"], "frac_of_nominal_8000": tail_gbs / 8000.0, "launches": tail_n,
                                    "kernel": ("gather_finalize_kernel (overlap blend + normalise + arg-max + counters in one "
                                               "pass over the per-patch probability slots)"
                                               if (stitch_iso and not by_patch and os.environ.get("DCL_GATHER", "1") != "0")
                                               else "accumulate / stitch_copy / finalize_labels"),
                                    "share_of_step": tail_ms / prof["prof_ms"], "peak_source": pk["source"] + " copy",
                                    "note": "timed inside the profiled steps, right after the last patch forward",
                                    **({"isolated": stitch_iso} if stitch_iso else {})},
            "clocks": clocks,
            "label_hist": counts[:4].tolist(),
            "parity": parity,
        }
        if bf16_rec is not None:
            line["bf16"] = bf16_rec
        if sharded_rec is not None:
            line["single_volume_sharded"] = sharded_rec
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = min(6, n_patches)
            v, sec, _wall, _n = cpu_reference_volumes_per_s(args.workload, n_cpu, threads, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": v, "unit": "volumes/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_cpu} of {n_patches} oracle patch forwards (torch CPU fp32) extrapolated to "
                                              f"{n_patches} + full-volume blend/argmax/Dice tail, one step; the reference arm "
                                              f"(--impl reference) runs all {n_patches}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# per-launch HBM bytes of the dominant kernels (bf16 B-format activations; x3 moves hi + lo planes = twice the bytes)
def _roll16_bytes(x3):
    plane = 16 * 128 ** 3 * 2 * (2 if x3 else 1)
    return {"min": 2 * plane, "with_residual": 3 * plane}


def dominant_roofline(prof, pk, precision):
    """Roofline line of the kernel with the largest share of the step (CUDA events around each of its launches on
    the launching stream; algorithmic FLOPs = 2 x MACs of the convolutions it ran).  `traffic` = DRAM bytes per
    launch from the committed ncu --set full capture of that kernel (profiles/), when one exists."""
    per_kind, prof_ms = prof["per_kind"], prof["prof_ms"]
    conv_ms, conv_n, conv_flops = prof["conv"]
    best = max(per_kind, key=lambda k: per_kind[k][0]) if any(v[1] for v in per_kind.values()) else None
    if best is None:      # fp32 mode: one FFMA kernel runs every convolution
        ms, n, flops, name = conv_ms, conv_n, conv_flops, "conv3d_k3_kernel (fp32 FFMA)"
    else:
        (ms, n, flops), name = per_kind[best], KINDS[best]
    tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02n_ncu_full_dominant_f16x3.json" if precision == "f16x3" else "r02_ncu_full_dominant_%s.json" % precision)
    if best is not None and os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("kind") == best:
            traffic = t.get("dram_bytes_per_launch")
    x3 = precision == "f16x3"
    rec = {"bound": "tensor", "achieved": tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": tf / pk["tensor"],
           "frac_of_burst_peak": tf / pk["tensor_burst"], "frac_of_nominal_2250": tf / 2250.0,
           "traffic": traffic, "kernel": name, "launches": n, "avg_launch_ms": ms / max(n, 1),
           "flops_per_launch": flops / max(n, 1), "share_of_step": ms / prof_ms,
           "peak_source": pk["source"] + " sustained bf16 dense"}
    # what the tensor pipe EXECUTES per algorithmic MAC: split-fp16 forms a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (+ a_lo*w_lo where
    # the weights ride stacked along N: the rolling kernels) - 3 bf16 MMAs in the GEMM / slab / stride-2 kernels, 4 in the
    # rolling kernels; north_star's "tensor-pipe utilisation against the dense bf16 peak" is executed MMA FLOPs / peak
    factor = (4 if best in (2, 3) else 3) if x3 else 1
    rec["tensor_pipe"] = {"executed_mma_flops_per_algorithmic_flop": factor, "executed_TFLOP/s": tf * factor,
                          "utilisation_of_sustained_peak": tf * factor / pk["tensor"],
                          "utilisation_of_nominal_2250": tf * factor / 2250.0,
                          "ncu_counter": "sm__pipe_tensor_subpipe_hmma_cycles_active / (8 x sm__cycles_elapsed): "
                                         "profiles/r02n_ncu_full_roll16_f16x3.csv (0.200 for the split-fp16 16-channel layer; "
                                         "0.148 before the issue loop moved to the uniform datapath and 0.118 for the bf16 one, "
                                         "profiles/r02_ncu_full_conv_kernels.csv; that counter's own peak is ~1.87 x the bf16 dense peak)"}
    if x3:
        rec["note"] = ("`achieved` / `frac` count ALGORITHMIC FLOPs (2 x MACs of the convolution); the split-fp16 kernels execute "
                       f"{factor} tensor-core MACs per algorithmic MAC, so the ceiling of `frac` is 1/{factor} and the pipe's own "
                       "utilisation is `tensor_pipe`")
    if best == 2 and n > 0:
        # the 16-channel 128^3 layers are HBM-side too: 2 (3 with the residual) passes over a 16-channel tensor
        b = _roll16_bytes(x3)
        gbs = b["min"] / (ms / n * 1e-3) / 1e9
        rec["hbm"] = {"algorithmic_bytes_per_launch": b["min"], "with_residual": b["with_residual"], "achieved_GB/s": gbs,
                      "frac": gbs / pk["hbm"], "frac_of_nominal_8000": gbs / 8000.0,
                      "note": "max(t_tensor, t_HBM) bound of SURVEY H1: both fractions are reported"}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "f16x3", "bf16x3", "bf16"],
                    help="f16x3 (default; bf16x3 is the same mode's first name) = the parity-grade split-operand tensor-core mode; "
                         "bf16 = the 2e-2 class mode")
    ap.add_argument("--no-bf16", action="store_true", help="skip the plain-bf16 sub-record of an f16x3 run")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the single-volume-sharded sub-record")
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
