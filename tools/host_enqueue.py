"""Is the device-resident loop host-bound?  Host time spent enqueueing a volume (perf_counter around the call, no
synchronisation) against the device time of the same volumes (CUDA events), for one caller thread and for two
(a handle and a stream each).  Usage: python tools/host_enqueue.py [workload] [steps]"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import dcl_b200  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "overlap50"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    mode, starts, n_patches = B.workload_plan(workload)
    sd = B.seed0_weights()
    vols = [B.synth_volume(i)[0].cuda() for i in range(3)]
    print("affinity", sorted(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())

    def make():
        eng = dcl_b200.Engine(dcl_b200.Precision.BF16)
        eng.load_state_dict(sd)
        return eng

    def run(eng, n, stream, out):
        with torch.cuda.stream(stream):
            t_host = 0.0
            for i in range(n):
                t0 = time.perf_counter()
                eng.predict_volume(vols[i % 3], mode, starts=starts, want_probs=False)
                t_host += time.perf_counter() - t0
            out.append(t_host)

    engines = [make(), make()]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for e, s in zip(engines, streams):
        run(e, 3, s, [])
    torch.cuda.synchronize()
    for workers in (1, 2, 1, 2):
        outs = [[] for _ in range(workers)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ts = [threading.Thread(target=run, args=(engines[w], steps, streams[w], outs[w])) for w in range(workers)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        vols_done = workers * steps
        print(f"workers={workers}: {vols_done / t_all:6.2f} volumes/s; wall per volume {t_all / vols_done * 1e3:6.2f} ms; "
              f"host enqueue per volume (per thread) {sum(o[0] for o in outs) / vols_done * 1e3:6.2f} ms; "
              f"all enqueued after {t_enq * 1e3:7.1f} of {t_all * 1e3:7.1f} ms ({n_patches} patches per volume)")


if __name__ == "__main__":
    main()
