# Round-end evidence on ONE B200 (run under gpurun from the repo root): writes gpurun_out/r02n_*; copy what is to be judged into profiles/.
T=r02n
python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
python bench.py --steps 10 --warmup 3 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_f16x3_overlap50.json
DCL_LANES=4 python bench.py --steps 10 --warmup 3 --no-bf16 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${T}_bench_lanes4.json
DCL_LANES=2 python bench.py --steps 10 --warmup 3 --no-bf16 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${T}_bench_lanes2.json
python bench.py --workload overlap50_aux --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${T}_bench_f16x3_overlap50_aux.json
python bench.py --workload reference8 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${T}_bench_f16x3_reference8.json
python tools/op_time.py x3 > gpurun_out/${T}_op_time_x3.log 2>&1
python tools/op_time.py > gpurun_out/${T}_op_time_bf16.log 2>&1
TRACE_X3=1 python tools/trace_kernel.py all > gpurun_out/${T}_trace_x3.log 2>&1
python tools/trace_kernel.py all > gpurun_out/${T}_trace_bf16.log 2>&1
python tools/one_patch.py f16x3 2 > gpurun_out/${T}_one_patch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${T}_launches_f16x3_one_patch.csv python tools/one_patch.py f16x3 1 > gpurun_out/${T}_ncu_launches.log 2>&1
python tools/op_time.py x3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3d_k3_roll --launch-skip 22 -c 4 -o gpurun_out/${T}_prof_roll_x3 -f python tools/op_time.py x3 > gpurun_out/${T}_ncu_full.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02n_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"], 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 2), round(d["roofline"]["frac"], 3),
              d.get("bf16", {}).get("value"), d.get("parity", {}).get("label_flip_frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
