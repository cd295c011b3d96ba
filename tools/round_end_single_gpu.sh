python -m pytest tests -x -q -m gpu > gpurun_out/r01m_pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/r01m_pytest_gpu.log; tail -4 gpurun_out/r01m_pytest_gpu.log
python bench.py > gpurun_out/r01m_bench_bf16_overlap50.json 2> gpurun_out/r01m_bench.err
DCL_LANES=4 python bench.py --no-cpu-baseline > gpurun_out/r01m_bench_lanes4.json 2>/dev/null
python bench.py --workload reference8 > gpurun_out/r01m_bench_bf16_reference8.json 2>/dev/null
python bench.py --workload overlap75 --no-cpu-baseline > gpurun_out/r01m_bench_bf16_overlap75.json 2>/dev/null
python bench.py --workload tta8 --steps 3 --no-cpu-baseline > gpurun_out/r01m_bench_bf16_tta8.json 2>/dev/null
python bench.py --precision fp32 --workload reference8 --steps 3 --no-cpu-baseline > gpurun_out/r01m_bench_fp32_reference8.json 2>/dev/null
python tools/volio_time.py > gpurun_out/r01m_volio_time.log 2>&1
python tools/stitch_time.py > gpurun_out/r01m_stitch_time.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r01m_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], round(d["value"],2), round(d["ms_per_step"],2), round(d["e2e"]["value"],2), round(d["roofline"]["frac"],3), round(d["roofline_accumulate"]["frac"],3), d["roofline_accumulate"].get("isolated",{}).get("gather_form_frac"))
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/r01m_volio_time.log
