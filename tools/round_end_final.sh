T=r02p
python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_f16x3_overlap50.json
python tools/op_time.py x3 > gpurun_out/${T}_op_time_x3.log 2>&1
python tools/op_time.py > gpurun_out/${T}_op_time_bf16.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench_f16x3_overlap50.json").read().strip().splitlines()[-1])
print(round(d["value"], 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 2), round(d["roofline"]["frac"], 3), d["bf16"]["value"], d["parity"]["label_flip_frac"], d["roofline"]["share_of_step"], d["model_tflops"])
PY
