#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_tc.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/r2c21_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c21_tests.log | cut -c1-400
for bal in 1 0; do MMR_TC_WGRAD_BAL=$bal timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c21_bench_bal$bal.json 2> gpurun_out/r2c21_bench_bal$bal.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c21_bench_bal$bal.json").read().strip().splitlines()[-1])
print("bal=$bal", d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
done
MMR_WGRAD_STREAM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c21_bench_1stream.json 2> gpurun_out/r2c21_bench_1stream.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c21_bench_1stream.json").read().strip().splitlines()[-1])
print("single stream", d["ms_per_step"])
PY
