#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c37_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2c37_tests.log | cut -c1-300
for i in 1 2; do timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c37_bench_$i.json 2> gpurun_out/r2c37_bench_$i.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c37_bench_$i.json").read().strip().splitlines()[-1])
k = d["kernel_time_ms_per_step"]
print("run $i", d["ms_per_step"], d["e2e"]["ms_per_step"], "fwd/bwd call", round(k["fusion_fwd_call"]["ms_per_step"], 4), round(k["fusion_bwd_call"]["ms_per_step"], 4), d["gpu_launches"])
PY
done
timeout 200 python bench.py --config mort8192 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c37_bench_mort.json 2> gpurun_out/r2c37_bench_mort.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c37_bench_mort.json").read().strip().splitlines()[-1])
print("mort8192", d["ms_per_step"], d["value"])
PY
