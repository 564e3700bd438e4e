#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/r2c23_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2c23_tests.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c23_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
