#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_tc.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/r2c24_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c24_tests.log | cut -c1-400
for pp in 1 0 1 0; do MMR_ATTN_FWD_PIPE=$pp timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c24_bench_pipe$pp.json 2> gpurun_out/r2c24_bench_pipe$pp.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c24_bench_pipe$pp.json").read().strip().splitlines()[-1])
print("pipe=$pp", d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
done
