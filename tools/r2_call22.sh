#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for pr in 0 1 0 1; do MMR_GRAPH_PRIORITY=$pr timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c22_bench_pr$pr.json 2> gpurun_out/r2c22_bench_pr$pr.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c22_bench_pr$pr.json").read().strip().splitlines()[-1])
print("prio=$pr", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
done
MMR_GRAPH_PRIORITY=1 MMR_TC_WGRAD_BAL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c22_bench_pr1bal.json 2> gpurun_out/r2c22_bench_pr1bal.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c22_bench_pr1bal.json").read().strip().splitlines()[-1])
print("prio=1 bal=1", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
