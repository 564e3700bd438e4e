#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_graph.py tests/test_gpu_train_steps.py tests/test_gpu_partial_fusion.py -q -m gpu -x > gpurun_out/r2c29_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c29_tests.log | cut -c1-300
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c29_bench_$i.json 2> gpurun_out/r2c29_bench_$i.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c29_bench_$i.json").read().strip().splitlines()[-1])
print("run $i", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
done
MMR_VARLEN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c29_bench_dense.json 2> gpurun_out/r2c29_bench_dense.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c29_bench_dense.json").read().strip().splitlines()[-1])
print("dense", d["ms_per_step"])
PY
