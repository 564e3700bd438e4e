#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MMR_TC_CLC=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x > gpurun_out/r2c25_tests_tc.log 2>&1; echo "tc tests (clc) rc=$?"; tail -5 gpurun_out/r2c25_tests_tc.log | cut -c1-300
MMR_TC_CLC=1 timeout 400 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/r2c25_tests_fusion.log 2>&1; echo "fusion tests (clc) rc=$?"; tail -3 gpurun_out/r2c25_tests_fusion.log | cut -c1-300
for c in 1 0 1 0; do MMR_TC_CLC=$c timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c25_bench_clc$c.json 2> gpurun_out/r2c25_bench_clc$c.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c25_bench_clc$c.json").read().strip().splitlines()[-1])
    print("clc=$c", d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
except Exception as e:
    print("clc=$c failed", e)
PY
done
for c in 1 0; do MMR_TC_CLC=$c MMR_TC_WGRAD_BAL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c25_bench_clc${c}_bal.json 2> gpurun_out/r2c25_bench_clc${c}_bal.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c25_bench_clc${c}_bal.json").read().strip().splitlines()[-1])
    print("clc=$c bal=1", d["ms_per_step"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("clc=$c bal failed", e)
PY
done
