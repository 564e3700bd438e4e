#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_late.so timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/r2c38_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c38_tests.log | cut -c1-300
for v in base late base late; do MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_$v.so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c38_bench_$v.json 2> gpurun_out/r2c38_bench_$v.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c38_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("$v failed", e)
PY
done
