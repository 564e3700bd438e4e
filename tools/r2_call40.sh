#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_emb.so timeout 600 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c40_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2c40_tests.log | cut -c1-200
for v in base emb base emb; do MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_$v.so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c40_bench_$v.json 2> gpurun_out/r2c40_bench_$v.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c40_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", d["ms_per_step"], d["e2e"]["ms_per_step"], round(d["kernel_time_ms_per_step"]["fusion_fwd_call"]["ms_per_step"], 4))
except Exception as e:
    print("$v failed", e)
PY
done
