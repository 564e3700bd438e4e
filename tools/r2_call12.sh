#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fusion.py tests/test_gpu_producer_proj.py -q -m gpu -x > gpurun_out/r2c12_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2c12_tests.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; cut -c1-300 gpurun_out/r2c12_bench.json
MMR_VARLEN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c12_bench_dense.json 2> gpurun_out/r2c12_bench_dense.err; cut -c1-300 gpurun_out/r2c12_bench_dense.json
