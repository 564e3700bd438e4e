#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c42_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2c42_tests.log | cut -c1-200
timeout 100 python tools/bench_routing.py --graph --iters 200 2>&1 | tail -1
timeout 100 python tools/bench_routing.py --graph --iters 50 --B 8192 --K 2 --variant mort 2>&1 | tail -1
