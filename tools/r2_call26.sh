#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_partial_fusion.py -q -m gpu > gpurun_out/r2c26_tests.log 2>&1; echo "partial tests rc=$?"; tail -40 gpurun_out/r2c26_tests.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_producer_proj.py -q -m gpu > gpurun_out/r2c26_tests2.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/r2c26_tests2.log | cut -c1-300
