#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_tail.py tests/test_gpu_train_steps.py -q -m gpu > gpurun_out/r2c41_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2c41_tests.log | cut -c1-200
