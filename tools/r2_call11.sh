#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c11_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c11_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; cut -c1-300 gpurun_out/r2c11_bench.json
timeout 300 python bench.py --config mort8192 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c11_bench_mort.json 2> gpurun_out/r2c11_bench_mort.err; cut -c1-300 gpurun_out/r2c11_bench_mort.json
timeout 300 python bench.py --config inspect --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c11_bench_inspect.json 2> gpurun_out/r2c11_bench_inspect.err; cut -c1-300 gpurun_out/r2c11_bench_inspect.json
timeout 120 python examples/train_step.py 4 2>&1 | tail -2
