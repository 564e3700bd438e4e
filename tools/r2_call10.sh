#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== dense plan (MMR_VARLEN=0)"; MMR_VARLEN=0 timeout 600 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c10_dense.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2c10_dense.log | cut -c1-300
echo "== packed"; timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c10_packed.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2c10_packed.log | cut -c1-300
echo "== bench dense vs packed"
MMR_VARLEN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c10_bench_dense.json 2> gpurun_out/r2c10_bench_dense.err; cut -c1-260 gpurun_out/r2c10_bench_dense.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c10_bench_packed.json 2> gpurun_out/r2c10_bench_packed.err; cut -c1-260 gpurun_out/r2c10_bench_packed.json
