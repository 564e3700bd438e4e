#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 CUBLAS=0
run() { echo "--- $*"; env "$@" timeout 100 python tools/bench_gemm.py 2>&1 | grep -v "rel err"; }
{
run MMR_TC_STAGES=8
run MMR_TC_STAGES=4
run MMR_TC_STAGES=3
} > gpurun_out/r2c7_gemm_stages.log 2>&1
cat gpurun_out/r2c7_gemm_stages.log
timeout 100 python tools/bench_gemm.py 2>&1 | grep "rel err"
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; cut -c1-300 gpurun_out/r2c7_bench.json
