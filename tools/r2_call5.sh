#!/usr/bin/env bash
# bottleneck experiments on the persistent tcgen05 GEMM: which part of a tile's life bounds the K = 256 shapes?
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 CUBLAS=0 ONLY=0,1,2
run() { echo "--- $*"; env "$@" timeout 100 python tools/bench_gemm.py 2>&1 | grep -v "rel err"; }
{
run MMR_TC_DBG=0
run MMR_TC_DBG=2
run MMR_TC_DBG=1
run MMR_TC_DBG=4
run MMR_TC_DBG=0 MMR_TC_PAIR_MIN_K=256
run MMR_TC_DBG=4 MMR_TC_PAIR_MIN_K=256
run MMR_TC_DBG=1 MMR_TC_PAIR_MIN_K=256
run MMR_TC_DBG=0 MMR_TC_STAGES=2
run MMR_TC_DBG=4 MMR_TC_STAGES=2
run MMR_TC_DBG=4 MMR_TC_PAIR=0
} > gpurun_out/r2c5_gemm_experiments.log 2>&1
cat gpurun_out/r2c5_gemm_experiments.log
