#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
M=1024 ITERS=2 timeout 60 python tools/bench_chain.py > gpurun_out/r2c4_chain_small.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c4_chain_small.log
timeout 120 python tools/bench_chain.py > gpurun_out/r2c4_chain_micro.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c4_chain_micro.log
echo "== chain + gemm parity"; timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu > gpurun_out/r2c4_tc_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2c4_tc_tests.log
echo "== bench: separate vs chain"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c4_bench_sep.json 2> gpurun_out/r2c4_bench_sep.err; cut -c1-300 gpurun_out/r2c4_bench_sep.json
MMR_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c4_bench_chain.json 2> gpurun_out/r2c4_bench_chain.err; cut -c1-300 gpurun_out/r2c4_bench_chain.json
echo "== gemm micro"; CUBLAS=0 timeout 120 python tools/bench_gemm.py > gpurun_out/r2c4_gemm.log 2>&1; cat gpurun_out/r2c4_gemm.log
echo "== inspect grads diag"; timeout 300 python tools/diag_inspect_grads.py > gpurun_out/r2c4_diag.log 2>&1; tail -8 gpurun_out/r2c4_diag.log
echo "== ncu chain"
ITERS=2 timeout 120 python tools/bench_chain.py > gpurun_out/r2c4_chain_plain.log 2>&1 && \
ITERS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_tc_kernel -s 1 -c 1 -o gpurun_out/r2c4_chain -f \
  python tools/bench_chain.py > gpurun_out/r2c4_ncu_chain.log 2>&1; echo "rc=$?"
