#!/usr/bin/env bash
# round 2, GPU call 2: the whole -m gpu suite with the un-gated goldens, the default bench line (with the reference eager
# path on the same GPU), chain-vs-separate FFN timings, and --set full captures of the chain kernel and the fc1 GEMM.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c2_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2c2_tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c2_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2c2_smoke.log
echo "== bench default"; timeout 400 python bench.py > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2c2_bench.json
echo "== bench with the chained FFN"; MMR_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c2_bench_chain.json 2> gpurun_out/r2c2_bench_chain.err; cut -c1-300 gpurun_out/r2c2_bench_chain.json
echo "== chain micro"; timeout 120 python tools/bench_chain.py > gpurun_out/r2c2_chain_micro.log 2>&1; cat gpurun_out/r2c2_chain_micro.log
echo "== ncu chain"
ITERS=2 timeout 120 python tools/bench_chain.py > gpurun_out/r2c2_chain_plain.log 2>&1 && \
ITERS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_tc_kernel -s 1 -c 1 -o gpurun_out/r2c2_chain -f \
  python tools/bench_chain.py > gpurun_out/r2c2_ncu_chain.log 2>&1; echo "rc=$?"
echo "== ncu fc1 / fc2 gemm"
ITERS=2 ONLY=1,2 CUBLAS=0 timeout 120 python tools/bench_gemm.py > gpurun_out/r2c2_gemm_plain.log 2>&1 && \
ITERS=2 ONLY=1,2 CUBLAS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -o gpurun_out/r2c2_gemm -f \
  python tools/bench_gemm.py > gpurun_out/r2c2_ncu_gemm.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2c2_*
