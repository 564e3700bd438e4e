#!/usr/bin/env bash
# round 2, GPU call 3: pipelined chain kernel -- micro check/timing first (bounded), then parity, then the step A/B,
# then the whole -m gpu suite.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== chain micro (small M first: a protocol bug must not eat the budget)"
M=1024 ITERS=2 timeout 60 python tools/bench_chain.py > gpurun_out/r2c3_chain_small.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c3_chain_small.log
timeout 120 python tools/bench_chain.py > gpurun_out/r2c3_chain_micro.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c3_chain_micro.log
MMR_CHAIN_PAIR=0 timeout 120 python tools/bench_chain.py > gpurun_out/r2c3_chain_micro_cl1.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c3_chain_micro_cl1.log
echo "== chain parity"; timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu > gpurun_out/r2c3_tc_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2c3_tc_tests.log
echo "== bench: separate vs chain"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c3_bench_sep.json 2> gpurun_out/r2c3_bench_sep.err; cut -c1-300 gpurun_out/r2c3_bench_sep.json
MMR_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c3_bench_chain.json 2> gpurun_out/r2c3_bench_chain.err; cut -c1-300 gpurun_out/r2c3_bench_chain.json
echo "== other configs"
timeout 400 python bench.py --config mort8192 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_mort8192.json 2> gpurun_out/r2c3_bench_mort8192.err; cut -c1-300 gpurun_out/r2c3_bench_mort8192.json
timeout 400 python bench.py --config inspect --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_inspect.json 2> gpurun_out/r2c3_bench_inspect.err; cut -c1-300 gpurun_out/r2c3_bench_inspect.json
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c3_tests.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2c3_tests.log
