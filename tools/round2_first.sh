#!/usr/bin/env bash
# First GPU actions of round 2 (one gpurun call, ~2 GPU-minutes): everything that was written after round 1's GPU budget
# was spent and is therefore still unverified on hardware.
#
#   gpurun --timeout 300 -- 'bash tools/round2_first.sh'
#
# 1. weight-stationary tcgen05 GEMM variant (MMR_TC_WS=1): parity of the GEMM unit tests + the bf16 golden cases, then the
#    A/B of the micro-benchmark and of the whole step;
# 2. the long-sequence reference goldens on the GPU (MMR_TEST_LONG_GOLDEN=1).
# Results land in gpurun_out/r2_first_*.log.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== WS parity"; MMR_TC_WS=1 timeout 120 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2_first_ws_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_first_ws_tests.log
echo "== GEMM micro-benchmark, default vs WS"
timeout 60 python tools/bench_gemm.py > gpurun_out/r2_first_gemm_default.log 2>&1; tail -12 gpurun_out/r2_first_gemm_default.log
MMR_TC_WS=1 timeout 60 python tools/bench_gemm.py > gpurun_out/r2_first_gemm_ws.log 2>&1; tail -12 gpurun_out/r2_first_gemm_ws.log
echo "== step, default vs WS"
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_first_bench_default.json 2> gpurun_out/r2_first_bench_default.err; cut -c1-260 gpurun_out/r2_first_bench_default.json
MMR_TC_WS=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_first_bench_ws.json 2> gpurun_out/r2_first_bench_ws.err; cut -c1-260 gpurun_out/r2_first_bench_ws.json
echo "== step with the device-side loss tail"
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --fused-loss > gpurun_out/r2_first_bench_fusedloss.json 2> gpurun_out/r2_first_bench_fusedloss.err; cut -c1-260 gpurun_out/r2_first_bench_fusedloss.json
echo "== tcgen05 attention with per-patient 3-D tensor maps (zero fill past a patient's last token)"
MMR_ATTN_TC_MAP3D=1 timeout 120 python -m pytest tests/test_gpu_fusion.py -q -m gpu -k tcgen05 > gpurun_out/r2_first_attn_map3d.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_first_attn_map3d.log
echo "== tcgen05 attention with two K / V stages (prefetch), then the engine comparison incl. the variants"
MMR_ATTN_TC_PREFETCH=1 timeout 120 python -m pytest tests/test_gpu_fusion.py -q -m gpu -k tcgen05 > gpurun_out/r2_first_attn_prefetch.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_first_attn_prefetch.log
timeout 90 python tools/bench_attn_engines.py --quick --variants > gpurun_out/r2_first_attn_engines.jsonl 2> gpurun_out/r2_first_attn_engines.err; cat gpurun_out/r2_first_attn_engines.jsonl
echo "== the integrated training-step example (graph capture incl. the loss tail)"
timeout 90 python examples/train_step.py 6 > gpurun_out/r2_first_train_step.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_first_train_step.log
echo "== long-sequence goldens on the GPU"
MMR_TEST_LONG_GOLDEN=1 timeout 120 python -m pytest tests/test_gpu_fusion.py -q -m gpu -k long > gpurun_out/r2_first_long.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_first_long.log
