#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_packed_cache.py tests/test_gpu_graph.py tests/test_gpu_train_steps.py -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; cut -c1-300 gpurun_out/r2c8_bench.json
timeout 120 python examples/train_step.py 4 2>&1 | tail -3
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2c8_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:routing_ -s 4 -c 2 -o gpurun_out/r2c8_routing -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2c8_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2c8_ncu.log
