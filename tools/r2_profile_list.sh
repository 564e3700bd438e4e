#!/usr/bin/env bash
# launch list of two eagerly issued steps (per-kernel time shares) -> gpurun_out/r2_launches.csv
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_all.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2_ncu_list.log 2>&1; echo "launch list rc=$?"
tail -2 gpurun_out/r2_ncu_list.log
