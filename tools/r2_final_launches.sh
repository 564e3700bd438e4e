#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2zz_plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2zz_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2zz_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2zz_launches.csv > gpurun_out/r2zz_launches.txt 2>/dev/null; head -24 gpurun_out/r2zz_launches.txt; tail -1 gpurun_out/r2zz_launches.txt
