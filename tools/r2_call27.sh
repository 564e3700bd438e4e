#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_partial_fusion.py > gpurun_out/r2z_partial_fusion.json 2> gpurun_out/r2z_partial_fusion.err; echo "rc=$?"; tail -1 gpurun_out/r2z_partial_fusion.json; tail -3 gpurun_out/r2z_partial_fusion.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_partial_launches.csv python tools/bench_partial_fusion.py --iters 2 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2z_partial_launches.csv 2>/dev/null | head -24
