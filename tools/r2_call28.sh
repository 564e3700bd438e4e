#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_partial_fusion.py > gpurun_out/r2z_partial_fusion.json 2> gpurun_out/r2z_partial_fusion.err; echo "rc=$?"; tail -1 gpurun_out/r2z_partial_fusion.json; tail -3 gpurun_out/r2z_partial_fusion.err
