#!/usr/bin/env bash
# ncu evidence for round 2 in ONE gpurun call (1 GPU; run only after `python bench.py` exited 0 without ncu):
#
#   gpurun --timeout 600 -- 'bash tools/round2_profile.sh [ENV=VALUE ...]'      e.g.  MMR_TC_WS=1
#
# 1. launch list of two eagerly issued steps (per-kernel time shares)  -> gpurun_out/r2_launches.csv
# 2. --set full capture of the GEMM launches of one step              -> gpurun_out/r2_gemm_step.ncu-rep
# Read them here with:
#   python tools/summarize_launches.py gpurun_out/r2_launches.csv > profiles/r2_launches.txt
#   ncu -i gpurun_out/r2_gemm_step.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|gpu__time_duration|sm__pipe_tensor'
set -u
for kv in "$@"; do export "$kv"; done
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 236 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 102 -c 34 -o gpurun_out/r2_gemm_step -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r2_* | head
