#!/usr/bin/env bash
# round-2 ncu evidence for the final build: launch list of the eagerly issued steps + --set full capture of the GEMM launches
# of one step (dram bytes per launch for roofline.traffic, tensor-pipe activity).
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2f_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2f_ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2f_plain2.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 102 -c 34 -o gpurun_out/r2f_gemm_step -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2f_ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r2f_*
