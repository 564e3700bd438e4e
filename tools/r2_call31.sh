#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for mb in 3 4 5 3 4 5; do MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_ln$mb.so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c31_bench_ln$mb.json 2> gpurun_out/r2c31_bench_ln$mb.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c31_bench_ln$mb.json").read().strip().splitlines()[-1])
    print("ln minb=$mb", d["ms_per_step"], d["e2e"]["ms_per_step"], round(d["kernel_time_ms_per_step"]["fusion_bwd_call"]["ms_per_step"], 4))
except Exception as e:
    print("minb=$mb failed", e)
PY
done
