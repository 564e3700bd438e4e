#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/r2c32_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c32_tests.log | cut -c1-300
for mb in 3 4 3 4; do MMR_B200_LIB=$PWD/multimodalrouting_b200/csrc/ab/libmmr_lnh$mb.so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c32_bench_lnh$mb.json 2> gpurun_out/r2c32_bench_lnh$mb.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c32_bench_lnh$mb.json").read().strip().splitlines()[-1])
    print("hoisted loads, minb=$mb", d["ms_per_step"], d["e2e"]["ms_per_step"], round(d["kernel_time_ms_per_step"]["fusion_bwd_call"]["ms_per_step"], 4))
except Exception as e:
    print("minb=$mb failed", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ln_rows_bwd -s 20 -c 4 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph 2>/dev/null | grep -E "ln_rows_bwd" | cut -d, -f5,12-15 | head -12
