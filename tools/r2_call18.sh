#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c18_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c18_tests.log | cut -c1-300
for sp in 1 0; do MMR_RT_SPLIT=$sp timeout 300 python tools/bench_routing.py --graph --iters 200 2>&1 | tail -1; done
for sp in 1 0; do MMR_RT_SPLIT=$sp timeout 300 python tools/bench_routing.py --graph --iters 50 --B 8192 --K 2 --variant mort 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c18_rt_launches.csv python tools/bench_routing.py --iters 3 > gpurun_out/r2c18_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c18_rt_launches.csv 2>/dev/null | grep -E "rs_|wgrad|routing|bias"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c18_bench.json").read().strip().splitlines()[-1])
print("new", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2c18_tests_all.log 2>&1; echo "all tests rc=$?"; tail -3 gpurun_out/r2c18_tests_all.log | cut -c1-300
