#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c16_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c16_tests.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c16_rt_launches.csv python tools/bench_routing.py --iters 3 > gpurun_out/r2c16_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c16_rt_launches.csv 2>/dev/null | grep -E "rs_|wgrad|routing|bias"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c16_bench.json").read().strip().splitlines()[-1])
print("new", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
