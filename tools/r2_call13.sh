#!/usr/bin/env bash
# split routing path: parity (routing tests first, then the whole GPU suite), then the bench with both paths
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/r2c13_tests_rt.log 2>&1; echo "rt rc=$?"; tail -15 gpurun_out/r2c13_tests_rt.log | cut -c1-400
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c13_tests.log 2>&1; echo "all rc=$?"; tail -15 gpurun_out/r2c13_tests.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; cut -c1-200 gpurun_out/r2c13_bench.json; python - <<'PY'
import json
for f in ("gpurun_out/r2c13_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("roofline_routing"))
    except Exception as e:
        print(f, "ERR", e)
PY
MMR_RT_SPLIT=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c13_bench_old.json 2> gpurun_out/r2c13_bench_old.err; python - <<'PY'
import json
for f in ("gpurun_out/r2c13_bench_old.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("roofline_routing"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 300 python bench.py --config mort8192 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c13_bench_mort.json 2> gpurun_out/r2c13_bench_mort.err; python - <<'PY'
import json
for f in ("gpurun_out/r2c13_bench_mort.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d.get("roofline_routing"))
    except Exception as e:
        print(f, "ERR", e)
PY
