#!/usr/bin/env bash
# last validation of the final sources: the whole GPU suite and smoke()
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2zz_tests_final.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r2zz_tests_final.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2zz_smoke2.log
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2zz_bench_last.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2zz_bench_last.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
