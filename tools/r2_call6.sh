#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in pheno_inspect pheno_sharp4 pheno_missing mort_missing pheno_tl256; do
  echo "=== $n"; timeout 300 python tools/diag_inspect_grads.py $n 2>&1 | grep -v Warning | tail -10
done > gpurun_out/r2c6_diag.log 2>&1
cat gpurun_out/r2c6_diag.log | cut -c1-420
echo "== bench (pairs for K=256)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; cut -c1-300 gpurun_out/r2c6_bench.json
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c6_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2c6_tests.log
