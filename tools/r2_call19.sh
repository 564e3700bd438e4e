#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c19_tests_all.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r2c19_tests_all.log | cut -c1-400
for tf in 1 0; do MMR_TF32=$tf timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c19_bench_tf$tf.json 2> gpurun_out/r2c19_bench_tf$tf.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c19_bench_tf$tf.json").read().strip().splitlines()[-1])
print("tf32=$tf", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
done
