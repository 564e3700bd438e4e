#!/usr/bin/env bash
# final build: the whole GPU suite, smoke(), the default bench line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests_final.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2z_tests_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2z_smoke.log
timeout 600 python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_bench_default.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "routing", d["roofline_routing"]["ms_per_step"], d["clocks"], "launches", d["gpu_launches"])
print("gpu eager ref", d.get("reference_gpu_eager", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
