#!/usr/bin/env bash
# last N=1 lines of round 2 on the final build (references on)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py > gpurun_out/r2zz_bench_pheno512.json 2> gpurun_out/r2zz_bench_pheno512.err; echo "pheno512 rc=$?"
timeout 300 python bench.py --config inspect --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2zz_bench_inspect.json 2> gpurun_out/r2zz_bench_inspect.err; echo "inspect rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2zz_smoke.log
python - <<'PY'
import json
for n in ("pheno512", "inspect"):
    d = json.loads(open(f"gpurun_out/r2zz_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "routing ms", round(d["roofline_routing"]["ms_per_step"], 4), d["clocks"]["reasons"], d["gpu_launches"])
    if "reference_gpu_eager" in d and d["reference_gpu_eager"]:
        print("  gpu eager", d["reference_gpu_eager"]["value"], "x", d["reference_gpu_eager"]["speedup_device_resident"], d["reference_gpu_eager"]["speedup_e2e"], "cpu", d["cpu_baseline"]["value"])
PY
