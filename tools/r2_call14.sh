#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_routing.py > gpurun_out/r2c14_rt.log 2>&1; tail -2 gpurun_out/r2c14_rt.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c14_rt_launches.csv python tools/bench_routing.py --iters 3 > gpurun_out/r2c14_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c14_rt_launches.csv 2>/dev/null | head -40
MMR_RT_SPLIT=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c14_rt_launches_old.csv python tools/bench_routing.py --iters 3 > gpurun_out/r2c14_ncu_old.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c14_rt_launches_old.csv 2>/dev/null | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c14_rt_launches_mort.csv python tools/bench_routing.py --iters 3 --B 8192 --K 2 --variant mort > gpurun_out/r2c14_ncu_mort.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c14_rt_launches_mort.csv 2>/dev/null | head -30
# attention backward: 3 warps per head
for w in 2 3; do MMR_ATTN_BWD_WPH=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c14_bench_wph$w.json 2> gpurun_out/r2c14_bench_wph$w.err; python - <<PY
import json
d = json.loads(open("gpurun_out/r2c14_bench_wph$w.json").read().strip().splitlines()[-1])
print("wph$w", d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d.get("kernel_time_ms_per_step", {}).items()})
PY
done
MMR_ATTN_BWD_WPH=3 timeout 600 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x > gpurun_out/r2c14_tests_wph3.log 2>&1; echo "wph3 tests rc=$?"; tail -3 gpurun_out/r2c14_tests_wph3.log | cut -c1-300
