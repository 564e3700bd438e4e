#!/usr/bin/env bash
# final round-2 evidence on one B200: bench lines (3 configs, references on), launch list of the eagerly issued step,
# --set full capture of the split routing kernels (fwd + bwd) at B=512 K=25 and B=8192 K=2
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_pheno512.json 2> gpurun_out/r2z_bench_pheno512.err; echo "pheno512 rc=$?"; cut -c1-300 gpurun_out/r2z_bench_pheno512.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r2z_bench_reference.json
timeout 300 python bench.py --config mort8192 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_bench_mort8192.json 2> gpurun_out/r2z_bench_mort8192.err; echo "mort rc=$?"
timeout 300 python bench.py --config inspect --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_bench_inspect.json 2> gpurun_out/r2z_bench_inspect.err; echo "inspect rc=$?"
python - <<'PY'
import json
for n in ("pheno512", "mort8192", "inspect"):
    try:
        d = json.loads(open(f"gpurun_out/r2z_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline_routing"]["ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "ERR", e)
PY
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2z_plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2z_ncu_list.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2z_launches.csv > gpurun_out/r2z_launches.txt 2>/dev/null; head -30 gpurun_out/r2z_launches.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_ -s 8 -c 6 -f -o gpurun_out/r2z_rs python tools/bench_routing.py --iters 3 > gpurun_out/r2z_ncu_rs.log 2>&1; echo "rs capture rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:rs_ -s 8 -c 6 -f -o gpurun_out/r2z_rs_mort python tools/bench_routing.py --iters 3 --B 8192 --K 2 --variant mort > gpurun_out/r2z_ncu_rs_mort.log 2>&1; echo "rs mort capture rc=$?"
for f in r2z_rs r2z_rs_mort; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null; done
for sp in 1 0; do MMR_RT_SPLIT=$sp timeout 300 python tools/bench_routing.py --graph --iters 200 2>&1 | tail -1; MMR_RT_SPLIT=$sp timeout 300 python tools/bench_routing.py --graph --iters 50 --B 8192 --K 2 --variant mort 2>&1 | tail -1; done
