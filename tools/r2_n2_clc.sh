#!/usr/bin/env bash
# N=2: does the dynamic (cluster launch control) GEMM tile scheduler make the overlapped all-reduce pay? (charged 2x)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # name, extra bench args, env...
  local name=$1; shift; local extra=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $extra > gpurun_out/r2z_n2clc_$name.json 2> gpurun_out/r2z_n2clc_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2z_n2clc_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:20s} {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} patients/s  e2e {d['e2e']['ms_per_step']:.3f}  {d['config'].get('grad_allreduce')}")
except Exception as e:
    print(f"{name:20s} failed: {e}")
PY
}
run static_after "" MMR_TC_CLC=0
run clc_after "" MMR_TC_CLC=1
run clc_overlap "--overlap" MMR_TC_CLC=1
run clc_overlap_cta8 "--overlap" MMR_TC_CLC=1 NCCL_MAX_CTAS=8
run static_overlap "--overlap" MMR_TC_CLC=0
