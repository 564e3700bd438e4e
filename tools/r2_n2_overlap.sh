#!/usr/bin/env bash
# N=2: gradient all-reduce after the backward vs overlapped per layer block, with an NCCL CTA budget (charged 2x)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # name, extra bench args, env...
  local name=$1; shift; local extra=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $extra > gpurun_out/r2z_n2_$name.json 2> gpurun_out/r2z_n2_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2z_n2_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:20s} {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} patients/s  e2e {d['e2e']['ms_per_step']:.3f}  {d['config'].get('grad_allreduce')}")
except Exception as e:
    print(f"{name:20s} failed: {e}")
PY
}
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_n2_n1.json 2> gpurun_out/r2z_n2_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2z_n2_n1.json').read().strip().splitlines()[-1]); print('N=1 same box', d['ms_per_step'])"
run default "" NCCL_DEBUG=WARN
run overlap "--overlap" NCCL_DEBUG=WARN
run overlap_cta4 "--overlap" NCCL_MAX_CTAS=4
run overlap_cta8 "--overlap" NCCL_MAX_CTAS=8
run default_cta8 "" NCCL_MAX_CTAS=8
