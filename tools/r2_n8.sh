#!/usr/bin/env bash
# 8-GPU lines for round 2 (charged 8x: four short runs): the headline config, the overlapped all-reduce, and the two other
# multi-GPU configs of BASELINE.json (configs[2] Mort global batch 8192 strong scaling, configs[4] INSPECT shapes 256/GPU).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # name, extra bench args
  local name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_n8_$name.json 2> gpurun_out/r2_n8_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_n8_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:12s} {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} patients/s  e2e {d['e2e']['value']:.0f}  reasons={d['clocks']['reasons']} sm={d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"{name:12s} failed: {e}")
PY
}
run pheno512
run overlap --overlap
run mort8192 --config mort8192
run inspect --config inspect --steps 10
