#!/usr/bin/env bash
# final 8-GPU lines (charged 8x): N=1 on the same box for the same-box ratio, then N=8 pheno512, N=2, N=8 mort8192
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() {
  python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:24s} {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} patients/s  e2e {d['e2e']['value']:.0f}  reasons={d['clocks']['reasons']} sm={d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"{name:24s} failed: {e}")
PY
}
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z_n1_samebox.json 2> gpurun_out/r2z_n1_samebox.err; show r2z_n1_samebox
run() {  # name, nproc, extra bench args
  local name=$1; local n=$2; shift; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  show $name
}
run r2z_n8_pheno512 8
run r2z_n2_pheno512 2
run r2z_n4_pheno512 4
run r2z_n8_mort8192 8 --config mort8192
