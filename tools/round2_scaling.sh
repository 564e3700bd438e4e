#!/usr/bin/env bash
# 8-GPU all-reduce experiments for round 2 (charged 8x: keep it short).  The step is 5.34 ms at N=1 and 5.87 ms at N=8: the
# 0.5 ms is the NCCL AVG all-reduce of 79 MB of fp32 gradients issued after the backward; overlapping it was equal in round 1
# because NCCL's CTAs compete with the persistent one-CTA-per-SM GEMMs.  Things not yet tried: NVLS (in-switch reduction:
# few CTAs, multimem) with a small CTA budget, so that the overlapped variant (--overlap) stops stealing SMs.
#
#   gpurun --gpus 8 --timeout 400 -- 'bash tools/round2_scaling.sh'
set -u
mkdir -p gpurun_out
run() {  # name, extra bench args, env...
  local name=$1; shift; local extra=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline $extra > gpurun_out/r2_scale_$name.json 2> gpurun_out/r2_scale_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2_scale_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:28s} {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} patients/s  reasons={d['clocks']['reasons']}")
except Exception as e:
    print(f"{name:28s} failed: {e}")
PY
}
run default        ""          NCCL_DEBUG=WARN
run nvls           ""          NCCL_ALGO=NVLS
run nvls_cta8      ""          NCCL_ALGO=NVLS NCCL_MAX_CTAS=8
run overlap        "--overlap" NCCL_DEBUG=WARN
run overlap_nvls8  "--overlap" NCCL_ALGO=NVLS NCCL_MAX_CTAS=8
run overlap_cta4   "--overlap" NCCL_MAX_CTAS=4
