#!/usr/bin/env bash
# raw gradient all-reduce of the path's flat buffers at N=8 under NCCL settings (charged 8x: short)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift
  echo "== $name"; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/bench_allreduce.py 2>&1 | grep -E "us per step|NVLS|error|Error" | tail -6
}
run default NCCL_DEBUG=WARN
run nvls NCCL_ALGO=NVLS
run tree NCCL_ALGO=Tree
run ring NCCL_ALGO=Ring
run default_ctas32 NCCL_MIN_CTAS=32
