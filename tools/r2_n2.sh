#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/check_dp.py > gpurun_out/r2_check_dp.log 2>&1; echo "check_dp rc=$?"; tail -6 gpurun_out/r2_check_dp.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference --batch 64 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_n2_ref.json
