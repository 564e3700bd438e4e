#!/usr/bin/env bash
# --set full rows for the attention and LayerNorm kernels of the final build (one launch each, layer 1 of the second step)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none -k regex:"attn_fwd_single|attn_bwd_fused|ln_rows_fwd|ln_rows_bwd|wgrad_tc" -s 60 -c 12 -f -o gpurun_out/r2z_attn_ln python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2z_attn_ln.log 2>&1; echo "capture rc=$?"
ncu -i gpurun_out/r2z_attn_ln.ncu-rep --page raw --csv > gpurun_out/r2z_attn_ln.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2z_attn_ln.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    print({w.split(".")[0][-44:]: r[i] for w, i in idx})
PY
