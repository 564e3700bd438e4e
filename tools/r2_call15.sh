#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/bench_routing.py > gpurun_out/r2c15_rt.log 2>&1; tail -1 gpurun_out/r2c15_rt.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c15_rt_launches.csv python tools/bench_routing.py --iters 3 > gpurun_out/r2c15_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c15_rt_launches.csv 2>/dev/null | grep -E "rs_|wgrad|routing|bias" 
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_ -s 10 -c 6 -o gpurun_out/r2c15_rs python tools/bench_routing.py --iters 3 > gpurun_out/r2c15_ncu_full.log 2>&1
ncu -i gpurun_out/r2c15_rs.ncu-rep --page raw --csv > gpurun_out/r2c15_rs_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2c15_rs_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "launch__grid_size", "launch__waves_per_multiprocessor", "launch__registers_per_thread", "sm__cycles_active.avg",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_st.sum", "smsp__inst_executed_op_local_ld.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i] for i in idx})
PY
