#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none -k regex:"ln_rows_fwd|embed_fwd|embed_bwd|pool_fwd|rowplan|unfold|ln_rows_bwd_kernel<__nv_bfloat16, 1>|colsum" -s 14 -c 14 -f -o gpurun_out/r2z_rows python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2z_rows.log 2>&1; echo "capture rc=$?"
ncu -i gpurun_out/r2z_rows.ncu-rep --page raw --csv > gpurun_out/r2z_rows.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2z_rows.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "launch__grid_size"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    print({w.split(".")[0][-30:]: r[i][:60] for w, i in idx})
PY
