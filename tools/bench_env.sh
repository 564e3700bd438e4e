#!/bin/bash
# usage: tools/bench_env.sh "VAR=a VAR=b ..."   -- runs bench.py once per env assignment and prints the step time + kernel classes
for kv in "$@"; do
  env $kv python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$kv', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernel_time_ms_per_step'].items()})"
done
