#!/bin/bash
# usage: tools/bench_batch.sh 64 128 256 ...   -- patients/s of the full step vs batch size
for b in "$@"; do
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('batch $b', round(d['value']), 'patients/s', round(d['ms_per_step'],3), 'ms')"
done
