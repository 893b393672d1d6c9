#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "1 1" "8 1" "16 1" "32 1" "64 1" "16 2"; do
  set -- $cfg
  KMC_ENQ_TIMING=1 timeout 300 python scripts/bench_sweep64.py 30 $1 $2 > gpurun_out/r2l_s64_w$1_t$2.json 2> gpurun_out/r2l_s64_w$1_t$2.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2l_s64_w$1_t$2.json')); print($1, $2, round(b['events_per_s'], 1), round(b['seconds'], 3), b['host_seconds_rank0'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2l_s64_w$1_t$2.err').read()[-600:])
PY
done
timeout 300 python scripts/bench_sweep64.py 100 32 1 > gpurun_out/r2l_s64_100ev.json 2>/dev/null; cat gpurun_out/r2l_s64_100ev.json | cut -c1-500
