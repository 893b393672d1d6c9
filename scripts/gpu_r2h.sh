#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
timeout 120 python bench.py --no-cpu > gpurun_out/r2h_bench$i.json 2> gpurun_out/r2h_bench$i.err; echo "rc=$?"; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2h_bench$i.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['us_per_line_search_in_kernel'])
except Exception as e:
    print('bench failed', e)
PY
done
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=1 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2h_phase1.log 2>&1; tail -4 gpurun_out/r2h_phase1.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=70 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2h_phase70.log 2>&1; tail -4 gpurun_out/r2h_phase70.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=70 timeout 120 python bench.py --no-cpu --steps 6 2>&1 | grep "phase us/ls" | tail -8 | cut -c1-330
