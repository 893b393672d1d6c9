#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
timeout 120 python bench.py --no-cpu > gpurun_out/r2g_bench$i.json 2> gpurun_out/r2g_bench$i.err; echo "rc=$?"; tail -6 gpurun_out/r2g_bench$i.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2g_bench$i.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['frac'], b['kernel_families_ms'])
except Exception as e:
    print('bench failed', e)
PY
done
timeout 120 python bench.py --no-cpu --no-flush > gpurun_out/r2g_bench3.json 2> gpurun_out/r2g_bench3.err; echo "rc=$?"; tail -4 gpurun_out/r2g_bench3.err
nvidia-smi --query-gpu=name,memory.used --format=csv
