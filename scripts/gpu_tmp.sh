#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 120 python bench.py --no-cpu > gpurun_out/r2ag_bench$i.json 2> gpurun_out/r2ag_bench$i.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2ag_bench$i.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['us_per_line_search_in_kernel'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2ag_bench$i.err').read()[-800:])
PY
done
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=1 timeout 300 python scripts/profile_relax.py 4000 400 2>&1 | tail -4 | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2ag_pytest.log 2>&1; tail -4 gpurun_out/r2ag_pytest.log
timeout 300 python bench.py --workload cfg5_4096x4096 --no-cpu --steps 10 --warmup 2 --prerelax 20000 > gpurun_out/r2ag_cfg5_single.json 2> gpurun_out/r2ag_cfg5_single.err; python -c "
import json; b=json.load(open('gpurun_out/r2ag_cfg5_single.json')); print({k:b[k] for k in ('value','us_per_line_search','line_searches_per_event')})"
timeout 120 python scripts/bench_sweep64.py 30 8 2>/dev/null | tail -1 | cut -c1-300
