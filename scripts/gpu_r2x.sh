#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2x_pytest.log 2>&1; tail -6 gpurun_out/r2x_pytest.log
for i in 1 2; do
timeout 120 python bench.py --no-cpu > gpurun_out/r2x_bench$i.json 2> gpurun_out/r2x_bench$i.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2x_bench$i.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['us_per_line_search_in_kernel'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2x_bench$i.err').read()[-800:])
PY
done
KMC_STRESS_WORKLOAD=cfg5_4096x4096 KMC_STRESS_PRERELAX=20000 timeout 120 python scripts/stress_events.py 2 2>&1 | tail -2 | cut -c1-300
timeout 300 python bench.py --workload cfg5_4096x4096 --no-cpu --steps 10 --warmup 2 --prerelax 20000 > gpurun_out/r2x_cfg5_single.json 2> gpurun_out/r2x_cfg5_single.err; python -c "
import json; b=json.load(open('gpurun_out/r2x_cfg5_single.json')); print({k:b[k] for k in ('value','us_per_line_search','line_searches_per_event')}, b['prerelax'], b['kernel_families_ms'])"
timeout 900 python scripts/seeds_study.py 24 200 1e-3 1e-2 5e-2 > gpurun_out/r2x_seeds.json 2> gpurun_out/r2x_seeds.err; cut -c1-1800 gpurun_out/r2x_seeds.json; tail -3 gpurun_out/r2x_seeds.err | cut -c1-300
