#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2e_pytest.log 2>&1; tail -12 gpurun_out/r2e_pytest.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=1 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2e_phase1.log 2>&1; tail -4 gpurun_out/r2e_phase1.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=70 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2e_phase70.log 2>&1; tail -4 gpurun_out/r2e_phase70.log
KMC_DEBUG_E=8 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2e_blocks.log 2>&1; tail -3 gpurun_out/r2e_blocks.log | cut -c1-300
timeout 600 python bench.py --no-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    b = json.load(open('gpurun_out/r2e_bench.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['frac'], b['kernel_families_ms'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2e_bench.err').read()[-1500:])
PY
KMC_ENQ_TIMING=1 timeout 300 python scripts/bench_sweep64.py 30 16 > gpurun_out/r2e_s64_w16.json 2> gpurun_out/r2e_s64_w16.err; tail -c 700 gpurun_out/r2e_s64_w16.json; sort gpurun_out/r2e_s64_w16.err | tail -3
