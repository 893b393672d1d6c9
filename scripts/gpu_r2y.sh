#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent_cell or incremental" > gpurun_out/r2y_cells.log 2>&1; tail -25 gpurun_out/r2y_cells.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2y_pytest.log 2>&1; tail -6 gpurun_out/r2y_pytest.log
timeout 120 python bench.py --no-cpu > gpurun_out/r2y_bench1.json 2> gpurun_out/r2y_bench1.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2y_bench1.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['traffic'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2y_bench1.err').read()[-800:])
PY
timeout 900 python scripts/seeds_study.py 32 200 1e-3 1e-2 > gpurun_out/r2y_seeds.json 2> gpurun_out/r2y_seeds.err; cut -c1-1500 gpurun_out/r2y_seeds.json; tail -3 gpurun_out/r2y_seeds.err | cut -c1-300
