#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2ab_pytest.log 2>&1; tail -4 gpurun_out/r2ab_pytest.log
timeout 120 python bench.py --no-cpu > gpurun_out/r2ab_bench1.json 2> gpurun_out/r2ab_bench1.err; python - <<PY
import json
try:
    b = json.load(open('gpurun_out/r2ab_bench1.json'))
    print({k: b[k] for k in ('value', 'us_per_line_search', 'line_searches_per_event', 'gpu_launches')}, b['e2e']['value'], b['roofline']['us_per_line_search_in_kernel'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2ab_bench1.err').read()[-800:])
PY
timeout 600 python scripts/bench_sweeps.py > gpurun_out/r2ab_sweeps.jsonl 2> gpurun_out/r2ab_sweeps.err; python - <<'PY'
import json
for l in open('gpurun_out/r2ab_sweeps.jsonl'):
    try: r=json.loads(l)
    except Exception: continue
    print(r['workload'], r['precision'], {k:(round(r[k]['us_per_launch'],1), '%.3g'%r[k].get('pair_evals_per_s',0)) for k in ('force_sweep','line_energy_20_candidates','rate_table','select')})
PY
