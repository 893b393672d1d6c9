#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_domain.py cfg5_4096x4096 20 6 > gpurun_out/r2n_domain1.json 2> gpurun_out/r2n_domain1.err; tail -c 1500 gpurun_out/r2n_domain1.json | cut -c1-1500; tail -3 gpurun_out/r2n_domain1.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_domain.py cfg5_4096x4096 20 6 > gpurun_out/r2n_domain2.json 2> gpurun_out/r2n_domain2.err; tail -c 1800 gpurun_out/r2n_domain2.json; tail -5 gpurun_out/r2n_domain2.err | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench2.json 2> gpurun_out/r2n_bench2.err; python - <<'PY'
import json
try:
    b = json.loads(open('gpurun_out/r2n_bench2.json').read().strip().splitlines()[-1])
    print({k: b[k] for k in ('value', 'n_gpus', 'us_per_line_search')}, 'cpu', b['cpu_baseline'] and b['cpu_baseline']['value'])
    print('sweep64', b['sweep64'])
    print('domain', b['domain_cfg5'])
except Exception as e:
    print('bench2 failed', e); print(open('gpurun_out/r2n_bench2.err').read()[-2000:])
PY
