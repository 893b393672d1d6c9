#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_domain.py cfg5_4096x4096 20 6 > gpurun_out/r2p_domain1.json 2> gpurun_out/r2p_domain1.err; python -c "
import json; b=json.load(open('gpurun_out/r2p_domain1.json')); print({k:b[k] for k in ('n_gpus','ms_per_sweep','halo_us','events_per_s','us_per_line_search')}, b['agreement_with_single_domain'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_domain.py cfg5_4096x4096 20 6 > gpurun_out/r2p_domain2.json 2> gpurun_out/r2p_domain2.err; python -c "
import json; b=json.loads(open('gpurun_out/r2p_domain2.json').read().strip().splitlines()[-1]); print({k:b[k] for k in ('n_gpus','ms_per_sweep','halo_us','events_per_s','us_per_line_search')}, b['agreement_with_single_domain'])"; tail -3 gpurun_out/r2p_domain2.err | cut -c1-300
timeout 300 python bench.py --workload cfg5_4096x4096 --no-cpu --steps 6 --warmup 2 --prerelax 20000 > gpurun_out/r2p_cfg5_single.json 2> gpurun_out/r2p_cfg5_single.err; tail -4 gpurun_out/r2p_cfg5_single.err | cut -c1-200; python -c "
import json; b=json.load(open('gpurun_out/r2p_cfg5_single.json')); print({k:b[k] for k in ('value','us_per_line_search','line_searches_per_event')}, b['prerelax'], b['kernel_families_ms'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "device_domain or line_search_candidates or bins_golden or vs_oracle" 2>&1 | tail -3
