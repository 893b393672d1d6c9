#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_domain" > gpurun_out/r2m_pytest.log 2>&1; tail -25 gpurun_out/r2m_pytest.log | cut -c1-250
timeout 600 python scripts/bench_domain.py cfg5_4096x4096 20 6 > gpurun_out/r2m_domain1.json 2> gpurun_out/r2m_domain1.err; tail -c 2500 gpurun_out/r2m_domain1.json; tail -5 gpurun_out/r2m_domain1.err | cut -c1-300
