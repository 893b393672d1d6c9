#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "event_modes or batched or seeded or trajectory_golden or local_relax" > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
echo "== plain launch for 1 block"
KMC_ENQ_TIMING=1 timeout 300 python scripts/bench_sweep64.py 30 16 > gpurun_out/r2c_s64_plain.json 2> gpurun_out/r2c_s64_plain.err; tail -c 700 gpurun_out/r2c_s64_plain.json; sort gpurun_out/r2c_s64_plain.err | tail -3
echo "== cluster(1) launch for 1 block"
KMC_CLUSTER1=1 KMC_ENQ_TIMING=1 timeout 300 python scripts/bench_sweep64.py 30 16 > gpurun_out/r2c_s64_cl1.json 2> gpurun_out/r2c_s64_cl1.err; tail -c 700 gpurun_out/r2c_s64_cl1.json; sort gpurun_out/r2c_s64_cl1.err | tail -3
echo "== 32 connections"
CUDA_DEVICE_MAX_CONNECTIONS=32 KMC_ENQ_TIMING=1 timeout 300 python scripts/bench_sweep64.py 30 16 > gpurun_out/r2c_s64_conn.json 2> gpurun_out/r2c_s64_conn.err; tail -c 700 gpurun_out/r2c_s64_conn.json; sort gpurun_out/r2c_s64_conn.err | tail -3
