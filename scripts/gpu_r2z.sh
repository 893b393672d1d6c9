#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent_cell or incremental" > gpurun_out/r2z_cells.log 2>&1; tail -25 gpurun_out/r2z_cells.log
