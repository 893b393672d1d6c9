#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/debug_event_modes.py > gpurun_out/r2b_modes.log 2>&1; tail -30 gpurun_out/r2b_modes.log
for w in 1 4 16 64; do
  timeout 300 python scripts/bench_sweep64.py 30 $w > gpurun_out/r2b_sweep64_w$w.json 2> gpurun_out/r2b_sweep64_w$w.err
  tail -c 900 gpurun_out/r2b_sweep64_w$w.json; tail -c 300 gpurun_out/r2b_sweep64_w$w.err
done
KMC_DEBUG_E=8 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2b_blocks.log 2>&1; tail -6 gpurun_out/r2b_blocks.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=1 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2b_phase1.log 2>&1; tail -4 gpurun_out/r2b_phase1.log
KMC_PHASE_TIMING=1 KMC_PHASE_BLOCK=70 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2b_phase70.log 2>&1; tail -4 gpurun_out/r2b_phase70.log
