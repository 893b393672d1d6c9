#!/bin/bash
# round 2, GPU call A: fixture of the new bench protocol, the whole GPU test tier, first bench line, phase-timing baseline
set -u
mkdir -p gpurun_out
echo "== fixture" ; date
timeout 900 python bench.py --record-fixture --no-cpu > gpurun_out/r2a_fixture.log 2> gpurun_out/r2a_fixture.err
tail -c 600 gpurun_out/r2a_fixture.err
if [ -s gpurun_out/bench_cfg2_256x256_10k.json ]; then cp gpurun_out/bench_cfg2_256x256_10k.json tests/golden/bench_cfg2_256x256_10k.json; fi
echo "== pytest gpu" ; date
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --durations=15 > gpurun_out/r2a_pytest.log 2>&1
tail -40 gpurun_out/r2a_pytest.log
echo "== bench" ; date
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json ; tail -c 500 gpurun_out/r2a_bench.err
echo "== phase timing" ; date
KMC_PHASE_TIMING=1 timeout 300 python scripts/profile_relax.py 4000 400 > gpurun_out/r2a_phase.log 2>&1
tail -5 gpurun_out/r2a_phase.log
echo "== sweep64" ; date
timeout 600 python scripts/bench_sweep64.py 30 16 > gpurun_out/r2a_sweep64.json 2> gpurun_out/r2a_sweep64.err
tail -c 800 gpurun_out/r2a_sweep64.json ; tail -c 500 gpurun_out/r2a_sweep64.err
date
