#!/bin/bash
# The measurement recipe of this repository on one B200 (run it through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh TAG'
# parity tests -> bench (twice) -> ncu launch list of the bench -> one full ncu capture of k_relax2.  Every ncu pass runs
# only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
TAG=${1:-check}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for i in 1 2; do
    timeout 300 python bench.py > gpurun_out/${TAG}_bench$i.json 2> gpurun_out/${TAG}_bench$i.err; tail -c 600 gpurun_out/${TAG}_bench$i.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/profile_relax.py 4000 200 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_relax2 -s 1 -c 1 -o gpurun_out/prof_relax2_${TAG} \
    python scripts/profile_relax.py 4000 200 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/prof_relax2_${TAG}.ncu-rep
