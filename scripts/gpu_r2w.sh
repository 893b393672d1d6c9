#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/seeds_study.py 32 200 1e-3 > gpurun_out/r2w_seeds.json 2> gpurun_out/r2w_seeds.err; tail -c 1500 gpurun_out/r2w_seeds.json; tail -3 gpurun_out/r2w_seeds.err | cut -c1-300
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2w_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2w_ncu_launches.log 2>&1
tail -2 gpurun_out/r2w_ncu_launches.log | cut -c1-200
python scripts/profile_relax.py 4000 200 > gpurun_out/r2w_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_relax2 -s 1 -c 1 -o gpurun_out/prof_relax2_r2w python scripts/profile_relax.py 4000 200 > gpurun_out/r2w_ncu_full.log 2>&1
tail -3 gpurun_out/r2w_ncu_full.log | cut -c1-200; ls -la gpurun_out/prof_relax2_r2w.ncu-rep
