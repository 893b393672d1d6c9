#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python bench.py --no-cpu --steps 10 --warmup 2 > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err; echo "plain rc=$?"
timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --no-cpu --steps 10 --warmup 2 > gpurun_out/r2i_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -v "^\[bench" gpurun_out/r2i_memcheck.log | head -60 | cut -c1-220
tail -5 gpurun_out/r2i_memcheck.log | cut -c1-300
