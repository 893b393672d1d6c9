#!/bin/bash
KMC_SYNC_DEBUG=1 timeout 300 python bench.py --workload cfg5_4096x4096 --no-cpu --steps 6 --warmup 2 --prerelax 20000 2>&1 | grep -v "^  File\|^    " | head -20 | cut -c1-300
