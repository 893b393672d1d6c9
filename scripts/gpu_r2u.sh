#!/bin/bash
export KMC_STRESS_WORKLOAD=cfg5_4096x4096 KMC_STRESS_PRERELAX=20000
timeout 120 python scripts/stress_events.py 1 dbg14 2>&1 | grep -v "^  File\|^    " | grep -v "mismatches 0 " | tail -6 | cut -c1-400
