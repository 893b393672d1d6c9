#!/bin/bash
export KMC_STRESS_WORKLOAD=cfg5_4096x4096 KMC_STRESS_PRERELAX=20000
for v in "dbg10" "dbg11" "dbg12"; do
  echo "== variant '$v'"
  timeout 120 python scripts/stress_events.py 1 $v 2>&1 | grep -v "^  File\|^    " | grep -v "abort 0 " | tail -5 | cut -c1-300
done
