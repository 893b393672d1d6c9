#!/bin/bash
export KMC_STRESS_WORKLOAD=cfg5_4096x4096 KMC_STRESS_PRERELAX=20000
for n in 8288 8140 8191; do
  echo "== n=$n"
  KMC_STRESS_N=$n timeout 120 python scripts/stress_events.py 1 2>&1 | grep -v "^  File\|^    " | tail -3 | cut -c1-300
done
export KMC_STRESS_WORKLOAD=cfg2_256x256_10k KMC_STRESS_PRERELAX=200000
for n in 9990 9950; do
  echo "== cfg2 n=$n"
  KMC_STRESS_N=$n timeout 120 python scripts/stress_events.py 3 2>&1 | grep -v "^  File\|^    " | tail -3 | cut -c1-300
done
