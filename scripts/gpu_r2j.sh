#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "" "notiming" "noflush" "ownstream" "nofast" "hostmode"; do
  timeout 200 python scripts/stress_events.py 60 $v 2>&1 | tail -2 | cut -c1-300
done
