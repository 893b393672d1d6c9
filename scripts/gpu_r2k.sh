#!/bin/bash
set -u
mkdir -p gpurun_out
ok=0; bad=0
for i in $(seq 1 14); do
  timeout 90 python bench.py --no-cpu > gpurun_out/r2k_b$i.json 2> gpurun_out/r2k_b$i.err; rc=$?
  if [ $rc -eq 0 ]; then ok=$((ok+1)); else bad=$((bad+1)); echo "run $i rc=$rc"; tail -3 gpurun_out/r2k_b$i.err | cut -c1-300; fi
done
echo "ok=$ok bad=$bad"
python - <<'PY'
import json, glob
v = []
for f in sorted(glob.glob('gpurun_out/r2k_b*.json')):
    try: v.append(json.load(open(f))['value'])
    except Exception: pass
print(sorted(v))
PY
