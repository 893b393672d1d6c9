"""BASELINE configs[0]: the reference's default run (W=18, Hs=10, empty surface, reference quirk modes) on the GPU:
events/s until ~100 adatoms, next to the literal reference's 0.21 events/s (first 45 events)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import Growth, _lib
nev = int(sys.argv[1]) if len(sys.argv) > 1 else 150
mapped = len(sys.argv) > 2 and sys.argv[2] == "mapped"
if mapped:
    from kmcislandgrowth_b200 import Constants
    Constants.hop_index_mode = 1; Constants.deposit_mode = 1
sim = Growth.Simulation()
rng = np.random.default_rng(1)
t0 = time.perf_counter(); marks = {}; ls = 0; err = None
for t in range(nev):
    try:
        rec = sim.step(rng.uniform(), rng.uniform())
    except _lib.KmcError as e:
        err = "event %d: %s" % (t, e); break
    ls += rec["line_searches"]
    if t + 1 in (45, 100, nev): marks[t + 1] = time.perf_counter() - t0
n = sim.ctx.state_count()
print(json.dumps({"events": t + 1, "adatoms": n, "seconds_at": marks, "events_per_s_first_45": 45 / marks[45] if 45 in marks else None,
                  "events_per_s_all": (t + 1) / (time.perf_counter() - t0), "line_searches": ls,
                  "us_per_line_search": (time.perf_counter() - t0) / max(ls, 1) * 1e6, "modes": "mapped" if mapped else "reference quirks", "stopped": err}))
