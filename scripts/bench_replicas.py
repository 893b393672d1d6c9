"""BASELINE configs[3]: temperature x flux sweep of independent replica trajectories (default geometry), this rank's share
run sequentially and with N concurrent host threads on one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmcislandgrowth_b200 import replicas
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 40
grid = replicas.sweep_grid()[::max(1, 64 // nrep)][:nrep]
out = {}
for threads in (1, 4, 8):
    t0 = time.perf_counter()
    res = replicas.run_sweep(grid, nev, device=0, threads=threads, max_line_searches=20000)
    dt = time.perf_counter() - t0
    out["threads_%d" % threads] = {"seconds": round(dt, 3), "events_per_s": nrep * nev / dt, "line_searches": sum(r["line_searches"] for r in res)}
    summary = [(r["replica"], r["n_adatoms"], r["hops"], r["line_searches"]) for r in res]
    if threads == 1:
        first = summary
    out["identical_results"] = (summary == first)
print(json.dumps({"replicas": nrep, "events_each": nev, **out}))
