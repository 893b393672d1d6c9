"""diagnostics: histogram of per-atom displacements over a whole line search (KMC_DEBUG_E=7) on the bench trajectory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads, Growth
gv, sub, ad = workloads.make("cfg2_256x256_10k")
sim = Growth.Simulation(adatoms=ad, substrate=sub, params=runtime.params_from(gv))
os.environ["KMC_DEBUG_E"] = "7"
sys.stderr.write("prerelax\n")
st = sim.ctx.relax(sim.sbins, 16, 1e-4, 4000)
rng = np.random.default_rng(1234)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
us = rng.uniform(0, 1, (2 * 23, 2))
for s in range(n):
    r = sim.step(us[s, 0], us[s, 1], 0)
    sys.stderr.write("event %d kind %d ls %d\n" % (s, r["kind"], r["line_searches"]))
