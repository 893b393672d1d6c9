"""ncu target: a warm-up Relaxation (not profiled: use `-s 1`) followed by one bounded Relaxation in the
steady state of config 2 (few movers, rare rebuilds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmcislandgrowth_b200 import runtime, workloads, Growth
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gv, sub, ad = workloads.make("cfg2_256x256_10k")
sim = Growth.Simulation(adatoms=ad, substrate=sub, params=runtime.params_from(gv))
st = sim.ctx.relax(sim.sbins, 16, 1e-4, warm)
st = sim.ctx.relax(sim.sbins, 16, 1e-4, cap)
print("ls", st.line_searches, "restarts", st.restarts, "maxF", st.maxF)
