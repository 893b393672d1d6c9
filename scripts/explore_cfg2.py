import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads, Growth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_256x256_10k"
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
gv, sub, ad = workloads.make(name)
p = runtime.params_from(gv)
sim = Growth.Simulation(adatoms=ad, substrate=sub, params=p)
ctx = sim.ctx
t0 = time.time()
st = ctx.relax(sim.sbins, 16, 1e-4, cap)
ctx.synchronize()
dt = time.time() - t0
print("pre-relax: ls=%d restarts=%d maxF=%.3e status=%d scale=%d thr=%g  %.3fs  (%.1f us/ls) launches=%d" % (st.line_searches, st.restarts, st.maxF, st.status, st.final_scale, st.final_threshold, dt, dt/max(st.line_searches,1)*1e6, ctx.launch_count()))
rng = np.random.default_rng(11)
for t in range(8):
    u0, u1 = rng.uniform(), rng.uniform()
    t0 = time.time()
    rec = sim.step(u0, u1, max_line_searches=cap)
    dt = time.time() - t0
    print("event", rec, "%.3fs" % dt, flush=True)
r = ctx.rates(sim.adatoms, sim.sbins)
print("surface atoms", int(r["is_surface"].sum()), "rate range", r["rate"][r["is_surface"]>0].min(), r["rate"].max())
