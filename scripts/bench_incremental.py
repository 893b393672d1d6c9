"""Incremental rate table on the bench trajectory (config 2): atoms recomputed per event and time of the update against
the full sweep, for several displacement thresholds.  The trajectory itself runs with full sweeps (parity mode)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads, Growth
gv, sub, ad = workloads.make("cfg2_256x256_10k")
p = runtime.params_from(gv)
sim = Growth.Simulation(adatoms=ad, substrate=sub, params=p)
sim.ctx.relax(sim.sbins, 16, 1e-4, 4000)
thresholds = [0.0, 1e-4, 1e-3, 1e-2]
shadows = []
for thr in thresholds:
    c = runtime.Context(p, 0)
    sb = c.bins_create(sub)
    c.state_set(sim.adatoms)
    c.rates_update(sb, thr)
    shadows.append((c, sb))
rng = np.random.default_rng(1234)
us = rng.uniform(0, 1, (46, 2))
nev = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for t in range(nev):
    before = sim.adatoms.copy()
    r = sim.step(us[t, 0], us[t, 1], 0)
    x = sim.adatoms
    full = sim.ctx.rates(x, sim.sbins)
    t0 = time.perf_counter(); sim.ctx.rates(x, sim.sbins); sim.ctx.synchronize(); t_full = time.perf_counter() - t0
    line = "event %2d kind %d ls %5d | full sweep %.0f us |" % (t, r["kind"], r["line_searches"], t_full * 1e6)
    for thr, (c, sb) in zip(thresholds, shadows):
        # replay the placement on the shadow (same state before the event -> same removal / append), then hand it the
        # relaxed positions of the main trajectory
        if r["kind"] == 1:
            c.hop_place(sb, int(r["hop_k"]), us[t, 1])
        else:
            c.deposit_place(sb, us[t, 1])
        c.state_move(x)
        t0 = time.perf_counter(); inc = c.rates_update(sb, thr); t_inc = time.perf_counter() - t0
        m = full["is_surface"].astype(bool)
        err = float(np.max(np.abs(inc["rate"][m] - full["rate"][m]) / full["rate"][m])) if m.any() else 0.0
        flips = int(np.sum(inc["is_surface"].astype(bool) != m))
        line += " thr %g: redone %5d (%.0f us) max rel err %.1e flips %d |" % (thr, inc["recomputed"], t_inc * 1e6, err, flips)
    print(line, flush=True)
