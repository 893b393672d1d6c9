"""Diagnostic: per-block phase time sums of k_relax2 (KMC_DEBUG_E=8) for the configs[1] film and for the same film shifted by half
a box in x (periodic).  If the slow blocks stay at the same block ids the imbalance is SM placement, if they move it is data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KMC_DEBUG_E"] = "8"
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads, Growth, Periodic
gv, sub, ad = workloads.make("cfg2_256x256_10k")
for shift in (0.0,):
    x = ad.copy()
    x[0::2] = workloads.wrap_x(x[0::2] + shift * gv.L, gv.L)
    sim = Growth.Simulation(adatoms=x, substrate=sub, params=runtime.params_from(gv))
    sys.stderr.write("---- shift %.1f L: warm-up\n" % shift)
    sim.ctx.relax(sim.sbins, 16, 1e-4, 2000)
    sys.stderr.write("---- shift %.1f L: measured\n" % shift)
    st = sim.ctx.relax(sim.sbins, 16, 1e-4, 400)
    sim.close()
