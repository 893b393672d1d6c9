"""Small fixed workload for ncu: a few line-search energy sweeps + force sweeps + a rate table on cfg2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_256x256_10k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gv, sub, ad = workloads.make(name)
ctx = runtime.Context(runtime.params_from(gv), 0)
sb = ctx.bins_create(sub)
F = ctx.forces(ad, sb, 3)
for _ in range(reps):
    e = ctx.line_energies(ad, F, sb, 16, 20)
    F = ctx.forces(ad, sb, 3)
r = ctx.rates(ad, sb)
print("energies", e[:3], "launches", ctx.launch_count())
