import os, sys
sys.path.insert(0, "/root/repo")
from kmcislandgrowth_b200 import runtime, workloads, Growth
gv, sub, ad = workloads.make("cfg2_256x256_10k")
sim = Growth.Simulation(adatoms=ad, substrate=sub, params=runtime.params_from(gv))
os.environ.pop("KMC_DEBUG_E", None)
st = sim.ctx.relax(sim.sbins, 16, 1e-4, 3000)
for dbg in (sys.argv[1:] or ["0", "1"]):
    os.environ["KMC_DEBUG_E"] = dbg
    sys.stderr.write("debug=%s\n" % dbg)
    x = sim.adatoms.copy()
    st = sim.ctx.relax(sim.sbins, 16, 1e-4, 300)
    sim.set_adatoms(x)
