import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kmcislandgrowth_b200 import Growth, runtime, workloads
gv, sub, ad0 = workloads.make("cfg2_256x256_10k")
sim = Growth.Simulation(adatoms=ad0, substrate=sub, device=0, params=runtime.params_from(gv))
ctx = sim.ctx; ctx.use_torch_stream()
rng = np.random.default_rng(1234)
ctx.relax(sim.sbins, 16, 1e-4, 4000)
us = rng.uniform(0, 1, (46, 2))
for s in range(3): sim.step(us[s,0], us[s,1], 0)
flush = torch.empty(256*2**20, dtype=torch.uint8, device="cuda")
state0 = sim.adatoms.copy()
for trial in range(6):
    sim.set_adatoms(state0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tf = ts = 0.0
    t0 = time.perf_counter(); e0.record()
    for s in range(20):
        a = time.perf_counter(); flush.zero_(); b = time.perf_counter()
        sim.step(us[3+s,0], us[3+s,1], 0); c = time.perf_counter()
        tf += b - a; ts += c - b
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("trial %d wall %.1f ms  events %.1f ms  flush-launch %.2f ms  steps %.1f ms" % (trial, (t1-t0)*1e3, e0.elapsed_time(e1), tf*1e3, ts*1e3))
