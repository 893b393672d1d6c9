"""debug: which event mode / step of tests/test_gpu_parity.py::test_event_modes_agree fails, and the state around it"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import Growth, runtime, workloads
for W, Hs, Ha, n, aa in [(18, 10, 10, 30, 0), (64, 16, 16, 900, 1)]:
    gv = workloads.constants_for(W, Hs, Ha, aa_mode=aa, hop_index_mode=1 if aa else 0, deposit_mode=1 if aa else 0)
    sub = workloads.substrate(gv)
    ad = workloads.film(gv, n, 5)
    us = np.random.default_rng(77).uniform(0, 1, (10, 2))
    us[3, 0] = 0.999999999
    us[7, 0] = 1.0 - 2 ** -53
    for mode in (1, 0):
        sim = Growth.Simulation(adatoms=ad, substrate=sub, device=0, params=runtime.params_from(gv))
        sim.set_modes(event_mode=mode)
        for s in range(10):
            try:
                rec = sim.step(us[s, 0], us[s, 1], 40)
                a = sim.adatoms
                print("W", W, "mode", mode, "step", s, rec["kind"], rec["hop_k"], rec["line_searches"], rec["status"], "n", a.size // 2,
                      "ymin %.3f ymax %.3f nan %d" % (a[1::2].min(), a[1::2].max(), int(np.isnan(a).sum())), "u", us[s])
            except Exception as e:
                a = sim.adatoms
                print("W", W, "mode", mode, "step", s, "FAILED", e, "n", a.size // 2, "ymin %.3f ymax %.3f nan %d" % (a[1::2].min(), a[1::2].max(), int(np.isnan(a).sum())),
                      "new_x", us[s, 1] * gv.L, "L", gv.L, "Dmax", gv.Dmax)
                break
        sim.close()
