"""stress: repeat the bench's event loop (configs[1]) until something fails; variants isolate what the failure needs.
    python scripts/stress_events.py REPS [notiming] [noflush] [ownstream] [nofast] [hostmode]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
flags = set(sys.argv[2:])
if "nofast" in flags:
    os.environ["KMC_DEBUG_E"] = "9"
for f in flags:
    if f.startswith("dbg"):
        os.environ["KMC_DEBUG_E"] = f[3:]
from kmcislandgrowth_b200 import Growth, runtime, workloads
reps = int(sys.argv[1])
name = os.environ.get("KMC_STRESS_WORKLOAD", "cfg2_256x256_10k")
if os.environ.get("KMC_STRESS_N"):
    W_, Hs_, Ha_, n_, seed_, aa_ = workloads.CONFIGS[name]
    workloads.CONFIGS[name] = (W_, Hs_, Ha_, int(os.environ["KMC_STRESS_N"]), seed_, aa_)
gv, sub, ad0 = workloads.make(name)
sim = Growth.Simulation(adatoms=ad0, substrate=sub, device=0, params=runtime.params_from(gv))
ctx = sim.ctx
if "ownstream" not in flags:
    ctx.use_torch_stream()
if "hostmode" in flags:
    sim.set_modes(event_mode=0)
st, total = workloads.bench_prerelax(ctx, sim.sbins, int(os.environ.get("KMC_STRESS_PRERELAX", "200000")))
print("prerelax", total, int(st.status), flush=True)
state0 = sim.adatoms.copy()
us = workloads.bench_uniforms(64)
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
ref = None
t0 = time.perf_counter()
for r in range(reps):
    ctx.state_set(state0)
    if "notiming" not in flags:
        ctx.timing_reset(); ctx.timing(True)
    sig = []
    try:
        for s in range(23):
            if "noflush" not in flags:
                flush.zero_()
            rec = sim.step(us[s, 0], us[s, 1], int(os.environ.get("KMC_STRESS_CAP", "0")))
            if os.environ.get("KMC_STRESS_VERBOSE"):
                print(s, rec, flush=True)
            sig.append((rec["kind"], rec["hop_k"], rec["line_searches"], rec["restarts"]))
    except Exception as e:
        print("FAILED rep %d step %d after %.1fs: %s" % (r, len(sig), time.perf_counter() - t0, str(e)[:200]), flush=True)
        sys.exit(3)
    if ref is None:
        ref = sig
    elif sig != ref:
        bad = [i for i, (a, b) in enumerate(zip(sig, ref)) if a != b]
        print("MISMATCH rep %d at steps %s: %s vs %s" % (r, bad[:4], sig[bad[0]], ref[bad[0]]), flush=True)
        sys.exit(4)
print("ok %d reps, %.1fs, flags %s" % (reps, time.perf_counter() - t0, sorted(flags)), flush=True)
