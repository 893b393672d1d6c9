"""Streaming kernels on an HBM-sized input: cell-list build (Bins.PutInBins) of the 4096x4096 substrate of BASELINE
config 5 (16.8 M atoms, 268 MB of positions; larger than the 126 MB L2) and PutAllInBox on the same array."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kmcislandgrowth_b200 import runtime, workloads
PEAK = 6509.1
gv = workloads.constants_for(4096, 4096, 10, aa_mode=1)
sub = workloads.substrate(gv)
n = sub.size // 2
ctx = runtime.Context(runtime.params_from(gv), 0)
d = torch.from_numpy(sub).cuda()
ctx.use_torch_stream()
for _ in range(2):
    b = ctx.bins_create(d); b.close()
ctx.timing_reset(); ctx.timing(True)
reps = 5
for _ in range(reps):
    b = ctx.bins_create(d); b.close()
    ctx.put_all_in_box(d)
ctx.timing(False)
out = {"atoms": n, "cells": gv.nbins_x * gv.nbins_y}
tot = 0
for fam in ("bin_count", "bin_scan", "bin_scatter", "bin_order", "misc_wrap"):
    ms, cnt = ctx.timing_read(fam)
    out[fam + "_ms"] = round(ms / max(cnt, 1), 3)
    if fam.startswith("bin_"): tot += ms / max(cnt, 1)
abytes = 44 * n + 8 * out["cells"]
out["bin_build_ms"] = round(tot, 3)
out["bin_build_algorithmic_GBs"] = abytes / (tot * 1e-3) / 1e9
out["bin_build_hbm_frac"] = out["bin_build_algorithmic_GBs"] / PEAK
out["wrap_GBs"] = 32 * n / (out["misc_wrap_ms"] * 1e-3) / 1e9
out["wrap_hbm_frac"] = out["wrap_GBs"] / PEAK
print(json.dumps(out))
