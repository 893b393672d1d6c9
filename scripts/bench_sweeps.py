"""Throughput of the sweep kernels (forces, 20-candidate line energies, rate table, cell-list build, selection)
on BASELINE configs 2 and 3, fp64 and fp32 pair terms: pair evaluations/s and algorithmic-bytes HBM fraction
(SURVEY.md §8d).  CUDA-event timing via the library's per-family timers.  Prints one JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmcislandgrowth_b200 import runtime, workloads

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
reps = 20
for name in ("cfg2_256x256_10k", "cfg3_1024x1024_500k"):
    gv, sub, ad = workloads.make(name)
    n = ad.size // 2
    for prec in (0, 1):
        gv.precision = prec
        ctx = runtime.Context(runtime.params_from(gv), 0)
        sb = ctx.bins_create(sub)
        ab = ctx.bins_create(ad)
        _, offs_a = ctx.nearby_atoms(ab, ad)
        _, offs_s = ctx.nearby_atoms(sb, ad)
        pairs = int(offs_a[n]) + int(offs_s[n])          # ordered pair visits of one force sweep
        # active substrate atoms / touched cells for the byte formula
        cells_a = np.unique(ctx.bins_export(ab)[0])
        ab.close()
        F = ctx.forces(ad, sb, 3)
        for _ in range(3):
            ctx.forces(ad, sb, 3); ctx.line_energies(ad, F, sb, 16, 20); ctx.rates(ad, sb)
        ctx.timing_reset(); ctx.timing(True)
        for _ in range(reps):
            ctx.forces(ad, sb, 3); ctx.line_energies(ad, F, sb, 16, 20); r = ctx.rates(ad, sb); ctx.select(r["rate"], r["is_surface"], 0.5)
        ctx.timing(False)
        out = {"workload": name, "adatoms": n, "precision": "fp32 pair terms" if prec else "fp64", "ordered_pairs_per_force_sweep": pairs}
        for fam, label, bytes_per in (("sweep_forces", "force_sweep", 40 * n + 16 * int(offs_s[n] > 0) * 0 + 8 * len(cells_a)),
                                      ("line_energy", "line_energy_20_candidates", 32 * n + 8 * len(cells_a) + 160),
                                      ("sweep_rates", "rate_table", 33 * n + 16 * n),
                                      ("bin_", "cell_list_build_all", 44 * n),
                                      ("select", "select", 8 * n)):
            ms, cnt = ctx.timing_read(fam)
            if cnt:
                per = ms / cnt
                out[label] = {"us_per_launch": round(per * 1e3, 2), "launches": cnt}
                if label == "force_sweep":
                    out[label].update(pair_evals_per_s=pairs / (per * 1e-3), algorithmic_GBs=bytes_per / (per * 1e-3) / 1e9,
                                      hbm_frac=bytes_per / (per * 1e-3) / 1e9 / PEAK, fp_gflops=24 * pairs / (per * 1e-3) / 1e9)
                if label == "line_energy_20_candidates":
                    out[label].update(pair_evals_per_s=20 * (int(offs_a[n]) // 2 + int(offs_s[n])) / (per * 1e-3),
                                      algorithmic_GBs=bytes_per / (per * 1e-3) / 1e9, hbm_frac=bytes_per / (per * 1e-3) / 1e9 / PEAK)
                if label in ("rate_table", "select"):
                    out[label].update(algorithmic_GBs=bytes_per / (per * 1e-3) / 1e9, hbm_frac=bytes_per / (per * 1e-3) / 1e9 / PEAK)
        ctx.timing_reset()
        print(json.dumps(out))
        sb.close(); ctx.close()
