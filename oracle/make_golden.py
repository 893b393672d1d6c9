#!/usr/bin/env python3
"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (test infrastructure, never shipped).

The reference (/root/reference, Python 2.7) holds no tests or golden vectors of its own
(SURVEY.md §4), so parity would be unpinned.  This script makes the reference runnable on
Python 3 by a purely mechanical, in-memory conversion written to a
temporary directory OUTSIDE the repo, imports it, evaluates it on seeded inputs and dumps the
input/output pairs as small .npz / .json fixtures under tests/golden/.  No reference source is
copied into the repository; only numbers computed by it are.

Run in the build container (where /root/reference exists):
    python oracle/make_golden.py [--out tests/golden] [--events 14]
The GPU box never runs this; it only reads the committed fixtures.

Mechanical conversion applied (value-identical to Python 2 semantics):
  1. ``len(X)/2``  -> ``len(X)//2``       (all lengths are even; Periodic.py:23,51,56,60 etc.)
  2. ``print X``   -> ``print(X)``        (2DGrowth.py:112,139,161,163,197,323,326)
  3. ``np.array(ragged)`` -> 1-D object array (Bins.py:42,111; NumPy>=1.24 refuses ragged)
  4. 2DGrowth.py: functions only (lines 1-303), matplotlib stubbed, Pool(1) -> in-process map
  5. random.random / random.randint patched to draw from an explicit uniform stream
     (randint(0, n-1) -> int(u*n)); consumption order per event documented in SURVEY §8c.
"""
import argparse
import importlib
import io
import json
import os
import re
import sys
import tempfile
import time
import types

import numpy as np

REF = os.environ.get("KMC_REFERENCE", "/root/reference")


def convert_reference(dst):
    """Write the mechanically converted py3 copy of the reference into `dst` (a tmp dir)."""
    len_div = re.compile(r"len\(([^()]*)\)/2")
    ragged_helper = (
        "\n\ndef _obj_array(seq):\n"
        "\tout = np.empty(len(seq), dtype=object)\n"
        "\tfor _i, _v in enumerate(seq):\n"
        "\t\tout[_i] = _v\n"
        "\treturn out\n"
    )
    for name in ("Constants.py", "Periodic.py", "Bins.py", "Energy.py", "2DGrowth.py"):
        src = io.open(os.path.join(REF, name), encoding="utf-8").read()
        src = len_div.sub(r"len(\1)//2", src)
        if name == "Bins.py":
            src = src.replace("return np.array(bins)", "return _bins_obj(bins)")
            src = src.replace(
                "return (np.array(nearest_adatoms), np.array(nearest_substrate))",
                "return (_obj_array(nearest_adatoms), _obj_array(nearest_substrate))",
            )
            # bins stay plain nested python lists inside a 1-D object array of columns
            src += ragged_helper + (
                "\n\ndef _bins_obj(bins):\n"
                "\tout = np.empty(len(bins), dtype=object)\n"
                "\tfor _i, _c in enumerate(bins):\n"
                "\t\tout[_i] = _c\n"
                "\treturn out\n"
            )
        if name == "2DGrowth.py":
            lines = src.split("\n")[:303]          # functions only; drop the top-level loop
            src = "\n".join(lines)
            src = re.sub(r"^(\s*)print (.*)$", r"\1print(\2)", src, flags=re.M)
            src = src.replace("from matplotlib import pyplot as plt", "plt = None")
            src = src.replace("from multiprocessing import Pool", "")
            src = src.replace(
                "pool = Pool(1)",
                "class _InProcPool(object):\n\tdef map(self, f, xs):\n\t\treturn list(map(f, xs))\npool = _InProcPool()",
            )
            name = "Growth.py"
        with io.open(os.path.join(dst, name), "w", encoding="utf-8") as f:
            f.write(src)


class UniformStream(object):
    """Explicit uniform stream standing in for the `random` module."""

    def __init__(self, us):
        self.us = list(us)
        self.pos = 0

    def random(self):
        u = self.us[self.pos]
        self.pos += 1
        return u

    def randint(self, a, b):
        u = self.random()
        return a + int(u * (b - a + 1))


def load_reference(tmp, argv=None):
    """(Re)import the converted reference with Constants configured through sys.argv."""
    old_argv = sys.argv
    sys.argv = ["2DGrowth.py"] + ([str(a) for a in argv] if argv else [])
    if tmp not in sys.path:
        sys.path.insert(0, tmp)
    try:
        mods = {}
        for n in ("Constants", "Periodic", "Bins", "Energy", "Growth"):
            if n in sys.modules:
                mods[n] = importlib.reload(sys.modules[n])
            else:
                mods[n] = importlib.import_module(n)
    finally:
        sys.argv = old_argv
    return types.SimpleNamespace(**mods)


def constants_dict(gv):
    keys = ["E_a", "E_s", "r_a", "r_s", "W", "E_as", "r_as", "sqrt3", "Hs", "Ha", "L", "Dmin", "Dmax", "D",
            "bin_size", "nbins_x", "nbins_y", "boltzmann", "beta", "deposition_rate"]
    return {k: (int(getattr(gv, k)) if isinstance(getattr(gv, k), (int, np.integer)) else float(getattr(gv, k)))
            for k in keys}


def ragged_to_flat(obj):
    """object array / list of 1-D float arrays -> (flat values, offsets in doubles)."""
    vals, offs = [], [0]
    for a in obj:
        a = np.asarray(a, dtype=np.float64).ravel()
        vals.append(a)
        offs.append(offs[-1] + a.size)
    return (np.concatenate(vals) if vals else np.zeros(0)), np.asarray(offs, dtype=np.int64)


def bins_to_flat(bins):
    """reference bins[nx][ny] (lists of floats) -> flat values + per-column row counts + offsets."""
    ncols = len(bins)
    rows = np.asarray([len(c) for c in bins], dtype=np.int64)
    vals, offs = [], [0]
    for c in bins:
        for cell in c:
            a = np.asarray(cell, dtype=np.float64).ravel()
            vals.append(a)
            offs.append(offs[-1] + a.size)
    return (np.concatenate(vals) if vals else np.zeros(0)), rows, np.asarray(offs, dtype=np.int64), ncols


def make_adatoms(rng, gv, n, layers_sigma=0.02, spread=1.0):
    """Seeded adatom set: perturbed hex layers above the substrate, random subset of sites.

    Distances stay >= ~0.8 r_a so LJ terms remain O(1); x in (-L/2, L/2].
    """
    sites = []
    nl = max(1, int(np.ceil(n / float(gv.W) * 1.6)) + 1)
    for layer in range(nl):
        y = (layer + 1) * gv.r_a * gv.sqrt3 / 2
        for k in range(gv.W):
            x = k * gv.r_a * spread + (gv.r_a / 2 if layer % 2 == 0 else 0.0)
            sites.append((x, y))
    sites = np.asarray(sites)
    # favour low layers: weights decay with height so islands form
    w = np.exp(-sites[:, 1] / (gv.r_a * 1.5))
    w /= w.sum()
    pick = rng.choice(len(sites), size=min(n, len(sites)), replace=False, p=w)
    pts = sites[pick] + rng.normal(0.0, layers_sigma, size=(len(pick), 2))
    R = pts.reshape(-1).copy()
    return np.asarray([v for v in R], dtype=np.float64)


def static_case(ref, rng, name, ns, out, with_rates_upto=100):
    """All per-call hot-path functions on seeded inputs for the currently loaded Constants."""
    gv, P, B, E, G = ref.Constants, ref.Periodic, ref.Bins, ref.Energy, ref.Growth
    data = {}
    meta = {"constants": constants_dict(gv), "sets": []}

    # --- Periodic (Periodic.py:8-62) ---
    xs = np.concatenate([
        rng.uniform(-3 * gv.L, 3 * gv.L, 40),
        np.array([0.0, gv.L / 2, -gv.L / 2, gv.L, -gv.L, gv.L * 1.5, -gv.L * 1.5, gv.L / 2 + 1e-9,
                  -gv.L / 2 - 1e-9, 24.3, 25.0, -30.0, 1e-300, -1e-300]),
    ])
    data["putinbox_in"] = xs
    data["putinbox_out"] = np.asarray([P.PutInBox(float(x)) for x in xs])
    a = rng.uniform(-gv.L, gv.L, 2 * 7)
    a[1::2] = rng.uniform(gv.Dmin, gv.Dmax, 7)
    b = rng.uniform(-gv.L, gv.L, 2 * 9)
    b[1::2] = rng.uniform(gv.Dmin, gv.Dmax, 9)
    data["dist_a"], data["dist_b"] = a.copy(), b.copy()
    data["dist_out"] = np.asarray(P.Distances(a.copy(), b.copy()))
    data["disp_out"] = np.asarray(P.Displacement(a.copy(), a[::-1].copy()))
    data["displs_out"] = np.asarray(P.Displacements(float(a[0]), float(a[1]), b.copy()))
    data["distance_out"] = np.asarray([P.Distance(a[0:2].copy(), b[2 * j:2 * j + 2].copy()) for j in range(9)])

    # --- Bins (Bins.py:9-70) ---
    px = np.concatenate([rng.uniform(-gv.L / 2 - 10, gv.L / 2 + 10, 60),
                         np.array([0.0, gv.L / 2, -gv.L / 2, 24.3, -24.29, -25.0, 30.0, gv.L / 2 - 1e-12])])
    py = np.concatenate([rng.uniform(gv.Dmin - 3, gv.Dmax + 12, 60),
                         np.array([0.0, 2.34, gv.Dmin, gv.Dmax, -21.04, -24.0, 30.0, 2.7])])
    data["bin_px"], data["bin_py"] = px, py
    data["binindex_out"] = np.asarray([B.BinIndex(float(x), float(y)) for x, y in zip(px, py)], dtype=np.int64)
    nbx, nby, nbo = [], [], [0]
    for x, y in zip(px, py):
        (cx, cy) = B.NearbyBins(float(x), float(y))
        nbx += list(cx)
        nby += list(cy)
        nbo.append(len(nbx))
    data["nearbybins_x"] = np.asarray(nbx, dtype=np.int64)
    data["nearbybins_y"] = np.asarray(nby, dtype=np.int64)   # always 3 per point
    data["nearbybins_xoff"] = np.asarray(nbo, dtype=np.int64)

    # --- substrate (2DGrowth.py:18-37) + its bins (Bins.py:24-42) ---
    substrate = G.InitSubstrate()
    sbins = B.PutInBins(substrate)
    data["substrate"] = substrate
    v, rows, offs, ncols = bins_to_flat(sbins)
    data["sbins_vals"], data["sbins_rows"], data["sbins_offs"] = v, rows, offs
    # NearbyAtoms on substrate bins at probe points (only points with valid python indexing)
    qx = np.concatenate([rng.uniform(-gv.L / 2, gv.L / 2, 24), np.array([-gv.L / 2 + 0.1, gv.L / 2 - 0.1, 0.0])])
    qy = np.concatenate([rng.uniform(0.5, 3 * gv.r_a, 24), np.array([2.7, 2.7, 2.7])])
    na_vals, na_offs = ragged_to_flat([B.NearbyAtoms(float(x), float(y), sbins) for x, y in zip(qx, qy)])
    data["nearby_qx"], data["nearby_qy"], data["nearby_vals"], data["nearby_offs"] = qx, qy, na_vals, na_offs
    data["surf_energy_q"] = np.asarray([E.AdatomSurfaceEnergy(np.array([x, y]), sbins) for x, y in zip(qx, qy)])
    data["as_forces_q"] = np.asarray(E.AdatomSubstrateForces(np.column_stack([qx, qy]).reshape(-1), sbins))

    # --- seeded adatom sets (Energy.py, Bins.py:86-111, 2DGrowth.py:46-62,201-227) ---
    for n in ns:
        ad = make_adatoms(rng, gv, n)
        n_eff = len(ad) // 2
        key = "%s_n%d" % (name, n_eff)
        t0 = time.time()
        rec = {"n": n_eff, "key": key}
        data[key + "_ad"] = ad.copy()
        data[key + "_faa"] = np.asarray(E.AdatomAdatomForces(ad.copy()))
        data[key + "_fas"] = np.asarray(E.AdatomSubstrateForces(ad.copy(), sbins))
        data[key + "_relax_energy"] = np.float64(E.RelaxEnergy((ad.copy(), sbins)))
        data[key + "_total_energy"] = np.float64(E.TotalEnergy(ad.copy(), sbins)) if n_eff >= 1 else np.float64(0)
        # 20-candidate line-search energies exactly as 2DGrowth.py:120-122 builds them
        dx0 = data[key + "_faa"] + data[key + "_fas"]
        scale = 16
        xs20 = [(ad + a_ * dx0 / 20 / scale, sbins) for a_ in range(20)]
        data[key + "_ls_energies"] = np.asarray([E.RelaxEnergy(t) for t in xs20])
        # probes (HoppingEnergy, Energy.py:170-172)
        pr = np.column_stack([rng.uniform(-gv.L / 2, gv.L / 2, 6), rng.uniform(2.0, 4 * gv.r_a, 6)])
        data[key + "_probes"] = pr.reshape(-1)
        data[key + "_hop_energy"] = np.asarray([E.HoppingEnergy(p.copy(), ad.copy(), sbins) for p in pr])
        (nn_a, nn_s) = B.NearestNeighbors(ad.copy(), sbins)
        data[key + "_nn_a_vals"], data[key + "_nn_a_offs"] = ragged_to_flat(nn_a)
        data[key + "_nn_s_vals"], data[key + "_nn_s_offs"] = ragged_to_flat(nn_s)
        surf = G.SurfaceAtoms(ad.copy(), sbins)
        data[key + "_surface_pos"] = np.asarray(surf)
        if n_eff <= with_rates_upto:
            if len(surf) > 0:
                sidx = [list(Ds).index(0) for Ds in P.Distances(surf, ad.copy())]
            else:
                sidx = []
            data[key + "_surface_idx"] = np.asarray(sidx, dtype=np.int64)
            data[key + "_uappx"] = np.asarray([E.U_appx(i, ad.copy(), sbins) for i in sidx])
            data[key + "_uloc"] = np.asarray([E.U_loc(i, ad.copy(), sbins) for i in sidx])
            data[key + "_uideal"] = np.asarray([E.U_loc_ideal(i, ad.copy(), sbins) for i in sidx], dtype=np.float64)
            data[key + "_deltaU"] = np.asarray([E.DeltaU(i, ad.copy(), sbins) for i in sidx])
            Rk = G.HoppingRates(ad.copy(), sbins)
            data[key + "_rates"] = np.asarray(Rk, dtype=np.float64)
            pj = G.HoppingPartialSums(Rk)
            data[key + "_pj"] = np.asarray(pj, dtype=np.float64)
            Rd = G.DepositionRate()
            Rtot = G.TotalRate(Rd, Rk)
            data[key + "_Rd"], data[key + "_Rtot"] = np.float64(Rd), np.float64(Rtot)
            us = np.concatenate([rng.uniform(0, 1, 200), np.array([0.0, 1.0 - 2 ** -53])])
            sel = []
            for u in us:
                r = u * Rtot
                hop = [p for p in pj if p > r]
                sel.append(pj.index(hop[0]) - 1 if len(hop) > 0 else -1)   # -1 = deposit (2DGrowth.py:324-329)
            data[key + "_sel_u"], data[key + "_sel_idx"] = us, np.asarray(sel, dtype=np.int64)
            rec["rates"] = True
        rec["seconds"] = round(time.time() - t0, 2)
        meta["sets"].append(rec)
        print("  %s: n=%d done in %.1fs" % (key, n_eff, rec["seconds"]), flush=True)
    np.savez_compressed(os.path.join(out, "static_%s.npz" % name), **data)
    with open(os.path.join(out, "static_%s.json" % name), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def relaxation_trace(ref, out, seed=7):
    """One Relaxation call (2DGrowth.py:97-166) traced line search by line search."""
    gv, B, E, G = ref.Constants, ref.Bins, ref.Energy, ref.Growth
    rng = np.random.default_rng(seed)
    sbins = B.PutInBins(G.InitSubstrate())
    ad = make_adatoms(rng, gv, 9, layers_sigma=0.05)
    trace = {"energies": [], "forces_max": []}
    real_relax = E.RelaxEnergy
    cur = []

    def spy(args):
        v = real_relax(args)
        cur.append(v)
        if len(cur) == 20:
            trace["energies"].append(list(cur))
            del cur[:]
        return v

    E.RelaxEnergy = spy
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        res = G.Relaxation(ad.copy(), sbins)
    finally:
        sys.stdout = old
        E.RelaxEnergy = real_relax
    np.savez_compressed(os.path.join(out, "relax_trace.npz"), ad0=ad, result=np.asarray(res),
                        energies=np.asarray(trace["energies"]), stdout=np.asarray(buf.getvalue()))
    print("  relaxation trace: %d line searches" % len(trace["energies"]), flush=True)


def trajectory(ref, out, n_events, seed=1):
    """Seeded run of the KMC loop (2DGrowth.py:317-329) with an explicit uniform stream."""
    gv, B, E, G = ref.Constants, ref.Bins, ref.Energy, ref.Growth
    import random as _random
    rng = np.random.default_rng(seed)
    us = rng.uniform(0, 1, 4 * n_events + 8)
    stream = UniformStream(us)
    _random.random = stream.random
    _random.randint = stream.randint
    substrate = G.InitSubstrate()
    sbins = B.PutInBins(substrate)
    adatoms = np.array([])
    states, kinds, idxs, rtots, rs, nls, used = [], [], [], [], [], [], []
    ls_count = [0]
    real_relax = E.RelaxEnergy

    def spy(args):
        ls_count[0] += 1
        return real_relax(args)

    E.RelaxEnergy = spy
    old = sys.stdout
    t0 = time.time()
    try:
        for t in range(n_events):
            sys.stdout = io.StringIO()
            Rk = G.HoppingRates(adatoms, sbins)
            pj = G.HoppingPartialSums(Rk)
            Rd = G.DepositionRate()
            Rtot = G.TotalRate(Rd, Rk)
            r = _random.random() * Rtot
            hop = [p for p in pj if p > r]
            ls_count[0] = 0
            if len(hop) > 0:
                k = pj.index(hop[0]) - 1
                adatoms = G.HopAtom(k, adatoms, sbins)
                kinds.append(1)
                idxs.append(k)
            else:
                adatoms = G.DepositAdatom(adatoms, sbins)
                kinds.append(0)
                idxs.append(-1)
            sys.stdout = old
            rtots.append(Rtot)
            rs.append(r)
            nls.append(ls_count[0] // 20)
            used.append(stream.pos)
            states.append(np.asarray(adatoms, dtype=np.float64).copy())
            print("  event %d kind=%d n=%d line_searches=%d (%.0fs)" % (t, kinds[-1], len(adatoms) // 2, nls[-1], time.time() - t0), flush=True)
    finally:
        sys.stdout = old
        E.RelaxEnergy = real_relax
    vals, offs = ragged_to_flat(states)
    np.savez_compressed(os.path.join(out, "trajectory.npz"), uniforms=us, states_vals=vals, states_offs=offs,
                        kinds=np.asarray(kinds), hop_idx=np.asarray(idxs), rtot=np.asarray(rtots), r=np.asarray(rs),
                        line_searches=np.asarray(nls), uniforms_used=np.asarray(used), seconds=np.float64(time.time() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
    ap.add_argument("--events", type=int, default=14)
    ap.add_argument("--skip-trajectory", action="store_true")
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    os.makedirs(out, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="kmc_ref_py3_")
    convert_reference(tmp)
    print("converted reference in", tmp)

    # case 1: default Constants.py (W=18)
    ref = load_reference(tmp)
    print("default constants", constants_dict(ref.Constants))
    static_case(ref, np.random.default_rng(101), "default", [1, 4, 25, 60, 100, 200], out)
    relaxation_trace(ref, out)
    # case 2: W=24 (L/bin_size exactly 8.0 -> x==L/2 lands in column nbins_x), misfit E_a!=E_s, r_a!=r_s
    ref = load_reference(tmp, [1.9, 2.3, 2.55, 2.7, 24, "_x/"])
    static_case(ref, np.random.default_rng(202), "w24misfit", [4, 30, 80], out)
    # case 3: W=20 (L/bin_size = 6.67 -> partial last column, the case the seam hack was written for)
    ref = load_reference(tmp, [2.083, 2.083, 2.7, 2.7, 20, "_x/"])
    static_case(ref, np.random.default_rng(303), "w20", [4, 30, 80], out)
    if not args.skip_trajectory:
        ref = load_reference(tmp)
        trajectory(ref, out, args.events)
    print("golden vectors written to", out)


if __name__ == "__main__":
    main()
