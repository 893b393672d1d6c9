"""CPU-side checks: the C-ABI library loads and exports every symbol include/kmc_b200.h declares,
the host-side constants reproduce the reference's Constants.py bit for bit, the product fails
loudly without a GPU, and the oracle's large-system mode is self-consistent."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle as orc


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from kmcislandgrowth_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libkmcb200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.load().kmc_version() >= 100


def test_struct_layout_matches_header():
    from kmcislandgrowth_b200 import _lib
    assert ctypes.sizeof(_lib.KmcParams) == 11 * 8 + 6 * 4
    assert ctypes.sizeof(_lib.KmcRelaxStats) == 3 * 8 + 2 * 8 + 2 * 4


@pytest.mark.parametrize("name", ["default", "w24misfit", "w20"])
def test_constants_match_reference(name):
    from kmcislandgrowth_b200 import Constants, runtime, workloads
    c = json.load(open(os.path.join(GOLDEN, "static_%s.json" % name)))["constants"]
    gv = workloads.constants_for(c["W"], E_a=c["E_a"], E_s=c["E_s"], r_a=c["r_a"], r_s=c["r_s"])
    for k, v in c.items():
        assert getattr(gv, k) == v, k                    # bit-exact, incl. the fp-fragile nbins_x
    p = runtime.params_from(gv)
    assert p.Rd == 6.0 * 18 / c["W"] * 1.0 and p.omega == 1.0e12
    op = orc.make_params(c["E_a"], c["E_s"], c["r_a"], c["r_s"], c["W"])
    for k in ("L", "bin_size", "Dmin", "Dmax", "E_a", "r_a", "E_as", "r_as", "beta", "omega", "Rd", "nbins_x", "nbins_y"):
        assert getattr(p, k) == getattr(op, k), k
    assert Constants.W == 18 and Constants.nbins_x == 6      # module defaults untouched


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kmcislandgrowth_b200 import Energy, _lib
    with pytest.raises(_lib.KmcError) as e:
        Energy.AdatomAdatomForces(np.array([0.0, 1.0, 2.7, 1.0]))
    assert "no CPU fallback" in str(e.value)


def test_workloads_are_deterministic_and_sane():
    from kmcislandgrowth_b200 import workloads
    gv, sub, ad = workloads.make("cfg2_256x256_10k")
    assert (gv.nbins_x, gv.W, gv.Hs) == (86, 256, 256) and sub.size == 2 * 65536 and ad.size == 20000
    gv2, sub2, ad2 = workloads.make("cfg2_256x256_10k")
    assert np.array_equal(ad, ad2) and np.array_equal(sub, sub2)
    assert np.all(np.abs(ad[0::2]) <= gv.L / 2) and ad[1::2].min() > 1.5
    # substrate fixture equals the oracle's restatement of InitSubstrate
    g = workloads.constants_for(18)
    assert np.array_equal(workloads.substrate(g), orc.init_substrate(18, 10, 2.7, 48.6))
    # top of the film stays inside the cell grid (no y-wrap aliasing)
    assert ad[1::2].max() < gv.Dmin + (gv.nbins_y - 1) * gv.bin_size


def test_oracle_bins3x3_mode_consistency():
    """aa_mode 1 == aa_mode 0 restricted to the stencil; forces are the gradient of the energy."""
    from kmcislandgrowth_b200 import workloads
    gv = workloads.constants_for(48, 12, 10, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 300, seed=9)
    o1 = orc.Oracle(orc.make_params(W=48, Hs=12, Ha=10, aa_mode=1))
    sb = o1.PutInBins(sub)
    F = o1.AdatomAdatomForces(ad) + o1.AdatomSubstrateForces(ad, sb)
    e0 = o1.RelaxEnergy((ad, sb))
    h = 1e-6
    for i in (0, 7, 123, 401):
        d = np.zeros_like(ad)
        d[i] = h
        num = -(o1.RelaxEnergy((ad + d, sb)) - o1.RelaxEnergy((ad - d, sb))) / (2 * h)
        assert abs(num - F[i]) < 1e-5 * max(1.0, abs(F[i]))
    assert np.isfinite(e0)
    # select: dense-with-zeros == compacted list
    rate, surf, du = o1.HoppingRatesDense(ad, sb)
    comp = rate[surf]
    pj = np.concatenate([[0.0], np.cumsum(comp)])
    for u in np.linspace(0, 0.999, 17):
        k, dense, rtot = o1.Select(rate, u)
        r = u * rtot
        hop = np.nonzero(pj > r)[0]
        want = int(hop[0]) - 1 if hop.size else -1
        assert k == want
        if k >= 0:
            assert dense == np.nonzero(surf)[0][k]


def test_island_statistics_against_brute_force():
    """Growth.IslandStatistics: connected components under d < 1.2 r_a with the x-periodic minimum image."""
    from kmcislandgrowth_b200 import Growth, workloads
    gv = workloads.constants_for(48, 12, 10)
    ad = workloads.film(gv, 150, seed=3, sigma=0.05)
    got = Growth.IslandStatistics(ad, gv)
    o = orc.Oracle(orc.make_params(W=48, Hs=12, Ha=10))
    d = o.Distances(ad, ad)
    n = ad.size // 2
    adj = d < 1.2 * gv.r_a
    lab = -np.ones(n, dtype=int)
    comp = 0
    for i in range(n):
        if lab[i] >= 0:
            continue
        stack = [i]
        lab[i] = comp
        while stack:
            k = stack.pop()
            for j in np.nonzero(adj[k])[0]:
                if lab[j] < 0:
                    lab[j] = comp
                    stack.append(j)
        comp += 1
    assert got["n_islands"] == comp
    assert got["sizes"] == sorted(np.bincount(lab).tolist(), reverse=True)
    # the golden trajectory of the reference: island statistics per event are reproducible observables
    z = np.load(os.path.join(GOLDEN, "trajectory.npz"))
    offs, vals = z["states_offs"], z["states_vals"]
    g18 = workloads.constants_for(18)
    last = Growth.IslandStatistics(vals[offs[-2]:offs[-1]], g18)
    assert last["n_islands"] >= 1 and sum(last["sizes"]) == (offs[-1] - offs[-2]) // 2
