"""Pins the CPU oracle (oracle/kmc_oracle.c) against outputs of the REFERENCE ITSELF.

tests/golden/*.npz were produced by oracle/make_golden.py, which runs the reference's own
Periodic.py / Bins.py / Energy.py / 2DGrowth.py (mechanically converted to Python 3) on seeded
inputs.  The reference ships no tests or golden vectors, so these are the pins.

Tolerances: integers / bin ids / surface sets / selected event index exact; wrapped positions
bit-exact; energies, forces, rates within 1e-12 relative (summation order and pow() only).
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc

CASES = ["default", "w24misfit", "w20"]


def load(name):
    z = np.load(os.path.join(GOLDEN, "static_%s.npz" % name))
    meta = json.load(open(os.path.join(GOLDEN, "static_%s.json" % name)))
    c = meta["constants"]
    o = orc.Oracle(E_a=c["E_a"], E_s=c["E_s"], r_a=c["r_a"], r_s=c["r_s"], W=c["W"])
    return z, meta, c, o


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor > 0 else 1e-300)))


@pytest.mark.parametrize("name", CASES)
def test_constants(name):
    z, meta, c, o = load(name)
    p = o.p
    for k in ("L", "bin_size", "Dmin", "Dmax", "E_a", "r_a", "E_as", "r_as", "beta"):
        assert getattr(p, k) == c[k], k                     # bit-exact derived constants
    assert (p.nbins_x, p.nbins_y) == (c["nbins_x"], c["nbins_y"])


def test_survey_anchor_values():
    """SURVEY.md §4 anchor table (computed there with the converted reference)."""
    o = orc.Oracle()
    assert (o.p.L, o.p.bin_size, o.p.nbins_x, o.p.nbins_y) == (48.6, 8.100000000000001, 6, 7)
    assert o.p.Dmin == -23.38268590217984 and o.p.beta == 9.670805771536886
    assert o.PutInBox(24.3) == 24.3 and abs(o.PutInBox(25.0) + 23.6) < 1e-12 and abs(o.PutInBox(-30) - 18.6) < 1e-12
    assert o.BinIndex(0, 0) == (2, 2) and o.BinIndex(24.3, 2.34) == (5, 3) and o.BinIndex(-25, -24) == (0, 0)
    assert o.BinIndex(30, 30) == (6, 6)
    assert [o.NearbyBins(-24.3 + 8.1 * k + 1, 2.0)[0] for k in range(6)] == \
        [[5, 0, 1, 4], [0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5, 0], [4, 5, 0]]
    assert o.NearbyBins(0.0, -22.0)[1] == [6, 0, 1]
    sub = orc.init_substrate(18, 10, 2.7, 48.6)
    sb = o.PutInBins(sub)
    assert abs(o.AdatomSurfaceEnergy([0, 2.7], sb) - (-3.3828171662046143)) < 1e-13
    f = o.AdatomSubstrateForces([0, 2.7], sb)
    assert np.allclose(f, [-0.00473487071730617, -1.2791320385082703], rtol=1e-12, atol=1e-14)
    ad = np.array([0, 2.338, 2.7, 2.338, 1.35, 4.676, 10, 2.4])
    assert abs(o.RelaxEnergy((ad, sb)) - (-10.48381743784729)) < 1e-12
    ua, ul, ui, du, na, ns = o.Barriers(ad, sb)
    assert list(zip(na, ns)) == [(3, 1), (3, 1), (3, 0), (1, 2)]
    assert np.allclose(du, [1.8236014173734407, 1.8078460607394078, -2.4479823304411434, -0.6988651525566687], rtol=1e-12)
    rate, surf, _ = o.HoppingRatesDense(ad, sb)
    assert surf.all()
    assert np.allclose(rate, [4.561262554361068e19, 3.9166334353343095e19, 52.30329997297852, 1160865825.637719], rtol=1e-11)


@pytest.mark.parametrize("name", CASES)
def test_periodic(name):
    z, meta, c, o = load(name)
    got = np.array([o.PutInBox(x) for x in z["putinbox_in"]])
    assert np.array_equal(got, z["putinbox_out"])            # bit-exact wrap (Periodic.py:8-20)
    assert np.array_equal(o.PutAllInBox(np.column_stack([z["putinbox_in"], z["putinbox_in"]]).reshape(-1))[0::2],
                          z["putinbox_out"])
    assert relerr(o.Distances(z["dist_a"], z["dist_b"]), z["dist_out"]) < 1e-15
    a = z["dist_a"]
    assert np.array_equal(o.Displacement(a, a[::-1].copy()), z["disp_out"])


@pytest.mark.parametrize("name", CASES)
def test_bins(name):
    z, meta, c, o = load(name)
    got = np.array([o.BinIndex(x, y) for x, y in zip(z["bin_px"], z["bin_py"])])
    assert np.array_equal(got, z["binindex_out"])            # bit-exact (Bins.py:9-22)
    xo = z["nearbybins_xoff"]
    for i, (x, y) in enumerate(zip(z["bin_px"], z["bin_py"])):
        xs, ys = o.NearbyBins(x, y)
        assert xs == list(z["nearbybins_x"][xo[i]:xo[i + 1]]), (x, y)
        assert ys == list(z["nearbybins_y"][3 * i:3 * i + 3]), (x, y)
    # substrate + its cell list
    sub = orc.init_substrate(c["W"], c["Hs"], c["r_s"], c["L"])
    assert np.array_equal(sub, z["substrate"])
    sb = o.PutInBins(sub)
    rows, offs, vals = z["sbins_rows"], z["sbins_offs"], z["sbins_vals"]
    counts = sb.cell_counts().reshape(o.p.nbins_x, o.p.nbins_y)
    k = 0
    outside = 0      # cells the reference grew beyond the addressable grid (x == L/2 when L/bin_size is integral)
    for cx in range(len(rows)):
        for cy in range(rows[cx]):
            cnt = (offs[k + 1] - offs[k]) // 2
            if cx < o.p.nbins_x and cy < o.p.nbins_y:
                assert counts[cx, cy] == cnt
            else:
                outside += cnt
            k += 1
    assert counts.sum() + outside == (offs[-1] // 2)
    assert (sb.cell_of_atom() < 0).sum() == outside
    # NearbyAtoms: same atoms in the same order (Bins.py:72-84)
    no = z["nearby_offs"]
    for i, (x, y) in enumerate(zip(z["nearby_qx"], z["nearby_qy"])):
        assert np.array_equal(o.NearbyAtoms(x, y, sb), z["nearby_vals"][no[i]:no[i + 1]])
    e = np.array([o.AdatomSurfaceEnergy([x, y], sb) for x, y in zip(z["nearby_qx"], z["nearby_qy"])])
    assert relerr(e, z["surf_energy_q"]) < 1e-12
    q = np.column_stack([z["nearby_qx"], z["nearby_qy"]]).reshape(-1)
    assert np.allclose(o.AdatomSubstrateForces(q, sb), z["as_forces_q"], rtol=1e-11, atol=1e-13)


def _sets():
    out = []
    for name in CASES:
        meta = json.load(open(os.path.join(GOLDEN, "static_%s.json" % name)))
        for s in meta["sets"]:
            out.append((name, s["key"], bool(s.get("rates"))))
    return out


@pytest.mark.parametrize("name,key,has_rates", _sets())
def test_adatom_sets(name, key, has_rates):
    z, meta, c, o = load(name)
    sb = o.PutInBins(z["substrate"])
    ad = z[key + "_ad"]
    n = ad.size // 2
    fscale = max(1.0, float(np.max(np.abs(z[key + "_faa"]))), float(np.max(np.abs(z[key + "_fas"]))))
    assert np.max(np.abs(o.AdatomAdatomForces(ad) - z[key + "_faa"])) < 1e-11 * fscale
    assert np.max(np.abs(o.AdatomSubstrateForces(ad, sb) - z[key + "_fas"])) < 1e-11 * fscale
    assert abs(o.RelaxEnergy((ad, sb)) - z[key + "_relax_energy"]) < 1e-12 * abs(z[key + "_relax_energy"])
    if n >= 1:
        assert abs(o.TotalEnergy(ad, sb) - z[key + "_total_energy"]) <= 1e-12 * max(1.0, abs(z[key + "_total_energy"]))
    dx0 = z[key + "_faa"] + z[key + "_fas"]
    ls = np.array([o.RelaxEnergy((ad + a * dx0 / 20 / 16, sb)) for a in range(20)])
    assert relerr(ls, z[key + "_ls_energies"]) < 1e-12
    assert int(np.argmin(ls)) == int(np.argmin(z[key + "_ls_energies"]))
    pr = z[key + "_probes"].reshape(-1, 2)
    he = np.array([o.HoppingEnergy(p, ad, sb) for p in pr])
    assert relerr(he, z[key + "_hop_energy"]) < 1e-12
    na, ns, (av, ao), (sv, so) = o.NearestNeighbors(ad, sb)
    assert np.array_equal(ao, z[key + "_nn_a_offs"]) and np.array_equal(av, z[key + "_nn_a_vals"])
    assert np.array_equal(so, z[key + "_nn_s_offs"]) and np.array_equal(sv, z[key + "_nn_s_vals"])
    rate, surf, du = o.HoppingRatesDense(ad, sb)
    assert np.array_equal(ad.reshape(-1, 2)[surf].reshape(-1), z[key + "_surface_pos"])
    if has_rates:
        sidx = z[key + "_surface_idx"]
        assert np.array_equal(np.nonzero(surf)[0], sidx)
        ua, ul, ui, dU, _, _ = o.Barriers(ad, sb)
        assert relerr(ua[sidx], z[key + "_uappx"]) < 1e-12
        assert relerr(ul[sidx], z[key + "_uloc"], 1e-3) < 1e-11
        assert np.allclose(ui[sidx], z[key + "_uideal"], rtol=1e-15)
        assert np.max(np.abs(dU[sidx] - z[key + "_deltaU"])) < 1e-11
        assert relerr(rate[sidx], z[key + "_rates"]) < 1e-9     # exp amplifies |dDeltaU|*beta
        # selection is checked on the REFERENCE's own rates so that the index must be exact
        dense = np.zeros(n)
        dense[sidx] = z[key + "_rates"]
        for u, want in zip(z[key + "_sel_u"], z[key + "_sel_idx"]):
            k, didx, rtot = o.Select(dense, u)
            assert k == want
            assert rtot == z[key + "_Rtot"]
            if k >= 0:
                assert didx == sidx[k]


def test_relaxation_trace():
    """One full Relaxation call (2DGrowth.py:97-166): same number of line searches, same result."""
    z = np.load(os.path.join(GOLDEN, "relax_trace.npz"))
    o = orc.Oracle()
    sb = o.PutInBins(orc.init_substrate(18, 10, 2.7, 48.6))
    x, st, rc = o.Relaxation(z["ad0"], sb)
    assert rc == 0
    assert st.line_searches == z["energies"].shape[0]
    assert np.max(np.abs(x - z["result"])) < 1e-7


def test_trajectory():
    """Seeded KMC loop (2DGrowth.py:317-329) with the reference's explicit uniform stream."""
    path = os.path.join(GOLDEN, "trajectory.npz")
    z = np.load(path)
    o = orc.Oracle()
    sb = o.PutInBins(orc.init_substrate(18, 10, 2.7, 48.6))
    us, offs, vals = z["uniforms"], z["states_offs"], z["states_vals"]
    ad = np.zeros(0)
    pos = 0
    for t in range(len(z["kinds"])):
        # restart from the reference's state each event so one divergence cannot cascade
        ad_ref_prev = vals[offs[t - 1]:offs[t]] if t > 0 else np.zeros(0)
        ad2, kind, hk, rtot, st, rc = o.Event(ad_ref_prev, sb, us[pos:pos + 2])
        pos = int(z["uniforms_used"][t])
        assert rc == 0
        assert kind == z["kinds"][t], t
        assert hk == z["hop_idx"][t], t
        assert abs(rtot - z["rtot"][t]) <= 1e-9 * abs(z["rtot"][t])
        want = vals[offs[t]:offs[t + 1]]
        assert ad2.size == want.size
        assert st.line_searches == z["line_searches"][t], t
        assert np.max(np.abs(ad2 - want)) < 1e-6, t
