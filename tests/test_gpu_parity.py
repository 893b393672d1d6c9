"""GPU parity: the CUDA path (through the C-ABI / drop-in modules) against the golden vectors the
reference produced (tests/golden/) and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): energies, forces, rates within 1e-6 relative in fp64 (we
assert 1e-9 .. 1e-10, the measured agreement is ~1e-14); bin ids, stencils, bond counts, surface
flags and the selected event index bit-exact.  Forces are compared against the force scale
max|F| (net forces are differences of eV/A-sized pair terms).
"""
import json
import os
import types

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CASES = ["default", "w24misfit", "w20"]


@pytest.fixture(autouse=True)
def _fresh_constants():
    from kmcislandgrowth_b200 import Constants
    yield
    Constants.configure()


def setup_case(name):
    from kmcislandgrowth_b200 import Constants
    z = np.load(os.path.join(GOLDEN, "static_%s.npz" % name))
    meta = json.load(open(os.path.join(GOLDEN, "static_%s.json" % name)))
    c = meta["constants"]
    Constants.configure(c["E_a"], c["E_s"], c["r_a"], c["r_s"], c["W"])
    return z, meta, c


def relerr(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.mark.parametrize("name", CASES)
def test_periodic_golden(name):
    from kmcislandgrowth_b200 import Periodic
    z, meta, c = setup_case(name)
    got = np.array([Periodic.PutInBox(x) for x in z["putinbox_in"][:12]])
    assert np.array_equal(got, z["putinbox_out"][:12])
    R = np.column_stack([z["putinbox_in"], z["putinbox_in"]]).reshape(-1).copy()
    ret = Periodic.PutAllInBox(R)
    assert ret is R                                           # in place, returns its argument
    assert np.array_equal(R[0::2], z["putinbox_out"])         # bit-exact wrap
    assert np.array_equal(R[1::2], z["putinbox_in"])          # y untouched
    assert relerr(Periodic.Distances(z["dist_a"], z["dist_b"]), z["dist_out"]) < 1e-15
    a = z["dist_a"]
    assert np.array_equal(Periodic.Displacement(a, a[::-1].copy()), z["disp_out"])
    assert np.array_equal(Periodic.Displacements(float(a[0]), float(a[1]), z["dist_b"]), z["displs_out"])
    assert relerr([Periodic.Distance(a[0:2], z["dist_b"][2 * j:2 * j + 2]) for j in range(9)], z["distance_out"]) < 1e-15


@pytest.mark.parametrize("name", CASES)
def test_bins_golden(name):
    from kmcislandgrowth_b200 import Bins, Energy, Growth, runtime
    z, meta, c = setup_case(name)
    pts = np.column_stack([z["bin_px"], z["bin_py"]]).reshape(-1)
    got = runtime.context().bin_index(pts)
    assert np.array_equal(got, z["binindex_out"])             # bit-exact BinIndex
    xo = z["nearbybins_xoff"]
    for i in range(0, len(z["bin_px"]), 7):
        xs, ys = Bins.NearbyBins(z["bin_px"][i], z["bin_py"][i])
        assert xs == list(z["nearbybins_x"][xo[i]:xo[i + 1]])
        assert ys == list(z["nearbybins_y"][3 * i:3 * i + 3])
    sub = Growth.InitSubstrate()
    assert np.array_equal(sub, z["substrate"])
    sb = Bins.PutInBins(sub)
    # cell contents identical to the reference's ragged lists, in the same order
    rows, offs, vals = z["sbins_rows"], z["sbins_offs"], z["sbins_vals"]
    k = 0
    for cx in range(len(rows)):
        col = sb[cx] if cx < c["nbins_x"] else None
        for cy in range(rows[cx]):
            want = vals[offs[k]:offs[k + 1]]
            if col is not None and cy < c["nbins_y"]:
                have = np.asarray(col[cy]) if cy < len(col) else np.zeros(0)
                assert np.array_equal(have, want), (cx, cy)
            k += 1
    # NearbyAtoms: same atoms, same order; all queries in one launch as well
    no = z["nearby_offs"]
    q = np.column_stack([z["nearby_qx"], z["nearby_qy"]]).reshape(-1)
    out, o2 = runtime.context().nearby_atoms(sb.handle(), q)
    assert np.array_equal(o2 * 2, no) and np.array_equal(out, z["nearby_vals"])
    assert np.array_equal(Bins.NearbyAtoms(q[0], q[1], sb), z["nearby_vals"][no[0]:no[1]])
    e = np.array([Energy.AdatomSurfaceEnergy(q[2 * i:2 * i + 2], sb) for i in range(0, len(q) // 2, 5)])
    assert relerr(e, z["surf_energy_q"][::5]) < 1e-12
    assert np.allclose(Energy.AdatomSubstrateForces(q, sb), z["as_forces_q"], rtol=1e-10, atol=1e-13)


def _sets():
    out = []
    for name in CASES:
        meta = json.load(open(os.path.join(GOLDEN, "static_%s.json" % name)))
        for s in meta["sets"]:
            out.append((name, s["key"], bool(s.get("rates"))))
    return out


@pytest.mark.parametrize("name,key,has_rates", _sets())
def test_energy_golden(name, key, has_rates):
    from kmcislandgrowth_b200 import Bins, Energy, Growth, runtime
    z, meta, c = setup_case(name)
    sb = Bins.PutInBins(z["substrate"])
    ad = z[key + "_ad"]
    n = ad.size // 2
    fscale = max(1.0, float(np.max(np.abs(z[key + "_faa"]))), float(np.max(np.abs(z[key + "_fas"]))))
    assert np.max(np.abs(Energy.AdatomAdatomForces(ad) - z[key + "_faa"])) < 1e-10 * fscale
    assert np.max(np.abs(Energy.AdatomSubstrateForces(ad, sb) - z[key + "_fas"])) < 1e-10 * fscale
    assert np.max(np.abs(Energy.TotalForces(ad, sb) - (z[key + "_faa"] + z[key + "_fas"]))) < 1e-10 * fscale
    assert abs(Energy.RelaxEnergy((ad, sb)) - z[key + "_relax_energy"]) < 1e-11 * abs(z[key + "_relax_energy"])
    assert abs(Energy.TotalEnergy(ad, sb) - z[key + "_total_energy"]) <= 1e-11 * max(1.0, abs(z[key + "_total_energy"]))
    # the 20 line-search candidates: host-formed batch (pool.map replacement) and device-formed
    dx0 = z[key + "_faa"] + z[key + "_fas"]
    xs = [(ad + a * dx0 / 20 / 16, sb) for a in range(20)]
    ls = np.array(Energy.RelaxEnergyBatch(xs))
    assert relerr(ls, z[key + "_ls_energies"]) < 1e-11
    assert int(np.argmin(ls)) == int(np.argmin(z[key + "_ls_energies"]))
    ls2 = runtime.context().line_energies(ad, dx0, sb.handle(), 16, 20)
    assert relerr(ls2, z[key + "_ls_energies"]) < 1e-11
    pr = z[key + "_probes"].reshape(-1, 2)
    assert relerr(Energy.HoppingEnergies(pr, ad, sb), z[key + "_hop_energy"]) < 1e-11
    assert abs(Energy.HoppingEnergy(pr[0], ad, sb) - z[key + "_hop_energy"][0]) < 1e-11 * abs(z[key + "_hop_energy"][0])
    # bond lists: same positions in the same order (bit-exact)
    nn_a, nn_s = Bins.NearestNeighbors(ad, sb)
    ao, so = z[key + "_nn_a_offs"], z[key + "_nn_s_offs"]
    for i in range(n):
        assert np.array_equal(nn_a[i], z[key + "_nn_a_vals"][ao[i]:ao[i + 1]])
        assert np.array_equal(nn_s[i], z[key + "_nn_s_vals"][so[i]:so[i + 1]])
    assert np.array_equal(Growth.SurfaceAtoms(ad, sb), z[key + "_surface_pos"])
    if has_rates:
        sidx = z[key + "_surface_idx"]
        b = Energy.Barriers(ad, sb)
        assert np.array_equal(np.nonzero(b["is_surface"])[0], sidx)
        assert relerr(b["u_appx"][sidx], z[key + "_uappx"]) < 1e-11
        assert np.max(np.abs(b["u_loc"][sidx] - z[key + "_uloc"])) < 1e-11
        assert np.max(np.abs(b["delta_u"][sidx] - z[key + "_deltaU"])) < 1e-10
        assert relerr(b["rate"][sidx], z[key + "_rates"]) < 1e-8
        Rk = Growth.HoppingRates(ad, sb)
        assert relerr(Rk, z[key + "_rates"]) < 1e-8
        if len(sidx):
            i0 = int(sidx[0])
            assert abs(Energy.DeltaU(i0, ad, sb) - z[key + "_deltaU"][0]) < 1e-10
            assert Energy.U_loc_ideal(i0, ad, sb) == z[key + "_uideal"][0]
        # partial sums and selection on the REFERENCE's rates: bit-exact
        ref_rates = z[key + "_rates"]
        assert np.array_equal(np.array(Growth.HoppingPartialSums(ref_rates)), z[key + "_pj"])
        assert Growth.TotalRate(float(z[key + "_Rd"]), ref_rates) == float(z[key + "_Rtot"])
        dense = np.zeros(n)
        dense[sidx] = ref_rates
        flags = np.zeros(n, dtype=np.uint8)
        flags[sidx] = 1
        ctx = runtime.context()
        for u, want in zip(z[key + "_sel_u"][::4], z[key + "_sel_idx"][::4]):
            k, didx, rtot = ctx.select(dense, flags, u)
            assert k == want and rtot == z[key + "_Rtot"]
            if k >= 0:
                assert didx == sidx[k]


def test_relaxation_trace_golden():
    """One full Relaxation (2DGrowth.py:97-166): same line-search count as the reference, same result."""
    from kmcislandgrowth_b200 import Bins, Growth
    z = np.load(os.path.join(GOLDEN, "relax_trace.npz"))
    sb = Bins.PutInBins(Growth.InitSubstrate())
    stats = []
    x = Growth.Relaxation(z["ad0"], sb, stats=stats)
    assert stats[0].status == 0
    assert np.max(np.abs(x - z["result"])) < 1e-6
    assert stats[0].line_searches == z["energies"].shape[0]


def test_trajectory_golden():
    """The reference's seeded KMC loop, event by event from the reference's own states."""
    from kmcislandgrowth_b200 import Growth
    z = np.load(os.path.join(GOLDEN, "trajectory.npz"))
    us, offs, vals = z["uniforms"], z["states_offs"], z["states_vals"]
    sim = Growth.Simulation()
    pos = 0
    report = []
    for t in range(len(z["kinds"])):
        prev = vals[offs[t - 1]:offs[t]] if t > 0 else np.zeros(0)
        nls_ref = int(z["line_searches"][t])
        sim.set_adatoms(prev)
        rec = sim.step(us[pos], us[pos + 1], max_line_searches=0 if nls_ref < 2000 else 3000)
        pos = int(z["uniforms_used"][t])
        assert rec["kind"] == z["kinds"][t] and rec["hop_k"] == z["hop_idx"][t], t
        assert abs(rec["rtot"] - z["rtot"][t]) <= 1e-9 * abs(z["rtot"][t])
        want = vals[offs[t]:offs[t + 1]]
        got = sim.adatoms
        assert got.size == want.size
        if nls_ref < 2000:      # the one 41k-line-search event is chaotic: only its start is checked
            err = float(np.max(np.abs(got - want)))
            report.append((t, nls_ref, rec["line_searches"], err))
            assert err < 1e-5, report
            assert rec["line_searches"] == nls_ref, report
    sim.close()


def _oracle_for(gv):
    p = orc.make_params(gv.E_a, gv.E_s, gv.r_a, gv.r_s, gv.W, gv.Hs, gv.Ha, gv.temperature, gv.deposition_rate,
                        aa_mode=gv.aa_mode, hop_index_mode=gv.hop_index_mode, deposit_mode=gv.deposit_mode)
    return orc.Oracle(p)


@pytest.mark.parametrize("W,Hs,Ha,n,aa", [(18, 10, 10, 150, 1), (48, 12, 10, 700, 0), (48, 12, 10, 700, 1),
                                          (64, 16, 16, 1500, 1), (256, 256, 64, 10000, 1)])
def test_vs_oracle_sizes(W, Hs, Ha, n, aa):
    """bins3x3 / all-pairs modes against the oracle on seeded films up to BASELINE config 2's full size."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(W, Hs, Ha, aa_mode=aa)
    sub, ad = workloads.substrate(gv), workloads.film(gv, n, seed=W + n)
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    try:
        # bin ids bit-exact
        ab = ctx.bins_create(ad)
        cell = ctx.bins_export(ab)[0]
        ocell = o.PutInBins(ad).cell_of_atom()
        ncell = gv.nbins_x * gv.nbins_y
        assert np.array_equal(np.where(cell == ncell, -1, cell), ocell)
        ab.close()
        Fo = o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, osb)
        F = ctx.forces(ad, sb, 3)
        assert np.max(np.abs(F - Fo)) < 1e-9 * max(1.0, np.max(np.abs(Fo)))
        eo = o.RelaxEnergy((ad, osb))
        assert abs(ctx.relax_energy(ad, sb)[0] - eo) < 1e-10 * abs(eo)
        # 20 candidates
        ls = ctx.line_energies(ad, F, sb, 16, 20)
        lso = np.array([o.RelaxEnergy((ad + a * Fo / 20 / 16, osb)) for a in range(0, 20, 7)])
        assert relerr(ls[::7], lso) < 1e-10
        r = ctx.rates(ad, sb)
        rate_o, surf_o, du_o = o.HoppingRatesDense(ad, osb)
        ua, ul, ui, du, na, ns = o.Barriers(ad, osb)
        assert np.array_equal(r["n_a"], na) and np.array_equal(r["n_s"], ns)
        assert np.array_equal(r["is_surface"].astype(bool), surf_o)
        assert np.max(np.abs(r["delta_u"] - du_o)) < 1e-9 * max(1.0, np.max(np.abs(du_o)))
        assert relerr(r["rate"][surf_o], rate_o[surf_o]) < 1e-7
        for u in (0.0, 0.123, 0.5, 0.999999):
            assert ctx.select(rate_o, surf_o.astype(np.uint8), u) == o.Select(rate_o, u)
    finally:
        sb.close()
        ctx.close()


@pytest.mark.parametrize("W,Hs,Ha,n,aa,amp", [(18, 10, 10, 120, 0, 40.0), (18, 10, 10, 120, 1, 40.0), (48, 12, 10, 600, 1, 60.0),
                                                (48, 12, 10, 600, 0, 5.0), (256, 256, 64, 10000, 1, 8.0),
                                                (256, 256, 64, 10000, 1, 0.01)])
def test_line_search_candidates_with_cell_crossings(W, Hs, Ha, n, aa, amp):
    """The 20 candidates x + a*s/20/scale (2DGrowth.py:121,152) with directions large enough that many atoms
    change cell, leave the grid or cross the +-L/2 seam along the line: every candidate energy must equal
    RelaxEnergy of that candidate (per-candidate stencils, Energy.py:133-148 + Bins.py:44-84)."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(W, Hs, Ha, aa_mode=aa)
    sub, ad = workloads.substrate(gv), workloads.film(gv, n, seed=3 * W + n + aa)
    rng = np.random.default_rng(W + n)
    s = rng.normal(0.0, amp, ad.size)
    s[1::2] = np.abs(s[1::2]) * 0.5          # keep atoms from diving into each other / the substrate
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    try:
        for scale in (16, 64):
            got = ctx.line_energies(ad, s, sb, scale, 20)
            cands = [ad + a * s / 20 / scale for a in range(20)]
            pick = range(20) if n <= 1000 else (0, 1, 7, 13, 19)
            want = np.array([o.RelaxEnergy((cands[a], osb)) for a in pick])
            assert relerr(got[list(pick)], want) < 1e-10, (scale, got[list(pick)], want)
            # and the generic batched path agrees as well
            gen = ctx.relax_energy(np.concatenate([cands[a] for a in pick]), sb, n=ad.size // 2, B=len(list(pick)))
            assert relerr(gen, want) < 1e-10
            # how many atoms actually changed cell: the test must exercise the mover path when amp is large
            c0 = o.PutInBins(cands[0]).cell_of_atom()
            c19 = o.PutInBins(cands[19]).cell_of_atom()
            if amp >= 5.0 and scale == 16:
                assert (c0 != c19).sum() > 0
    finally:
        sb.close()
        ctx.close()


@pytest.mark.parametrize("name", ["cfg2_256x256_10k", "cfg3_1024x1024_500k"])
def test_fp32_vs_fp64_tolerance(name):
    """BASELINE configs[2]: one force sweep, one configuration energy and one rate table in fp64 and with fp32
    pair terms, against the fp64 oracle: 1e-6 relative (fp64) / 1e-4 (fp32); bin ids, bond counts and surface
    flags exact in both (they are always evaluated in fp64)."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv, sub, ad = workloads.make(name)
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    Fo = o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, osb)
    eo = o.RelaxEnergy((ad, osb))
    rate_o, surf_o, du_o = o.HoppingRatesDense(ad, osb)
    ua, ul, ui, du, na, ns = o.Barriers(ad, osb)
    ocell = o.PutInBins(ad).cell_of_atom()
    fscale = max(1.0, float(np.max(np.abs(Fo))))
    worst = {}
    for prec, tol in ((0, 1e-6), (1, 1e-4)):
        gv.precision = prec
        ctx = runtime.Context(runtime.params_from(gv), 0)
        sb = ctx.bins_create(sub)
        try:
            ab = ctx.bins_create(ad)
            cell = ctx.bins_export(ab)[0]
            assert np.array_equal(np.where(cell == gv.nbins_x * gv.nbins_y, -1, cell), ocell)
            ab.close()
            F = ctx.forces(ad, sb, 3)
            ferr = float(np.max(np.abs(F - Fo))) / fscale
            eerr = abs(float(ctx.relax_energy(ad, sb)[0]) - eo) / abs(eo)
            r = ctx.rates(ad, sb)
            assert np.array_equal(r["n_a"], na) and np.array_equal(r["n_s"], ns)
            assert np.array_equal(r["is_surface"].astype(bool), surf_o)
            rerr = relerr(r["rate"][surf_o], rate_o[surf_o])
            worst[prec] = (ferr, eerr, rerr)
            assert ferr < tol and eerr < tol and rerr < tol, (prec, worst)
        finally:
            sb.close()
            ctx.close()
    print("fp64 / fp32 errors (forces, energy, rates):", worst)
    assert worst[0][0] < 1e-9 and worst[0][1] < 1e-9       # fp64 is ~1e-13 in practice


@pytest.mark.parametrize("W,Hs,Ha,n,aa,cap", [(64, 16, 16, 1500, 1, 60), (48, 12, 10, 200, 0, 80), (18, 10, 10, 12, 0, 120)])
def test_relaxation_engines_agree(W, Hs, Ha, n, aa, cap):
    """The three Relaxation engines (host-driven launches, persistent kernel, persistent kernel on sorted state)
    and all three launch shapes of the latter (one block, one cluster, cooperative grid) follow the same
    trajectory: equal line-search / restart counts and positions to rounding; the oracle agrees as well."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(W, Hs, Ha, aa_mode=aa)
    sub, ad = workloads.substrate(gv), workloads.film(gv, n, seed=5 * W + n, sigma=0.05)
    res = {}
    for mode in (0, 1, 2):
        ctx = runtime.Context(runtime.params_from(gv), 0)
        ctx.set_relax_mode(mode)
        sb = ctx.bins_create(sub)
        ctx.state_set(ad)
        st = ctx.relax(sb, 16, 1e-4, cap)
        res[mode] = (ctx.state_get(), int(st.line_searches), int(st.restarts), st.maxF)
        sb.close()
        ctx.close()
    o = _oracle_for(gv)
    xo, sto, rc = o.Relaxation(ad, o.PutInBins(sub), max_line_searches=cap)
    for mode in (0, 1, 2):
        assert res[mode][1] == int(sto.line_searches) and res[mode][2] == int(sto.restarts), (mode, res[mode][1:], sto.line_searches)
        assert np.max(np.abs(res[mode][0] - xo)) < 1e-7, mode
    assert np.max(np.abs(res[2][0] - res[0][0])) < 1e-9 and np.max(np.abs(res[1][0] - res[0][0])) < 1e-9


def test_staged_operands_are_bit_identical_to_global_reads(monkeypatch):
    """k_relax2 reads the operands of the energy and force phases from shared-memory copies (TMA bulk copies / cached
    substrate atoms); KMC_DEBUG_E=6 makes it read global memory instead.  Where the operands come from must not change
    a single bit of the result: same line searches, same restarts, identical positions."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(64, 16, 16, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 1500, seed=77, sigma=0.05)
    out = {}
    for dbg in ("0", "6"):
        monkeypatch.setenv("KMC_DEBUG_E", dbg)
        ctx = runtime.Context(runtime.params_from(gv), 0)
        sb = ctx.bins_create(sub)
        ctx.state_set(ad)
        st = ctx.relax(sb, 16, 1e-4, 150)
        out[dbg] = (ctx.state_get().copy(), int(st.line_searches), int(st.restarts), float(st.maxF))
        sb.close()
        ctx.close()
    assert out["0"][1:] == out["6"][1:]
    assert np.array_equal(out["0"][0], out["6"][0])


def test_line_polynomial_and_per_candidate_paths_against_oracle():
    """The line-search energy is summed as per-pair Taylor polynomials for pairs that move little and per candidate for
    the others (kmc_line.cuh: pair_line_poly / pair_line_one).  A film with one strongly displaced atom makes the first
    line searches of the relaxation use BOTH paths (displacements from 1e-6 to 0.3 A along one line); the trajectory must
    follow the oracle: same line-search and restart counts, positions to 1e-7."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(48, 12, 12, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 600, seed=11, sigma=0.01)
    ad = ad.copy()
    ad[2 * 300] += 0.45          # one atom pushed far out of its site: eV/A forces on it and its neighbours
    ad[2 * 300 + 1] += 0.35
    o = _oracle_for(gv)
    xo, sto, rc = o.Relaxation(ad, o.PutInBins(sub), max_line_searches=120)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    ctx.state_set(ad)
    st = ctx.relax(sb, 16, 1e-4, 120)
    x = ctx.state_get()
    sb.close()
    ctx.close()
    assert (int(st.line_searches), int(st.restarts)) == (int(sto.line_searches), int(sto.restarts))
    assert np.max(np.abs(x - xo)) < 1e-7


def test_incremental_rate_table_exact_mode_follows_events():
    """kmc_rates_update with threshold 0 must equal the full sweep bit for bit, event after event, through the index
    shifts of HopAtom (np.delete + append, 2DGrowth.py:241,255) and DepositAdatom (:93); and an event loop that uses the
    incremental table (kmc_set_rate_mode 1, threshold 0) follows the same trajectory as the full-sweep loop."""
    from kmcislandgrowth_b200 import Growth, runtime, workloads
    gv = workloads.constants_for(48, 12, 12, aa_mode=1, hop_index_mode=1, deposit_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 500, seed=21, sigma=0.02)
    rng = np.random.default_rng(5)
    us = rng.uniform(0, 1, (10, 2))
    sims = [Growth.Simulation(adatoms=ad, substrate=sub, params=runtime.params_from(gv)) for _ in range(2)]
    sims[1].ctx.set_rate_mode(1, 0.0)
    kinds = []
    try:
        for s in sims:
            s.ctx.relax(s.sbins, 16, 1e-4, 100)
        first = sims[1].ctx.rates_update(sims[1].sbins, 0.0)
        assert first["recomputed"] == sims[1].ctx.state_count()          # no history yet: full sweep
        again = sims[1].ctx.rates_update(sims[1].sbins, 0.0)
        assert again["recomputed"] == 0                                   # nothing moved: nothing redone
        for t in range(10):
            r0 = sims[0].step(us[t, 0], us[t, 1], 40)
            r1 = sims[1].step(us[t, 0], us[t, 1], 40)
            kinds.append(r0["kind"])
            assert (r0["kind"], r0["line_searches"]) == (r1["kind"], r1["line_searches"])
            assert np.array_equal(sims[0].adatoms, sims[1].adatoms)
            full = sims[0].ctx.rates(sims[0].adatoms, sims[0].sbins)
            inc = sims[1].ctx.rates_update(sims[1].sbins, 0.0)
            assert np.array_equal(inc["rate"], full["rate"]) and np.array_equal(inc["is_surface"], full["is_surface"])
            assert np.array_equal(inc["delta_u"], full["delta_u"])
        assert 1 in kinds
        # a deposition (hop rates dwarf the deposition rate here, so it is placed explicitly): appended atom, no history
        for s_ in sims:
            s_.ctx.deposit_place(s_.sbins, 0.3)
            s_.ctx.relax(s_.sbins, 16, 1e-4, 40)
        full = sims[0].ctx.rates(sims[0].adatoms, sims[0].sbins)
        inc = sims[1].ctx.rates_update(sims[1].sbins, 0.0)
        assert inc["rate"].size == full["rate"].size == 501
        assert np.array_equal(inc["rate"], full["rate"]) and np.array_equal(inc["is_surface"], full["is_surface"])
    finally:
        for s in sims:
            s.close()


def test_incremental_rate_table_touches_only_the_neighbourhood():
    """Moving ONE atom dirties only the cells of its old and new 3x3 stencils: a small fraction of the film is recomputed,
    and the table still equals the full sweep exactly (threshold 0).  With a threshold, sub-threshold jitter of every atom
    recomputes nothing and leaves rates within the corresponding tolerance."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(128, 16, 16, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 1500, seed=8, sigma=0.02)       # 12 layers: every atom inside the bin grid
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    try:
        ctx.state_set(ad)
        assert ctx.rates_update(sb, 0.0)["recomputed"] == 1500
        x = ad.copy()
        top = int(np.argmax(x[1::2]))
        x[2 * top] += 0.37
        x[2 * top + 1] += 0.21
        # kmc_state_move: same atoms, new positions (kmc_state_set would drop the table's history)
        ctx.state_move(x)
        inc = ctx.rates_update(sb, 0.0)
        full = ctx.rates(x, sb)
        assert 0 < inc["recomputed"] < 250, inc["recomputed"]            # ~ the atoms of 2 x 9 cells, not 1500
        assert np.array_equal(inc["rate"], full["rate"]) and np.array_equal(inc["is_surface"], full["is_surface"])
        # sub-threshold jitter everywhere: nothing is redone, rates stay within beta * |dU/dr| * jitter
        rng = np.random.default_rng(3)
        y = x + rng.normal(0, 2e-5, x.shape)
        ctx.state_move(y)
        inc2 = ctx.rates_update(sb, 1e-3)
        assert inc2["recomputed"] == 0
        full2 = ctx.rates(y, sb)
        m = full2["is_surface"].astype(bool)
        assert np.array_equal(inc2["is_surface"].astype(bool), m)
        assert np.max(np.abs(inc2["rate"][m] - full2["rate"][m]) / full2["rate"][m]) < 5e-2
    finally:
        sb.close()
        ctx.close()


def test_grid_barrier_microbenchmark_publishes_data():
    """kmc_barrier_bench: every variant of the device-wide barrier must order the blocks' writes (no stale read of another
    block's word), and the barrier k_relax2 uses must cost microseconds, not milliseconds (profiles/r2_barrier_microbench.md)."""
    from kmcislandgrowth_b200 import runtime, workloads
    ctx = runtime.Context(runtime.params_from(workloads.constants_for(18, 10, 10)), 0)
    try:
        for variant, cluster in ((0, 1), (1, 2), (2, 1), (3, 1)):
            us, stale = ctx.barrier_bench(variant, cluster, 3000)
            assert stale == 0, (variant, cluster, stale)
            assert 0.2 < us < 50.0, (variant, cluster, us)
    finally:
        ctx.close()


def Periodic_wrap(x, L):
    """Periodic.py:8-20 for one coordinate"""
    x = x % L
    return x - L if abs(x) > L / 2 else x


def test_persistent_cell_list_is_edited_in_place_and_equals_a_rebuild():
    """Row n1: the incremental rate table keeps ONE adatom cell list on the context and edits it (kmc_cells.cu) instead of
    re-binning every adatom (Bins.py:91).  After every kind of change -- one atom carried across cell borders, jitter of
    all atoms, a hop (np.delete + append, 2DGrowth.py:241,255), a deposition (:93), a relaxation, many movers at once
    (rebuild fallback) -- the edited list must be word for word the list a rebuild produces (kmc_cells_check == 0), and
    the small changes must really be in-place updates with a handful of edits."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(128, 16, 16, aa_mode=1, hop_index_mode=1, deposit_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, 1500, seed=8, sigma=0.02)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    rng = np.random.default_rng(12)
    try:
        ctx.set_event_mode(0)                                   # host-stepped placement: the path the incremental table runs on
        ctx.state_set(ad)
        ctx.relax(sb, 16, 1e-4, 200)
        ad = ctx.state_get()
        ctx.rates_update(sb, 0.0)                               # no history: full sweep over a scratch list
        ctx.rates_update(sb, 0.0)                               # first incremental update: the persistent list is built
        st = ctx.cells_stats()
        assert (st["full_builds"], st["in_place_updates"]) == (1, 0), st
        assert ctx.cells_check() == 0
        x = ad.copy()
        top = float(np.max(x[1::2]))
        for step in range(6):                                   # one atom at a time is lifted above the film and carried sideways
            k = int(rng.integers(0, 1500))
            x[2 * k] = Periodic_wrap(x[2 * k] + float(rng.choice([-1.0, 1.0])) * gv.bin_size * 1.1, gv.L)
            x[2 * k + 1] = top + 3.0 * (step + 1)
            ctx.state_move(x)
            inc = ctx.rates_update(sb, 0.0)
            full = ctx.rates(x, sb)
            assert np.array_equal(inc["rate"], full["rate"]) and np.array_equal(inc["is_surface"], full["is_surface"])
            st = ctx.cells_stats()
            assert st["full_builds"] == 1 and 1 <= st["last_edits"] <= 2, st
            assert ctx.cells_check() == 0
        y = ad + rng.normal(0, 1e-3, ad.shape)                  # all six back, and everybody moves a little: atoms near a cell wall change cell
        ctx.state_move(y)
        ctx.rates_update(sb, 0.0)
        st = ctx.cells_stats()
        assert ctx.cells_check() == 0 and st["full_builds"] == 1 and 6 <= st["last_edits"] <= 64, st
        n_updates = ctx.cells_stats()["in_place_updates"]
        # index shifts: HopAtom (every later atom moves down one slot in the array), a deposition, relaxations
        cur = ctx.state_get()
        ctx.hop_place(sb, int(np.argmax(cur[1::2])), 0.4)
        ctx.relax(sb, 16, 1e-4, 30)
        inc = ctx.rates_update(sb, 0.0)
        full = ctx.rates(ctx.state_get(), sb)
        assert np.array_equal(inc["rate"], full["rate"])
        assert ctx.cells_check() == 0
        ctx.deposit_place(sb, 0.7)
        ctx.relax(sb, 16, 1e-4, 30)
        ctx.rates_update(sb, 0.0)
        assert ctx.cells_check() == 0
        st = ctx.cells_stats()
        assert st["in_place_updates"] >= n_updates + 2, st
        # thresholded events use the same list
        ctx.set_rate_mode(1, 1e-3)
        for t in range(4):
            ctx.event(sb, float(rng.uniform()), float(rng.uniform()), 30)
            assert ctx.cells_check() == 0
        # hundreds of atoms change cell at once: more edits than are applied in place -> one rebuild, same list
        builds = ctx.cells_stats()["full_builds"]
        z = ctx.state_get()
        z[0::2] += 0.45 * gv.bin_size
        ctx.state_move(z)
        ctx.rates_update(sb, 0.0)
        assert ctx.cells_stats()["full_builds"] == builds + 1
        assert ctx.cells_check() == 0
    finally:
        sb.close()
        ctx.close()


@pytest.mark.parametrize("W,Hs,Ha,n,cap", [(512, 16, 48, 20000, 8), (1024, 16, 96, 80000, 3)])
def test_relaxation_large_chunks_follow_the_oracle(W, Hs, Ha, n, cap):
    """Films whose per-block chunk exceeds the shared-memory caches of k_relax2 (more than 80 atoms per block: stencil
    ranges from global memory; more than 512: several passes per phase; neighbourhoods beyond the staging capacity:
    global reads): the first line searches must follow the oracle exactly like the small cases do."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(W, Hs, Ha, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, n, seed=W + n, sigma=0.03)
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    xo, sto, rc = o.Relaxation(ad, osb, max_line_searches=cap)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    sb = ctx.bins_create(sub)
    try:
        ctx.state_set(ad)
        st = ctx.relax(sb, 16, 1e-4, cap)
        assert (int(st.line_searches), int(st.restarts)) == (int(sto.line_searches), int(sto.restarts))
        assert np.max(np.abs(ctx.state_get() - xo)) < 1e-8
    finally:
        sb.close()
        ctx.close()


def test_checkpoint_and_island_statistics(tmp_path):
    """Simulation.save / load round trip, the KMC clock extension, and island statistics of a GPU trajectory equal to
    those of the reference's trajectory (same uniforms -> same states -> same observables)."""
    from kmcislandgrowth_b200 import Growth, workloads
    z = np.load(os.path.join(GOLDEN, "trajectory.npz"))
    us, offs, vals = z["uniforms"], z["states_offs"], z["states_vals"]
    sim = Growth.Simulation()
    pos = 0
    for t in range(8):
        sim.step(us[pos], us[pos + 1], u_clock=0.5)
        pos = int(z["uniforms_used"][t])
    assert sim.time > 0
    gv = workloads.constants_for(18)
    ref_state = vals[offs[7]:offs[8]]
    assert np.max(np.abs(sim.adatoms - ref_state)) < 1e-5
    assert Growth.IslandStatistics(sim.adatoms, gv) == Growth.IslandStatistics(ref_state, gv)
    path = str(tmp_path / "ckpt.npz")
    sim.save(path)
    sim2 = Growth.Simulation.load(path)
    assert np.array_equal(sim2.adatoms, sim.adatoms) and sim2.t == 8 and sim2.time == sim.time
    a = sim.step(us[pos], us[pos + 1])
    b = sim2.step(us[pos], us[pos + 1])
    assert a["kind"] == b["kind"] and a["line_searches"] == b["line_searches"] and np.array_equal(sim.adatoms, sim2.adatoms)
    sim.close()
    sim2.close()


def test_work_list_overflow_relaunch_and_large_deposit(monkeypatch):
    """(1) a work list that is too small makes the persistent kernel stop and the host grow it and relaunch from the
    untouched state: same result as with a large list; (2) DepositAdatom on the 10k-adatom film (top-down probe)."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv, sub, ad = workloads.make("cfg2_256x256_10k")
    out = {}
    for cap in (None, "64"):
        if cap:
            monkeypatch.setenv("KMC_ITEM_CAP", cap)
        ctx = runtime.Context(runtime.params_from(gv), 0)
        sb = ctx.bins_create(sub)
        ctx.state_set(ad)
        st = ctx.relax(sb, 16, 1e-4, 25)
        out[cap] = (ctx.state_get(), int(st.line_searches))
        if cap is None:
            n0 = ctx.state_count()
            ctx.deposit_place(sb, 0.3)
            assert ctx.state_count() == n0 + 1
            new = ctx.state_get()[-2:]
            o = _oracle_for(gv)
            want = o.DepositPlace(out[None][0], o.PutInBins(sub), 0.3)[-2:]
            assert np.array_equal(new, want)
        sb.close()
        ctx.close()
        monkeypatch.delenv("KMC_ITEM_CAP", raising=False)
    assert out["64"][1] == out[None][1] == 25
    assert np.array_equal(out["64"][0], out[None][0])


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16, 17, 18])
def test_seeded_trajectories_match_the_oracle(seed):
    """KMC trajectories over several seeds (default constants, reference quirk modes, relaxation capped at 1500 line
    searches on both sides): event kinds, hop indices, total rates, adatom counts, positions and the island statistics
    (north_star: island density / size distributions over seeds) agree event by event."""
    from kmcislandgrowth_b200 import Growth, workloads
    o = orc.Oracle()
    osb = o.PutInBins(orc.init_substrate(18, 10, 2.7, 48.6))
    gv = workloads.constants_for(18)
    rng = np.random.default_rng(seed)
    us = rng.uniform(0, 1, (16, 2))
    sim = Growth.Simulation()
    ad = np.zeros(0)
    try:
        for t in range(16):
            ad, kind, hk, rtot, st, rc = o.Event(ad, osb, us[t], max_line_searches=1500)
            rec = sim.step(us[t, 0], us[t, 1], max_line_searches=1500)
            assert (rec["kind"], rec["hop_k"]) == (kind, hk), (seed, t)
            assert abs(rec["rtot"] - rtot) <= 1e-8 * abs(rtot)
            assert rec["line_searches"] == st.line_searches, (seed, t, rec["line_searches"], st.line_searches)
            got = sim.adatoms
            assert got.size == ad.size and np.max(np.abs(got - ad)) < 1e-5, (seed, t)
            assert Growth.IslandStatistics(got, gv) == Growth.IslandStatistics(ad, gv)
            sim.set_adatoms(ad)          # continue both from the oracle's state: one ulp cannot cascade
    finally:
        sim.close()


@pytest.mark.parametrize("aa", [0, 1])
def test_relaxation_with_y_wrapped_stencils(aa):
    """A pillar that reaches the top bin row: there the stencil's y indices wrap ONCE to row 0 (Bins.py:62-65), i.e.
    the top of the pillar 'sees' the bottom substrate rows, and atoms above the grid are invisible to others.  All
    engines must follow the oracle through these non-contiguous stencils (generic stencil walk, three-range items)."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(18, 10, 10, aa_mode=aa)
    sub = workloads.substrate(gv)
    pts = []
    for layer in range(1, 15):            # up to y = 32.7 A: rows 3 .. 6 of 7, the last layers above the top row
        y = layer * gv.r_a * gv.sqrt3 / 2
        for k in range(3):
            pts.append((k * gv.r_a + (gv.r_a / 2 if layer % 2 == 1 else 0.0) - 2.7, y))
    rng = np.random.default_rng(4)
    ad = (np.asarray(pts) + rng.normal(0, 0.03, (len(pts), 2))).reshape(-1)
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    assert max(o.BinIndex(ad[2 * i], ad[2 * i + 1])[1] for i in range(len(pts))) >= gv.nbins_y - 1
    xo, sto, rc = o.Relaxation(ad, osb, max_line_searches=60)
    Fo = o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, osb)
    for mode in (0, 1, 2):
        ctx = runtime.Context(runtime.params_from(gv), 0)
        ctx.set_relax_mode(mode)
        sb = ctx.bins_create(sub)
        assert np.max(np.abs(ctx.forces(ad, sb, 3) - Fo)) < 1e-9 * max(1.0, np.max(np.abs(Fo)))
        assert abs(ctx.relax_energy(ad, sb)[0] - o.RelaxEnergy((ad, osb))) < 1e-10 * abs(o.RelaxEnergy((ad, osb)))
        ctx.state_set(ad)
        st = ctx.relax(sb, 16, 1e-4, 60)
        assert (int(st.line_searches), int(st.restarts)) == (int(sto.line_searches), int(sto.restarts)), mode
        assert np.max(np.abs(ctx.state_get() - xo)) < 1e-7, mode
        sb.close()
        ctx.close()


def test_block_scan_selection_matches_sequential_order():
    """kmc_select's block-scan path (large tables) against the reference-order selection: identical index for
    random u; a mismatch is only legal when r sits within rounding of a partial sum (none expected here)."""
    from kmcislandgrowth_b200 import runtime, workloads
    gv = workloads.constants_for(64, 16, 16, aa_mode=1)
    ctx = runtime.Context(runtime.params_from(gv), 0)
    o = _oracle_for(gv)
    rng = np.random.default_rng(5)
    try:
        for n in (1, 31, 1000, 70001, 300000):
            surf = (rng.uniform(size=n) < 0.3).astype(np.uint8)
            rate = np.where(surf > 0, 10.0 ** rng.uniform(0, 6, n), 0.0)
            if n == 1000:
                rate[surf > 0] *= 1e-9          # every hop far below Rd: mostly deposits
            for mode in (1, 2):
                ctx.set_select_mode(mode)
                for u in np.concatenate([rng.uniform(size=12), [0.0, 1.0 - 2 ** -53]]):
                    got = ctx.select(rate, surf, u)
                    want = o.Select(rate, u)
                    assert got[:2] == want[:2], (n, mode, u, got, want)
                    assert abs(got[2] - want[2]) <= 1e-12 * want[2]
        ctx.set_select_mode(0)
    finally:
        ctx.close()


def test_edge_cases():
    """empty / single-atom inputs and atoms outside the cell grid"""
    from kmcislandgrowth_b200 import Bins, Energy, Growth, Periodic, runtime
    sb = Bins.PutInBins(Growth.InitSubstrate())
    assert Energy.AdatomAdatomForces(np.array([])).size == 0
    assert Energy.RelaxEnergy((np.array([]), sb)) == 0
    assert Growth.HoppingRates(np.array([]), sb) == []
    assert Growth.SurfaceAtoms(np.array([]), sb).size == 0
    assert Periodic.Distances(np.array([]), np.array([1.0, 2.0])).size == 0
    o = orc.Oracle()
    osb = o.PutInBins(orc.init_substrate(18, 10, 2.7, 48.6))
    one = np.array([3.3, 2.5])
    assert abs(Energy.RelaxEnergy((one, sb)) - o.RelaxEnergy((one, osb))) < 1e-12
    assert np.allclose(Energy.TotalForces(one, sb), o.AdatomSubstrateForces(one, osb), rtol=1e-11)
    # unwrapped line-search candidates: x beyond +-L/2 and y above the grid (Bins.py:60-65 single wrap)
    ad = np.array([24.35, 2.4, -24.4, 2.4, 0.0, 2.4, 2.7, 2.4, 5.0, 34.0, 7.0, 36.0])
    e, eo = Energy.RelaxEnergy((ad, sb)), o.RelaxEnergy((ad, osb))
    assert abs(e - eo) < 1e-11 * abs(eo)
    F = Energy.TotalForces(ad, sb)
    Fo = o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, osb)
    assert np.max(np.abs(F - Fo)) < 1e-10 * max(1.0, np.max(np.abs(Fo)))
    b = Energy.Barriers(ad, sb)
    ua, ul, ui, du, na, ns = o.Barriers(ad, osb)
    assert np.array_equal(b["n_a"], na) and np.array_equal(b["n_s"], ns)
    assert relerr(b["delta_u"], du, 1.0) < 1e-11      # atoms 0/1 sit 0.15 A apart across the seam: DeltaU ~ 5e15
    # bins3x3 mode with atoms outside the grid (invisible as neighbours, Bins.py:24-42 + :57-65)
    from kmcislandgrowth_b200 import Constants
    Constants.aa_mode = 1
    try:
        o1 = orc.Oracle(aa_mode=1)
        e1 = Energy.RelaxEnergy((ad, sb))
        eo1 = o1.RelaxEnergy((ad, o1.PutInBins(orc.init_substrate(18, 10, 2.7, 48.6))))
        assert abs(e1 - eo1) < 1e-11 * abs(eo1)
    finally:
        del Constants.aa_mode


# ---- the headline configuration itself (BASELINE.json configs[1]: 256x256 substrate, 10 000 adatoms) ------------------
def _bench_fixture(name="cfg2_256x256_10k"):
    return json.load(open(os.path.join(GOLDEN, "bench_%s.json" % name)))


def test_headline_events_follow_the_oracle_uncapped():
    """The bench trajectory itself: from the bench's relaxed start state, the first events of the seeded stream are run
    UNCAPPED on the GPU (kmc_event: one device-resident chain) and on the CPU oracle.  Same event kind and hop index, same
    number of line searches and restarts in every relaxation (each argmin of each line search agrees, or the counts would
    diverge), positions within 1e-7.  The events cover >= 300 line searches with restarts and re-sorts of the cell-sorted
    state, and the oracle confirms the first entries of the fixture the reference arm of bench.py extrapolates with
    (tests/golden/bench_cfg2_256x256_10k.json: recorded on the GPU)."""
    from kmcislandgrowth_b200 import Growth, runtime, workloads
    gv, sub, ad0 = workloads.make("cfg2_256x256_10k")
    fx = _bench_fixture()
    sim = Growth.Simulation(adatoms=ad0, substrate=sub, device=0, params=runtime.params_from(gv))
    st, total = workloads.bench_prerelax(sim.ctx, sim.sbins)
    assert int(st.status) == 0, "the set-up relaxation must converge (%d line searches)" % total
    if "prerelax_line_searches" in fx:
        assert total == fx["prerelax_line_searches"]
    o = _oracle_for(gv)
    orc.set_num_threads(os.cpu_count() or 1)
    osb = o.PutInBins(sub)
    us = workloads.bench_uniforms(8)
    state = sim.adatoms
    n_ls = n_restarts = n_resorts = 0
    budget = int(os.environ.get("KMC_TEST_HEADLINE_LS", "1400"))      # CPU cost ~55 ms per line search on 16 cores
    done = 0
    for s in range(8):
        if done >= 2 and n_ls + fx["line_searches"][s] > budget:
            break
        rec = sim.step(us[s, 0], us[s, 1], 0)
        diag = sim.ctx.relax_diag()
        want, kind, hk, rtot, ost, rc = o.Event(state, osb, us[s], 0)
        assert rc == 0 and rec["status"] == 0
        assert rec["kind"] == kind and rec["hop_k"] == hk, s
        assert abs(rec["rtot"] - rtot) <= 1e-9 * abs(rtot)
        assert rec["line_searches"] == int(ost.line_searches), (s, rec["line_searches"], int(ost.line_searches))
        assert rec["restarts"] == int(ost.restarts), s
        if "prerelax_line_searches" in fx:          # fixture of the current bench protocol (converged set-up relaxation)
            assert rec["line_searches"] == fx["line_searches"][s] and rec["kind"] == fx["kinds"][s], "bench fixture entry %d is stale" % s
        state = sim.adatoms
        assert state.size == want.size
        assert float(np.max(np.abs(state - want))) < 1e-7, s
        n_ls += rec["line_searches"]
        n_restarts += rec["restarts"]
        n_resorts += diag["resorts"]
        done += 1
    assert done >= 2 and n_ls >= 300, (done, n_ls)
    assert n_restarts >= 1 and n_resorts >= 1, (n_restarts, n_resorts)
    sim.close()


def test_event_modes_agree():
    """device-resident event chain (default) == host-stepped placement (first generation), bit for bit, on hops and deposits"""
    from kmcislandgrowth_b200 import Growth, runtime, workloads
    for W, Hs, Ha, n, aa in [(18, 10, 10, 30, 0), (64, 16, 16, 900, 1)]:
        # (mapped hop index and top-down deposition for both: a capped relaxation after the reference's in-film
        # deposition quirk throws atoms out of the grid, after which nothing can be deposited any more)
        gv = workloads.constants_for(W, Hs, Ha, aa_mode=aa, hop_index_mode=1, deposit_mode=1)
        sub = workloads.substrate(gv)
        ad = workloads.film(gv, n, 5)
        us = np.random.default_rng(77).uniform(0, 1, (10, 2))
        us[3, 0] = 0.999999999           # force deposits as well (the film's hop rates dwarf Rd)
        us[7, 0] = 1.0 - 2 ** -53
        outs = []
        for mode in (1, 0):
            sim = Growth.Simulation(adatoms=ad, substrate=sub, device=0, params=runtime.params_from(gv))
            sim.set_modes(event_mode=mode)
            recs = [sim.step(us[s, 0], us[s, 1], 40) for s in range(10)]
            outs.append((recs, sim.adatoms))
            sim.close()
        for a, b in zip(outs[0][0], outs[1][0]):
            assert (a["kind"], a["hop_k"], a["line_searches"], a["restarts"]) == (b["kind"], b["hop_k"], b["line_searches"], b["restarts"])
            assert a["rtot"] == b["rtot"]
        assert np.array_equal(outs[0][1], outs[1][1])
        assert {r["kind"] for r in outs[0][0]} == {0, 1}, "the event mix must cover hops and deposits"


def test_batched_replicas_equal_sequential_runs():
    """kmc_events_batch (configs[3]): R replicas with their own constants and uniform streams advanced together give
    exactly the trajectories of R sequential simulations"""
    from kmcislandgrowth_b200 import Growth, replicas, runtime
    grid = replicas.sweep_grid(2, 3)
    n_events = 12
    seq = []
    for r in range(len(grid)):
        gv = replicas.replica_constants(r, grid)
        sim = Growth.Simulation(adatoms=None, substrate=replicas.workloads.substrate(gv), device=0, params=runtime.params_from(gv))
        us = np.random.default_rng(1000 + r).uniform(0, 1, (n_events, 2))
        recs = [sim.step(us[t, 0], us[t, 1], 0) for t in range(n_events)]
        seq.append(([(x["kind"], x["hop_k"], x["line_searches"]) for x in recs], sim.adatoms))
        sim.close()
    batch = replicas.ReplicaBatch(list(range(len(grid))), grid, device=0)
    out = batch.run(n_events)
    for r in range(len(grid)):
        assert out["log"][r] == seq[r][0]
        assert np.array_equal(out["adatoms"][r], seq[r][1])
    batch.close()


def test_local_relaxation_matches_reference_semantics():
    """Growth.LocalRelaxation (2DGrowth.py:168-199): stencil atoms relaxed on their own (scale 4), then a global
    relaxation when max|F|^2 > 1e-3 -- against the same composition on the oracle's primitives"""
    from kmcislandgrowth_b200 import Bins, Growth, workloads
    gv = workloads.constants_for(18)
    o = _oracle_for(gv)
    sub = Growth.InitSubstrate()
    sb, osb = Bins.PutInBins(sub), o.PutInBins(sub)
    rng = np.random.default_rng(3)
    n = 14
    ad = np.empty(2 * n)
    ad[0::2] = (np.arange(n) % 9) * 2.7 - 12.0 + rng.normal(0, 0.05, n)
    ad[1::2] = (np.arange(n) // 9 + 1) * 2.338 + rng.normal(0, 0.05, n)
    around = ad[4:6].copy()
    stats = []
    got = Growth.LocalRelaxation(ad, sb, around, stats=stats)
    # the same steps on the oracle
    nearby = np.asarray(o.NearbyAtoms(around[0], around[1], o.PutInBins(ad)))
    D = o.Distances(nearby, ad)
    idx = sorted([int(np.nonzero(row == 0)[0][0]) for row in D], reverse=True)
    rest, relaxing = ad.copy(), []
    for i in idx:
        relaxing += [rest[2 * i], rest[2 * i + 1]]
        rest = np.delete(np.delete(rest, 2 * i), 2 * i)
    x, st1, rc = o.Relaxation(np.array(relaxing), osb, scale=4)
    allx = np.append(rest, x)
    F = o.AdatomAdatomForces(allx) + o.AdatomSubstrateForces(allx, osb)
    if float(np.max(F[0::2] ** 2 + F[1::2] ** 2)) > 1e-3:
        allx, st2, rc = o.Relaxation(allx, osb, scale=4)
        assert len(stats) == 2 and int(stats[1].line_searches) == int(st2.line_searches)
    assert int(stats[0].line_searches) == int(st1.line_searches)
    assert got.size == allx.size and float(np.max(np.abs(got - allx))) < 1e-6


def test_device_domain_world1_matches_the_oracle():
    """domain.DeviceDomain (the device-resident decomposition of configs[4]) with one rank, config-5 geometry scaled down:
    sweeps against the oracle, the host-driven decomposed relaxation against the persistent kernel, one event."""
    import torch
    from kmcislandgrowth_b200 import domain, runtime, workloads
    gv = workloads.constants_for(512, 64, 10, aa_mode=1, hop_index_mode=1, deposit_mode=1)
    sub = workloads.substrate(gv)
    ad = workloads.film(gv, 1100, 5)
    o = _oracle_for(gv)
    osb = o.PutInBins(sub)
    dd = domain.DeviceDomain(gv, sub, ad, 0, 1, 0)
    sw = dd.sweep()
    Fo = (o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, osb)).reshape(-1, 2)
    assert np.max(np.abs(sw["F"].cpu().numpy() - Fo)) < 1e-9 * np.max(np.abs(Fo))
    eo = o.RelaxEnergy((ad, osb))
    assert abs(float(sw["energy"]) - eo) < 1e-10 * abs(eo)
    rate_o, surf_o, du_o = o.HoppingRatesDense(ad, osb)
    assert np.array_equal(sw["surf_dense"].cpu().numpy().astype(bool), surf_o)
    assert np.allclose(sw["rate_dense"].cpu().numpy(), rate_o, rtol=1e-8)
    k, dense, rtot = dd.select(0.4321, sw)
    assert (k, dense) == tuple(o.Select(rate_o, 0.4321)[:2])
    # decomposed line searches: candidate energies against the oracle, then a capped relaxation against k_relax2
    F = sw["F"].contiguous()
    lx, ls = dd.exchange(dd.x, F)
    e = dd.ctx.domain_line_energies(lx.view(-1), ls.view(-1), dd.n_own, dd.sb, 16, 20).cpu().numpy()
    for a in (0, 7, 19):
        cand = ad + a * F.cpu().numpy().reshape(-1) / 20 / 16
        ea = o.RelaxEnergy((cand, osb))
        assert abs(e[a] - ea) < 1e-10 * abs(ea), a
    e19 = dd.ctx.domain_line_energies(lx.view(-1), ls.view(-1), dd.n_own, dd.sb, 16, 19).cpu().numpy()      # K != 20: batched path
    assert np.max(np.abs(e19 - e[:19])) < 1e-9 * np.max(np.abs(e))
    st = dd.relax(16, 1e-4, 12)
    ref = runtime.Context(runtime.params_from(gv), 0)
    rsb = ref.bins_create(sub)
    ref.state_set(ad)
    st2 = ref.relax(rsb, 16, 1e-4, 12)
    assert st["line_searches"] == int(st2.line_searches) == 12
    assert np.max(np.abs(dd.gather_positions().cpu().numpy().reshape(-1) - ref.state_get())) < 1e-8
    ev = dd.event(0.37, 0.31, 8)                    # one event on the decomposed state (placement on the gathered state)
    assert ev["kind"] in (0, 1) and dd.n_global == 1100 + (1 - ev["kind"]) and ev["line_searches"] == 8

    rsb.close(); ref.close(); dd.close()
    torch.cuda.synchronize()
