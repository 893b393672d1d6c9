"""The reference driver's imports (`import Constants as gv; import Periodic; import Bins; import Energy`,
2DGrowth.py:11-14) resolve to the B200 implementation when dropin/ is ahead on sys.path, and the drop-in modules
read the constants of whatever top-level `Constants` module is loaded (the reference's own file)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from conftest import ROOT


def _fake_reference_constants(W):
    """a stand-in for the reference's Constants.py with its exact derivation order (Constants.py:28-46)"""
    from kmcislandgrowth_b200 import Constants as own
    m = types.ModuleType("Constants")
    own.configure(2.083, 2.083, 2.7, 2.7, W, "_Images/", target=m)
    return m


def test_dropin_modules_resolve_and_export_the_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for name in ("Periodic", "Bins", "Energy"):
            sys.modules.pop(name, None)
        P, B, E = (importlib.import_module(n) for n in ("Periodic", "Bins", "Energy"))
        for f in ("PutInBox", "PutAllInBox", "Displacement", "Distance", "Displacements", "Distances"):
            assert callable(getattr(P, f))
        for f in ("BinIndex", "PutInBins", "NearbyBins", "NearbyAtoms", "NearestNeighbors"):
            assert callable(getattr(B, f))
        for f in ("PairwisePotential", "PairwisePotentials", "PairwiseForce", "PairwiseForces", "AdatomAdatomForces",
                  "AdatomSubstrateForces", "AdatomSurfaceSubsetEnergy", "AdatomSurfaceEnergy", "AdatomAdatomEnergy",
                  "ArbitraryAdatomEnergy", "TotalEnergy", "RelaxEnergy", "AdatomSurfaceForce", "AdatomSurfaceSubsetForce",
                  "AdatomAdatomForce", "ArbitraryAdatomForce", "HoppingEnergy", "U_appx", "U_loc", "U_loc_ideal", "C", "DeltaU"):
            assert callable(getattr(E, f)), f
        assert E.C() == 1.0
        import pickle
        bins = B.PutInBins(np.array([0.0, 1.0, 2.7, 1.0]))
        again = pickle.loads(pickle.dumps(bins))          # the reference ships bins through multiprocessing.Pool
        assert np.array_equal(again.R, bins.R)
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for name in ("Periodic", "Bins", "Energy"):
            sys.modules.pop(name, None)


@pytest.mark.gpu
def test_dropin_reads_the_loaded_constants_module():
    from kmcislandgrowth_b200 import Bins, Energy, Growth, Periodic, runtime
    from oracle import oracle as orc
    sys.modules["Constants"] = _fake_reference_constants(24)
    try:
        assert runtime.constants().W == 24
        sub = Growth.InitSubstrate()
        assert sub.size == 2 * 24 * 10
        o = orc.Oracle(W=24)
        assert np.array_equal(sub, orc.init_substrate(24, 10, 2.7, o.p.L))
        sb = Bins.PutInBins(sub)
        ad = np.array([0.0, 2.34, 2.7, 2.34, 32.4, 2.4])
        assert abs(Energy.RelaxEnergy((ad, sb)) - o.RelaxEnergy((ad, o.PutInBins(sub)))) < 1e-11
        assert Periodic.PutInBox(40.0) == o.PutInBox(40.0)
        assert Bins.BinIndex(32.4, 0.0) == o.BinIndex(32.4, 0.0) == (8, 2)       # x == L/2 lands in column nbins_x
        # pool.map replacement and pairwise helpers
        xs = [(ad + a * 0.01, sb) for a in range(20)]
        assert np.allclose(Energy.RelaxEnergyBatch(xs), [o.RelaxEnergy((x, o.PutInBins(sub))) for x, _ in xs], rtol=1e-12)
        assert abs(Energy.PairwisePotential(ad[0:2], ad[2:4], 2.083, 2.7) - orc._lib().orc_pair_potential(2.7, 2.083, 2.7)) < 1e-12
        f = Energy.PairwiseForces(ad[0:2], ad[2:6], 2.083, 2.7)
        fo = o.AdatomAdatomForces(ad)[0:2]
        assert np.allclose(f, fo, rtol=1e-10, atol=1e-12)
    finally:
        del sys.modules["Constants"]
        runtime.reset()


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class _Stream(object):
    """explicit uniform stream standing in for the `random` module inside the reference driver"""

    def __init__(self, us):
        self.us, self.pos = list(us), 0

    def random(self):
        u = self.us[self.pos]
        self.pos += 1
        return u

    def randint(self, a, b):
        return a + int(self.random() * (b - a + 1))


class _BatchPool(object):
    """the `pool` of 2DGrowth.py:16 whose map(Energy.RelaxEnergy, xs) is ONE batched device evaluation"""

    def __init__(self, energy):
        self.energy = energy

    def map(self, f, xs):
        if f is self.energy.RelaxEnergy:
            return [float(v) for v in self.energy.RelaxEnergyBatch(xs)]
        return list(map(f, xs))


@pytest.mark.gpu
def test_reference_driver_runs_on_the_dropin_modules():
    """The reference's OWN DepositAdatom / HopAtom / Relaxation / HoppingRates / HoppingPartialSums / TotalRate
    (2DGrowth.py:39-257, mechanically converted to Python 3 into the git-ignored baseline/_ref by
    oracle/make_ref_driver.py) run over dropin/{Periodic,Bins,Energy}: ten events of the reference's loop
    (2DGrowth.py:317-329) on the golden uniform stream reproduce the golden trajectory of the pure reference."""
    if not os.path.exists(os.path.join(REF_DIR, "RefGrowth.py")):
        pytest.skip("baseline/_ref/RefGrowth.py is absent: run `python oracle/make_ref_driver.py` where /root/reference exists")
    from kmcislandgrowth_b200 import runtime
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "trajectory.npz"))
    us, offs, vals = z["uniforms"], z["states_offs"], z["states_vals"]
    saved_path, saved_argv = list(sys.path), sys.argv
    names = ("Constants", "Periodic", "Bins", "Energy", "RefGrowth")
    saved_mods = {n: sys.modules.pop(n, None) for n in names}
    sys.argv = ["2DGrowth.py"]                      # Constants.py:7-13 reads argv: defaults
    sys.path[:0] = [os.path.join(ROOT, "dropin"), REF_DIR]
    runtime.reset()
    try:
        gv = importlib.import_module("Constants")   # the REFERENCE's constants module
        assert gv.__file__.startswith(REF_DIR) and gv.W == 18
        G = importlib.import_module("RefGrowth")    # the REFERENCE's driver functions
        E, B = sys.modules["Energy"], sys.modules["Bins"]
        assert E.__file__.startswith(os.path.join(ROOT, "dropin")) and B.__file__.startswith(os.path.join(ROOT, "dropin"))
        assert runtime.constants() is gv
        stream = _Stream(us)
        G.random = stream                           # the driver calls random.random() / random.randint()
        literal_pool = G.pool
        substrate = G.InitSubstrate()
        sbins = B.PutInBins(substrate)
        import io
        pos = 0
        for t in range(10):
            adatoms = np.array(vals[offs[t - 1]:offs[t]]) if t > 0 else np.array([])
            stream.pos = pos
            G.pool = literal_pool if t < 5 else _BatchPool(E)      # both forms of the pool.map replacement
            old = sys.stdout
            sys.stdout = io.StringIO()              # the driver prints maxF per line search (2DGrowth.py:161)
            try:
                Rk = G.HoppingRates(adatoms, sbins)                # 2DGrowth.py:318
                pj = G.HoppingPartialSums(Rk)                      # :319
                Rd = G.DepositionRate()                            # :320
                Rtot = G.TotalRate(Rd, Rk)                         # :321
                r = stream.random() * Rtot                         # :322
                hop = [p for p in pj if p > r]                     # :324
                if len(hop) > 0:
                    k = pj.index(hop[0]) - 1
                    adatoms = G.HopAtom(k, adatoms, sbins)         # :327
                    kind = 1
                else:
                    k = -1
                    adatoms = G.DepositAdatom(adatoms, sbins)      # :329
                    kind = 0
            finally:
                sys.stdout = old
            pos = int(z["uniforms_used"][t])
            assert stream.pos == pos, t
            assert kind == z["kinds"][t] and k == z["hop_idx"][t], t
            assert abs(Rtot - z["rtot"][t]) <= 1e-9 * abs(z["rtot"][t])
            want = vals[offs[t]:offs[t + 1]]
            got = np.asarray(adatoms, dtype=np.float64)
            assert got.size == want.size, t
            assert float(np.max(np.abs(got - want))) < 1e-5, t
    finally:
        sys.path[:] = saved_path
        sys.argv = saved_argv
        for n in names:
            sys.modules.pop(n, None)
            if saved_mods[n] is not None:
                sys.modules[n] = saved_mods[n]
        runtime.reset()
