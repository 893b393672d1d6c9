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
