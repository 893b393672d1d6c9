"""Drop-in for the reference's Energy.py: Lennard-Jones pair energies, forces and hop barriers,
evaluated by the sm_100a kernels of libkmcb200.  Same names / arguments / return values as the
reference; the live functions of Energy.py are all here, the dead or broken-arity ones
(Energy.py:97-109,150-168) are provided with their evident meaning.
"""
import numpy as np

from . import Bins, runtime
from ._lib import KMC_F_AA, KMC_F_AS


def _gv():
    return runtime.constants()


def _sb(substrate_bins):
    return Bins._as_bins(substrate_bins).handle()


_EMPTY = Bins.DeviceBins(np.zeros(0))


def _empty_bins():
    return _EMPTY.handle()


def _pair_ctx(E_ij, r_ij):
    """The context with its adatom-adatom constants set to (E_ij, r_ij) and all-pairs sums: the generic
    Pairwise* entry points take the pair constants as arguments (Energy.py:10,26,31,49)."""
    ctx = runtime.context()
    if E_ij == ctx.params.E_a and r_ij == ctx.params.r_a and ctx.params.aa_mode == 0:
        return ctx, None
    p = runtime.params_from(_gv(), E_a=float(E_ij), r_a=float(r_ij), aa_mode=0)
    old = ctx.params
    ctx.set_params(p)
    return ctx, old


def PairwisePotentials(Ri, Rjs, E_ij, r_ij):
    """Energy.py:26-29 -- list of LJ energies between Ri and every atom of Rjs (0 where d <= 0)."""
    Ri = np.asarray(Ri, dtype=np.float64)
    Rjs = np.asarray(Rjs, dtype=np.float64)
    m = Rjs.size // 2
    if m == 0:
        return []
    ctx, old = _pair_ctx(E_ij, r_ij)
    try:
        # one probe per neighbour against the single-atom set [Ri] gives the individual terms
        # (LJ is symmetric in the pair)
        e_aa, e_as, _, _ = ctx.probe(Rjs, Ri[:2], _empty_bins())
        out = [float(v) for v in e_aa]
    finally:
        if old is not None:
            ctx.set_params(old)
    return out


def PairwisePotential(Ri, Rj, E_ij, r_ij):
    """Energy.py:10-24"""
    return PairwisePotentials(Ri, Rj, E_ij, r_ij)[0]


def PairwiseForces(Ri, Rjs, E_ij, r_ij):
    """Energy.py:49-60 -- sum of LJ forces on Ri from the atoms of Rjs (pairs closer than 1e-2 skipped)."""
    Ri = np.asarray(Ri, dtype=np.float64)
    Rjs = np.asarray(Rjs, dtype=np.float64)
    if Rjs.size == 0:
        return np.array([0., 0.])
    ctx, old = _pair_ctx(E_ij, r_ij)
    try:
        # force on Ri = AA force on atom 0 of the set [Ri] + Rjs minus nothing else (self pair is skipped, r < 1e-2)
        allp = np.concatenate([Ri[:2], Rjs])
        F = ctx.forces(allp, _empty_bins(), KMC_F_AA)
    finally:
        if old is not None:
            ctx.set_params(old)
    return np.array([F[0], F[1]])


def PairwiseForce(Ri, Rj, E_ij, r_ij):
    """Energy.py:31-47 (dead in the reference; differs from PairwiseForces only in the r == 0 test)"""
    return PairwiseForces(Ri, Rj, E_ij, r_ij)


def AdatomAdatomForces(adatoms):
    """Energy.py:62-67 -- LJ force on every adatom from all adatoms (aa_mode 0) or its bin stencil (aa_mode 1)."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return np.array([])
    return runtime.context().forces(adatoms, _empty_bins(), KMC_F_AA)


def AdatomSubstrateForces(adatoms, substrate_bins):
    """Energy.py:69-75 -- LJ force on every adatom from the substrate atoms of its bin stencil."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return np.array([])
    return runtime.context().forces(adatoms, _sb(substrate_bins), KMC_F_AS)


def TotalForces(adatoms, substrate_bins):
    """AdatomAdatomForces + AdatomSubstrateForces in one launch (the sum formed at 2DGrowth.py:120,128,159)."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return np.array([])
    return runtime.context().forces(adatoms, _sb(substrate_bins), KMC_F_AA | KMC_F_AS)


def AdatomSurfaceSubsetEnergy(adatom, substrate_subset):
    """Energy.py:77-81"""
    gv = _gv()
    return sum(PairwisePotentials(adatom, substrate_subset, gv.E_as, gv.r_as))


def AdatomSurfaceEnergy(adatom, substrate_bins):
    """Energy.py:83-95 -- LJ energy between one adatom and the substrate atoms of its bin stencil."""
    adatom = np.asarray(adatom, dtype=np.float64)
    e_aa, e_as, _, _ = runtime.context().probe(adatom[:2], None, _sb(substrate_bins))
    return float(e_as[0])


def ArbitraryAdatomEnergy(Ri, adatoms):
    """Energy.py:111-119 -- LJ energy of a point against a set of adatoms (list of [x,y] or flat array)."""
    Ri = np.asarray(Ri, dtype=np.float64)
    ad = np.asarray(adatoms, dtype=np.float64).reshape(-1)
    if ad.size == 0:
        return 0
    e_aa, _, _, _ = runtime.context().probe(Ri[:2], ad, _empty_bins())
    return float(e_aa[0])


def AdatomAdatomEnergy(i, adatoms):
    """Energy.py:97-109 (dead in the reference: it indexes the flat array as if it were 2-D)"""
    ad = np.asarray(adatoms, dtype=np.float64).reshape(-1)
    return ArbitraryAdatomEnergy(ad[2 * i:2 * i + 2], ad)


def TotalEnergy(adatoms, substrate_bins):
    """Energy.py:121-131 (loop bound quirk kept; the driver ignores the value, 2DGrowth.py:116)"""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return 0
    return float(runtime.context().total_energy(adatoms, _sb(substrate_bins))[0])


def RelaxEnergy(args):
    """Energy.py:133-148 -- configuration energy: sum_{i<j} LJ_aa + sum_i substrate-stencil energy."""
    (xn, substrate_bins) = args
    xn = np.asarray(xn, dtype=np.float64)
    if xn.size == 0:
        return 0
    return float(runtime.context().relax_energy(xn, _sb(substrate_bins), n=xn.size // 2, B=1)[0])


def RelaxEnergyBatch(xs):
    """All candidates of one line search in ONE launch: xs = [(positions, substrate_bins), ...] exactly as
    2DGrowth.py:121,152 builds them for pool.map(Energy.RelaxEnergy, xs)."""
    if len(xs) == 0:
        return []
    sb = xs[0][1]
    X = np.ascontiguousarray([np.asarray(x, dtype=np.float64) for (x, _) in xs])
    if X.shape[1] == 0:
        return [0] * len(xs)
    e = runtime.context().relax_energy(X.reshape(-1), _sb(sb), n=X.shape[1] // 2, B=X.shape[0])
    return [float(v) for v in e]


def AdatomSurfaceSubsetForce(adatom, substrate_subset):
    """Energy.py:157-158"""
    gv = _gv()
    return PairwiseForces(adatom, substrate_subset, gv.E_as, gv.r_as)


def AdatomSurfaceForce(adatom, substrate_bins):
    """Energy.py:150-155 (broken arity in the reference; evident meaning)"""
    return AdatomSubstrateForces(np.asarray(adatom, dtype=np.float64)[:2], substrate_bins)


def ArbitraryAdatomForce(Ri, adatoms):
    """Energy.py:164-168"""
    gv = _gv()
    return PairwiseForces(Ri, np.asarray(adatoms, dtype=np.float64).reshape(-1), gv.E_a, gv.r_a)


def AdatomAdatomForce(i, adatoms):
    """Energy.py:160-162"""
    ad = np.asarray(adatoms, dtype=np.float64).reshape(-1)
    return ArbitraryAdatomForce(ad[2 * i:2 * i + 2], ad)


def HoppingEnergy(pos, adatoms, substrate_bins):
    """Energy.py:170-172 -- energy of a trial position against all adatoms + its substrate stencil."""
    pos = np.asarray(pos, dtype=np.float64)
    ad = np.asarray(adatoms, dtype=np.float64)
    e_aa, e_as, _, _ = runtime.context().probe(pos[:2], ad if ad.size else None, _sb(substrate_bins))
    return float(e_aa[0] + e_as[0])


def HoppingEnergies(positions, adatoms, substrate_bins):
    """the six trial sites of HopAtom (2DGrowth.py:249-252) in one launch"""
    P = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
    ad = np.asarray(adatoms, dtype=np.float64)
    e_aa, e_as, _, _ = runtime.context().probe(P, ad if ad.size else None, _sb(substrate_bins))
    return e_aa + e_as


def Barriers(adatoms, substrate_bins):
    """U_appx, U_loc, DeltaU, n_a, n_s, surface flag and rate of EVERY adatom in one pass
    (Energy.py:174-217, 2DGrowth.py:46-62,201-215)."""
    return runtime.context().rates(np.asarray(adatoms, dtype=np.float64), _sb(substrate_bins))


def U_appx(i, adatoms, substrate_bins):
    """Energy.py:174-181"""
    return float(Barriers(adatoms, substrate_bins)["u_appx"][i])


def U_loc(i, adatoms, substrate_bins):
    """Energy.py:183-192"""
    return float(Barriers(adatoms, substrate_bins)["u_loc"][i])


def U_loc_ideal(i, adatoms, substrate_bins):
    """Energy.py:194-202 -- -(n_a E_a + n_s E_as) with n_a counting the atom itself"""
    gv = _gv()
    b = Barriers(adatoms, substrate_bins)
    return -(int(b["n_a"][i]) * gv.E_a + int(b["n_s"][i]) * gv.E_as)


def C():
    """Energy.py:204-208"""
    return 1.0


def DeltaU(i, adatoms, substrate_bins):
    """Energy.py:210-217"""
    return float(Barriers(adatoms, substrate_bins)["delta_u"][i])
