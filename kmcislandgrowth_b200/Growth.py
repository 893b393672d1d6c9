"""Host mirror of the functions of the reference driver 2DGrowth.py that sit on (or directly drive)
the hot path: substrate fixture, rate table, selection, event application and relaxation.

Every function keeps the reference's name, arguments and return value and cites its lines; the
arithmetic runs on the GPU through libkmcb200.  `Simulation` is the fused form of the loop
2DGrowth.py:317-329 on a device-resident adatom array.

Randomness: the reference draws `random.random()` (2DGrowth.py:76,322) and `random.randint`
(:248).  Here every draw is a `random.random()` uniform u (randint(0, n-1) -> int(u*n)), or comes
from an explicit uniform stream, so that runs are reproducible across Python versions.
"""
import random

import numpy as np

from . import Bins, Energy, Periodic, runtime


def _gv():
    return runtime.constants()


def InitSubstrate():
    """2DGrowth.py:18-37 -- hexagonal W x Hs lattice below y = 0, wrapped into the box (input fixture)."""
    gv = _gv()
    i = np.arange(gv.Hs)
    x = np.arange(gv.W)
    X = x[None, :] * gv.r_s + np.where(i[:, None] % 2 == 1, gv.r_s / 2, 0.0)
    Y = np.broadcast_to((-i * gv.r_s * gv.sqrt3 / 2)[:, None], X.shape)
    R = np.empty(2 * gv.W * gv.Hs)
    R[0::2] = X.reshape(-1)
    R[1::2] = Y.reshape(-1)
    return Periodic.PutAllInBox(R)


def DepositionRate():
    """2DGrowth.py:39-43 (the second return at :44 is unreachable)"""
    gv = _gv()
    return 6.0 * 18 / gv.W * gv.deposition_rate


def SurfaceAtoms(adatoms, substrate_bins):
    """2DGrowth.py:46-62 -- positions of adatoms with n_a (incl. self) + n_s < 6."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return np.array([])
    na, ns = Bins.BondCounts(adatoms, substrate_bins)
    mask = (na + ns) < 6
    return adatoms.reshape(-1, 2)[mask].reshape(-1)


def HoppingRates(adatoms, substrate_bins):
    """2DGrowth.py:201-215 -- omega*exp(DeltaU*beta) for every surface adatom, ordered by adatom index."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return []
    b = Energy.Barriers(adatoms, substrate_bins)
    return [float(r) for r, s in zip(b["rate"], b["is_surface"]) if s]


def HoppingPartialSums(Rk):
    """2DGrowth.py:217-221 -- [0] + left-to-right cumulative sums (device k_select, sequential order)."""
    Rk = np.asarray(Rk, dtype=np.float64)
    if Rk.size == 0:
        return [0]
    k, dense, rtot, pj = runtime.context().select(Rk, np.ones(Rk.size, dtype=np.uint8), 0.0, want_pj=True)
    return [0] + [float(v) for v in pj[1:]]


def TotalRate(Rd, Rk):
    """2DGrowth.py:223-227"""
    Rk = np.asarray(Rk, dtype=np.float64)
    if Rk.size == 0:
        return Rd + 0
    ctx = runtime.context()
    old = ctx.params
    p = runtime.params_from(_gv(), Rd=float(Rd))
    ctx.set_params(p)
    try:
        k, dense, rtot = ctx.select(Rk, np.ones(Rk.size, dtype=np.uint8), 0.0)
    finally:
        ctx.set_params(old)
    return rtot


def SelectEvent(adatoms, substrate_bins, u):
    """2DGrowth.py:318-329 up to the branch: -> (k, dense_index, Rtot); k = -1 means deposit, otherwise k is
    the index into the surface list that the reference hands to HopAtom (:327)."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    ctx = runtime.context()
    if adatoms.size == 0:
        return -1, -1, DepositionRate()
    b = Energy.Barriers(adatoms, substrate_bins)
    return ctx.select(b["rate"], b["is_surface"], u)


def Relaxation(adatoms, substrate_bins, scale=16, threshold=1e-4, max_line_searches=0, stats=None):
    """2DGrowth.py:97-166 -- 20-point line search pseudo-CG with the reference's restart rules."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    if adatoms.size == 0:
        return adatoms
    ctx = runtime.context()
    ctx.state_set(adatoms)
    st = ctx.relax(Bins._as_bins(substrate_bins).handle(), scale, threshold, max_line_searches)
    if stats is not None:
        stats.append(st)
    return ctx.state_get()


def DepositAdatom(adatoms, substrate_bins, u=None, max_line_searches=0, stats=None):
    """2DGrowth.py:64-95 -- drop a new adatom at x = PutInBox(u*L) until it is within 2 r of something, relax."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    ctx = runtime.context()
    sb = Bins._as_bins(substrate_bins).handle()
    ctx.state_set(adatoms)
    ctx.deposit_place(sb, random.random() if u is None else u)
    st = ctx.relax(sb, 16, 1e-4, max_line_searches)
    if stats is not None:
        stats.append(st)
    return ctx.state_get()


def HopAtom(i, adatoms, substrate_bins, u=None, max_line_searches=0, stats=None):
    """2DGrowth.py:229-257 -- move adatom i next to a random surface atom within 3 r_a (best of 6 hex sites), relax."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    ctx = runtime.context()
    sb = Bins._as_bins(substrate_bins).handle()
    ctx.state_set(adatoms)
    ctx.hop_place(sb, i, random.random() if u is None else u)
    st = ctx.relax(sb, 16, 1e-4, max_line_searches)
    if stats is not None:
        stats.append(st)
    return ctx.state_get()


def IslandStatistics(adatoms, gv=None):
    """Islands = connected components of the adatoms under the reference's bond rule d < 1.2 r_a with the
    x-periodic minimum image (Bins.py:100, Periodic.py:27-36).  Analysis helper (host side, not on the hot path):
    -> dict(n_islands, sizes (descending), density per unit length, largest)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    gv = _gv() if gv is None else gv
    ad = np.asarray(adatoms, dtype=np.float64).reshape(-1, 2)
    n = ad.shape[0]
    if n == 0:
        return {"n_islands": 0, "sizes": [], "density": 0.0, "largest": 0}
    pts = ad.copy()
    pts[:, 0] = np.mod(pts[:, 0] + gv.L / 2, gv.L)
    span = float(pts[:, 1].max() - pts[:, 1].min()) + 10 * gv.r_a
    pts[:, 1] -= pts[:, 1].min()
    tree = cKDTree(pts, boxsize=[gv.L, max(span, 4 * gv.r_a) * 4])
    pairs = tree.query_pairs(1.2 * gv.r_a, output_type="ndarray")
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n)) if len(pairs) else coo_matrix((n, n))
    ncomp, lab = connected_components(g, directed=False)
    sizes = sorted(np.bincount(lab).tolist(), reverse=True)
    return {"n_islands": int(ncomp), "sizes": sizes, "density": ncomp / float(gv.L), "largest": int(sizes[0])}


class Simulation(object):
    """The KMC loop of 2DGrowth.py:305-329 on a device-resident trajectory.

        sim = Simulation()              # constants from the Constants module
        sim.step(u0, u1)                # one event: rates -> select -> hop/deposit -> relaxation
    """

    def __init__(self, adatoms=None, substrate=None, device=None, params=None):
        gv = _gv()
        self.params = params if params is not None else runtime.params_from(gv)
        self.ctx = runtime.Context(self.params, runtime.default_device() if device is None else device)
        sub = InitSubstrate() if substrate is None else np.asarray(substrate, dtype=np.float64)
        self.substrate = sub
        self.sbins = self.ctx.bins_create(sub)
        self.ctx.state_set(np.zeros(0) if adatoms is None else np.asarray(adatoms, dtype=np.float64))
        self.t = 0
        self.time = 0.0          # physical KMC clock (extension: the reference only counts events, 2DGrowth.py:316,343)
        self.log = []

    @property
    def adatoms(self):
        return self.ctx.state_get()

    def set_adatoms(self, adatoms):
        self.ctx.state_set(np.asarray(adatoms, dtype=np.float64))

    def step(self, u0=None, u1=None, max_line_searches=0, u_clock=None):
        """one event; with u_clock in (0, 1] the residence time dt = -ln(u_clock)/Rtot is added to self.time"""
        u0 = random.random() if u0 is None else u0
        u1 = random.random() if u1 is None else u1
        kind, hk, rtot, st = self.ctx.event(self.sbins, u0, u1, max_line_searches)
        self.t += 1
        if u_clock is not None and rtot > 0:
            self.time += -np.log(u_clock) / rtot
        rec = dict(t=self.t, kind=kind, hop_k=hk, rtot=rtot, line_searches=st.line_searches, restarts=st.restarts,
                   maxF=st.maxF, status=st.status)
        self.log.append(rec)
        return rec

    def save(self, path):
        """checkpoint: positions, substrate, constants, counters (npz)"""
        p = self.params
        np.savez_compressed(path, adatoms=self.adatoms, substrate=self.substrate, t=self.t, time=self.time,
                            params=np.array([getattr(p, f) for f, _ in p._fields_], dtype=np.float64))

    @classmethod
    def load(cls, path, device=None):
        from ._lib import KmcParams
        z = np.load(path)
        p = KmcParams()
        for (f, ct), v in zip(p._fields_, z["params"]):
            setattr(p, f, int(v) if "int" in ct.__name__ else float(v))
        sim = cls(adatoms=z["adatoms"], substrate=z["substrate"], device=device, params=p)
        sim.t, sim.time = int(z["t"]), float(z["time"])
        return sim

    def close(self):
        self.sbins.close()
        self.ctx.close()
