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


def LocalRelaxation(adatoms, substrate_bins, around, max_line_searches=0, stats=None):
    """2DGrowth.py:168-199 (defined by the reference, never called by its loop): the adatoms of the 3x3(+seam) stencil
    around `around` are taken out of the array (highest index first), relaxed ON THEIR OWN with scale=4, appended again;
    when the largest squared force of the whole configuration then exceeds 1e-3 a global Relaxation(scale=4) follows.
    Host logic over the device primitives (NearbyAtoms, Relaxation, forces), operation for operation -- including the
    reference's index lookup by exact zero distance (:185) and its repeated deletes for duplicated stencil entries."""
    adatoms = np.asarray(adatoms, dtype=np.float64)
    ctx = runtime.context()
    adatom_bins = Bins.PutInBins(adatoms)                                                   # :182
    nearby = np.asarray(Bins.NearbyAtoms(around[0], around[1], adatom_bins))                # :183
    idx = []
    if nearby.size:
        D = np.asarray(Periodic.Distances(nearby, adatoms))                                 # :184
        idx = [int(np.nonzero(row == 0)[0][0]) for row in D]
    relaxing = []
    for i in sorted(idx, reverse=True):                                                     # :185-190
        relaxing.append(adatoms[2 * i])
        relaxing.append(adatoms[2 * i + 1])
        adatoms = np.delete(adatoms, 2 * i)
        adatoms = np.delete(adatoms, 2 * i)
    relaxing = np.array(relaxing)
    relaxing = Relaxation(relaxing, substrate_bins, scale=4, max_line_searches=max_line_searches, stats=stats)   # :192
    adatoms = np.append(adatoms, relaxing)                                                  # :193
    if adatoms.size == 0:
        return adatoms
    F = np.asarray(Energy.AdatomAdatomForces(adatoms)) + np.asarray(Energy.AdatomSubstrateForces(adatoms, substrate_bins))   # :194
    maxF = float(np.max(F[0::2] ** 2 + F[1::2] ** 2))                                       # :195 (max over atoms of |F_i|^2)
    if maxF > 1e-3:                                                                         # :196-198
        adatoms = Relaxation(adatoms, substrate_bins, scale=4, threshold=1e-4, max_line_searches=max_line_searches, stats=stats)
    return adatoms


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
        self.rate_mode, self.rate_threshold, self.relax_mode, self.select_mode, self.event_mode = 0, 0.0, 2, 0, 1
        self.restored = None

    def set_modes(self, rate_mode=None, rate_threshold=None, relax_mode=None, select_mode=None, event_mode=None):
        """engine knobs (kmc_set_rate_mode / _relax_mode / _select_mode / _event_mode); remembered for checkpoints"""
        if rate_mode is not None or rate_threshold is not None:
            self.rate_mode = self.rate_mode if rate_mode is None else int(rate_mode)
            self.rate_threshold = self.rate_threshold if rate_threshold is None else float(rate_threshold)
            self.ctx.set_rate_mode(self.rate_mode, self.rate_threshold)
        if relax_mode is not None:
            self.relax_mode = int(relax_mode)
            self.ctx.set_relax_mode(self.relax_mode)
        if select_mode is not None:
            self.select_mode = int(select_mode)
            self.ctx.set_select_mode(self.select_mode)
        if event_mode is not None:
            self.event_mode = int(event_mode)
            self.ctx.set_event_mode(self.event_mode)

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

    # ---- state I/O -----------------------------------------------------------------
    @staticmethod
    def _npz(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"

    def save(self, path, rng=None, uniforms=None, uniform_pos=0):
        """Checkpoint (npz; '.npz' is appended when missing, here and in load): positions, substrate, constants,
        counters, the engine's mode knobs, and the random state -- Python's `random` state (what step() draws from
        when no uniforms are passed, 2DGrowth.py:76,248,322), optionally a numpy Generator `rng` and/or an explicit
        uniform stream with its read position -- so that a resumed run continues the SAME trajectory."""
        import pickle
        p = self.params
        extra = {}
        if rng is not None:
            extra["np_rng"] = np.frombuffer(pickle.dumps(rng.bit_generator.state), dtype=np.uint8)
        if uniforms is not None:
            extra["uniforms"] = np.asarray(uniforms, dtype=np.float64)
        np.savez_compressed(self._npz(path), adatoms=self.adatoms, substrate=self.substrate, t=self.t, time=self.time,
                            params=np.array([getattr(p, f) for f, _ in p._fields_], dtype=np.float64),
                            modes=np.array([self.rate_mode, self.relax_mode, self.select_mode, self.event_mode], dtype=np.int64),
                            rate_threshold=self.rate_threshold, uniform_pos=int(uniform_pos),
                            py_random=np.frombuffer(pickle.dumps(random.getstate()), dtype=np.uint8), **extra)

    @classmethod
    def load(cls, path, device=None, restore_random=True):
        """-> Simulation; `sim.restored` holds dict(np_rng=Generator or None, uniforms=array or None, uniform_pos=int)"""
        import pickle
        from ._lib import KmcParams
        z = np.load(cls._npz(path))
        p = KmcParams()
        for (f, ct), v in zip(p._fields_, z["params"]):
            setattr(p, f, int(v) if "int" in ct.__name__ else float(v))
        sim = cls(adatoms=z["adatoms"], substrate=z["substrate"], device=device, params=p)
        sim.t, sim.time = int(z["t"]), float(z["time"])
        if "modes" in z.files:
            rate_mode, relax_mode, select_mode, event_mode = (int(v) for v in z["modes"])
            sim.set_modes(rate_mode=rate_mode, rate_threshold=float(z["rate_threshold"]), relax_mode=relax_mode,
                          select_mode=select_mode, event_mode=event_mode)
        if restore_random and "py_random" in z.files:
            random.setstate(pickle.loads(z["py_random"].tobytes()))
        rng = None
        if "np_rng" in z.files:
            rng = np.random.default_rng()
            rng.bit_generator.state = pickle.loads(z["np_rng"].tobytes())
        sim.restored = {"np_rng": rng, "uniforms": z["uniforms"] if "uniforms" in z.files else None,
                        "uniform_pos": int(z["uniform_pos"]) if "uniform_pos" in z.files else 0}
        return sim

    def dump_frame(self, path, forces=True):
        """The data of the reference's per-event frames (2DGrowth.py:332-341: atoms + bin grid, PlotAtoms/PlotSubstrate :259-279;
        force-coloured atoms, PlotStresses :290-300) as an npz instead of a PNG (matplotlib is not part of the hot path):
        adatoms, substrate, surface flags, per-atom forces F_aa + F_as and |F|, the bin grid lines, t."""
        ad = self.adatoms
        n = ad.size // 2
        out = {"t": self.t, "adatoms": ad, "substrate": self.substrate}
        p = self.params
        out["bin_edges_x"] = -p.L / 2 + p.bin_size * np.arange(p.nbins_x + 1)
        out["bin_edges_y"] = p.Dmin + p.bin_size * np.arange(p.nbins_y + 1)
        if n:
            out["is_surface"] = np.asarray(self.ctx.rates(ad, self.sbins)["is_surface"])
            if forces:
                F = np.asarray(self.ctx.forces(ad, self.sbins, 3))
                out["forces"] = F
                out["force_norm"] = np.sqrt(F[0::2] ** 2 + F[1::2] ** 2)        # the colour of PlotStresses
        np.savez_compressed(self._npz(path), **out)
        return self._npz(path)

    def close(self):
        self.sbins.close()
        self.ctx.close()
