"""Process-wide access to the CUDA library: constants lookup, lazy per-process context, and a thin
object wrapper (`Context`) over the C-ABI of include/kmc_b200.h.

Arrays may be numpy arrays (host memory: the call copies in/out and synchronises) or torch CUDA
tensors (device memory: the call only enqueues kernels on the context's stream; torch is used for
nothing but the allocation).  No CUDA context is created at import time (the reference driver forks
a multiprocessing.Pool right after importing Energy, 2DGrowth.py:14-16); contexts are per process.
"""
import ctypes as C
import os
import sys

import numpy as np

from . import _lib
from ._lib import KMC_DEVICE, KMC_HOST, KmcError, KmcParams, KmcRelaxStats, check

_REQUIRED = ("L", "bin_size", "Dmin", "Dmax", "E_a", "r_a", "E_as", "r_as", "beta", "nbins_x", "nbins_y", "W")


def constants():
    """The module the reference calls `gv`: a top-level `Constants` if one is loaded (the reference's
    own Constants.py next to its driver), else this package's Constants."""
    m = sys.modules.get("Constants")
    if m is not None and all(hasattr(m, k) for k in _REQUIRED):
        return m
    from . import Constants as own
    return own


def params_from(gv, **over):
    """KmcParams from a Constants-like namespace; 2DGrowth.py:43 (Rd) and :210 (omega) included."""
    g = lambda k, d=None: over[k] if k in over else getattr(gv, k, d)
    p = KmcParams()
    p.L, p.bin_size, p.Dmin, p.Dmax = float(g("L")), float(g("bin_size")), float(g("Dmin")), float(g("Dmax"))
    p.E_a, p.r_a, p.E_as, p.r_as = float(g("E_a")), float(g("r_a")), float(g("E_as")), float(g("r_as"))
    p.beta = float(g("beta"))
    p.omega = float(g("omega", 1.0e12))
    p.Rd = float(over["Rd"]) if "Rd" in over else 6.0 * 18 / g("W") * g("deposition_rate", 1.0)
    p.nbins_x, p.nbins_y = int(g("nbins_x")), int(g("nbins_y"))
    p.aa_mode = int(g("aa_mode", 0))
    p.precision = int(g("precision", 0))
    p.hop_index_mode = int(g("hop_index_mode", 0))
    p.deposit_mode = int(g("deposit_mode", 0))
    return p


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _in(a, dtype=np.float64):
    """-> (keepalive, pointer, mem)"""
    if a is None:
        return None, None, None
    if _is_torch(a):
        if not a.is_cuda or not a.is_contiguous():
            raise ValueError("torch arguments must be contiguous CUDA tensors")
        return a, C.c_void_p(a.data_ptr()), KMC_DEVICE
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, C.c_void_p(a.ctypes.data), KMC_HOST


class _Out(object):
    """output allocation on the same side (host/device) as the inputs"""

    def __init__(self, like):
        self.torch = _is_torch(like)
        self.like = like

    def new(self, shape, dtype=np.float64):
        if self.torch:
            import torch
            td = {np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64, np.uint8: torch.uint8}[dtype]
            return torch.empty(shape, dtype=td, device=self.like.device)
        return np.empty(shape, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.data_ptr()) if _is_torch(a) else C.c_void_p(a.ctypes.data)


class BinsHandle(object):
    """Owner of one device cell list (KmcBins*)."""

    def __init__(self, ctx, handle, n):
        self.ctx, self.handle, self.n = ctx, handle, n

    def close(self):
        if self.handle is not None and self.ctx.handle is not None:
            self.ctx.lib.kmc_bins_destroy(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(object):
    """One KmcCtx: one device, one stream, one set of constants."""

    def __init__(self, params, device=0):
        self.lib = _lib.load()
        self.handle = None
        h = C.c_void_p()
        check(self.lib.kmc_create(C.byref(params), int(device), C.byref(h)))
        self.handle = h
        self.params = params
        self.device = int(device)
        self.pid = os.getpid()

    def close(self):
        if self.handle is not None:
            self.lib.kmc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.pid == os.getpid():
                self.close()
        except Exception:
            pass

    # ---- management
    def set_params(self, params):
        check(self.lib.kmc_set_params(self.handle, C.byref(params)))
        self.params = params

    def use_torch_stream(self):
        import torch
        check(self.lib.kmc_set_stream(self.handle, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def synchronize(self):
        check(self.lib.kmc_synchronize(self.handle))

    def launch_count(self):
        return int(self.lib.kmc_launch_count(self.handle))

    def timing(self, on):
        check(self.lib.kmc_timing_enable(self.handle, 1 if on else 0))

    def timing_reset(self):
        check(self.lib.kmc_timing_reset(self.handle))

    def timing_read(self, family=""):
        ms, cnt = C.c_double(0), C.c_int64(0)
        check(self.lib.kmc_timing_read(self.handle, family.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # ---- Periodic.py
    def put_all_in_box(self, R):
        """in place (Periodic.py:22-25)"""
        if _is_torch(R):
            check(self.lib.kmc_put_all_in_box(self.handle, _ptr(R), R.numel() // 2, KMC_DEVICE))
            return R
        if not (isinstance(R, np.ndarray) and R.dtype == np.float64 and R.flags.c_contiguous):
            raise ValueError("PutAllInBox mutates its argument: need a contiguous float64 numpy array")
        check(self.lib.kmc_put_all_in_box(self.handle, _ptr(R), R.size // 2, KMC_HOST))
        return R

    def distances(self, a, b):
        ka, pa, mem = _in(a)
        kb, pb, _ = _in(b)
        na, nb = _count(ka) // 2, _count(kb) // 2
        out = _Out(ka).new((na, nb))
        check(self.lib.kmc_distances(self.handle, pa, na, pb, nb, _ptr(out), mem))
        return out

    def displacement(self, Ri, Rj):
        ka, pa, mem = _in(Ri)
        kb, pb, _ = _in(Rj)
        k = _count(ka) // 2
        out = _Out(ka).new((2 * k,))
        check(self.lib.kmc_displacement(self.handle, pa, pb, k, _ptr(out), mem))
        return out

    # ---- Bins.py
    def bin_index(self, xy):
        k, p, mem = _in(xy)
        n = _count(k) // 2
        out = _Out(k).new((n, 2), np.int32)
        check(self.lib.kmc_bin_index(self.handle, p, n, _ptr(out), mem))
        return out

    def bins_create(self, R):
        k, p, mem = _in(R)
        n = _count(k) // 2
        h = C.c_void_p()
        check(self.lib.kmc_bins_create(self.handle, p, n, mem if mem is not None else KMC_HOST, C.byref(h)))
        return BinsHandle(self, h, n)

    def bins_export(self, bins):
        """-> cell_of_atom[n], cell_start[ncell+2], sorted_xy[2n], sorted_idx[n] (numpy)"""
        n = bins.n
        ncell2 = self.params.nbins_x * self.params.nbins_y + 2
        cell = np.empty(n, dtype=np.int32)
        start = np.empty(ncell2, dtype=np.int32)
        sxy = np.empty(2 * n)
        sidx = np.empty(n, dtype=np.int32)
        check(self.lib.kmc_bins_export(self.handle, bins.handle, _ptr(cell), _ptr(start), _ptr(sxy), _ptr(sidx), KMC_HOST))
        return cell, start, sxy, sidx

    def nearby_atoms(self, bins, qxy):
        """-> (xy values concatenated, offsets[nq+1] in atoms)  (Bins.py:72-84)"""
        k, p, mem = _in(qxy)
        nq = _count(k) // 2
        o = _Out(k)
        offs = o.new((nq + 1,), np.int64)
        check(self.lib.kmc_nearby_atoms(self.handle, bins.handle, p, nq, None, 0, _ptr(offs), mem))
        total = int(offs[nq])
        out = o.new((2 * total,))
        if total:
            check(self.lib.kmc_nearby_atoms(self.handle, bins.handle, p, nq, _ptr(out), total, _ptr(offs), mem))
        return out, offs

    def neighbors(self, xy, sbins, lists=False):
        """-> n_a (incl. self), n_s [, (a_xy, a_offs), (s_xy, s_offs)]  (Bins.py:86-111)"""
        k, p, mem = _in(xy)
        n = _count(k) // 2
        o = _Out(k)
        na, ns = o.new((n,), np.int32), o.new((n,), np.int32)
        if not lists:
            check(self.lib.kmc_neighbors(self.handle, p, n, sbins.handle, _ptr(na), _ptr(ns), None, 0, None, None, 0, None, mem))
            return na, ns
        ao, so = o.new((n + 1,), np.int64), o.new((n + 1,), np.int64)
        check(self.lib.kmc_neighbors(self.handle, p, n, sbins.handle, _ptr(na), _ptr(ns), None, 0, _ptr(ao), None, 0, _ptr(so), mem))
        ta, ts = int(ao[n]), int(so[n])
        ax, sx = o.new((2 * ta,)), o.new((2 * ts,))
        check(self.lib.kmc_neighbors(self.handle, p, n, sbins.handle, _ptr(na), _ptr(ns), _ptr(ax) if ta else None, ta, _ptr(ao),
                                     _ptr(sx) if ts else None, ts, _ptr(so), mem))
        return na, ns, (ax, ao), (sx, so)

    # ---- Energy.py
    def forces(self, xy, sbins, which=3):
        k, p, mem = _in(xy)
        n = _count(k) // 2
        out = _Out(k).new((2 * n,))
        if n:
            check(self.lib.kmc_forces(self.handle, p, n, sbins.handle, int(which), _ptr(out), mem))
        return out

    def relax_energy(self, xy, sbins, n=None, B=1):
        """xy: B configurations of n adatoms, flattened [B][2n] -> energies[B]  (Energy.py:133-148)"""
        k, p, mem = _in(xy)
        if n is None:
            n = _count(k) // 2 // B
        out = _Out(k).new((B,))
        check(self.lib.kmc_relax_energy(self.handle, p, n, B, sbins.handle, _ptr(out), mem))
        return out

    def line_energies(self, xy, direction, sbins, scale=16, K=20):
        """energies of x + a*s/20/scale, a = 0..K-1  (2DGrowth.py:121-122,152-153)"""
        k, p, mem = _in(xy)
        kd, pd, _ = _in(direction)
        n = _count(k) // 2
        out = _Out(k).new((K,))
        check(self.lib.kmc_line_energies(self.handle, p, pd, n, int(scale), int(K), sbins.handle, _ptr(out), mem))
        return out

    def total_energy(self, xy, sbins):
        k, p, mem = _in(xy)
        out = _Out(k).new((1,))
        check(self.lib.kmc_total_energy(self.handle, p, _count(k) // 2, sbins.handle, _ptr(out), mem))
        return out

    def probe(self, probes, xy, sbins):
        """-> e_aa[m], e_as[m], min_d_a[m], min_d_s[m]  (Energy.py:170-172, 2DGrowth.py:82-90)"""
        kp, pp, mem = _in(probes)
        kx, px, _ = _in(xy)
        m = _count(kp) // 2
        n = _count(kx) // 2 if kx is not None else 0
        o = _Out(kp)
        outs = [o.new((m,)) for _ in range(4)]
        check(self.lib.kmc_probe(self.handle, pp, m, px if n else None, n, sbins.handle, *[_ptr(t) for t in outs], mem))
        return outs

    def rates(self, xy, sbins):
        """-> dict(rate, is_surface, delta_u, u_appx, u_loc, n_a, n_s) for every adatom
        (Energy.py:174-217, 2DGrowth.py:46-62,201-215)"""
        k, p, mem = _in(xy)
        n = _count(k) // 2
        o = _Out(k)
        r = dict(rate=o.new((n,)), is_surface=o.new((n,), np.uint8), delta_u=o.new((n,)), u_appx=o.new((n,)),
                 u_loc=o.new((n,)), n_a=o.new((n,), np.int32), n_s=o.new((n,), np.int32))
        if n:
            check(self.lib.kmc_rates(self.handle, p, n, sbins.handle, _ptr(r["rate"]), _ptr(r["is_surface"]), _ptr(r["delta_u"]),
                                     _ptr(r["u_appx"]), _ptr(r["u_loc"]), _ptr(r["n_a"]), _ptr(r["n_s"]), mem))
        return r

    def select(self, rate_dense, is_surface, u, want_pj=False):
        """-> (k compacted or -1, dense index or -1, Rtot[, pj])  (2DGrowth.py:217-227,322-329)"""
        kr, pr, mem = _in(rate_dense)
        ks, ps, _ = _in(is_surface, np.uint8)
        n = _count(kr)
        o = _Out(kr)
        res, rtot = o.new((2,), np.int64), o.new((1,))
        pj = o.new((n + 1,)) if want_pj else None
        check(self.lib.kmc_select(self.handle, pr, ps, n, float(u), _ptr(res), _ptr(rtot), _ptr(pj), mem))
        if want_pj:
            return int(res[0]), int(res[1]), float(rtot[0]), pj
        return int(res[0]), int(res[1]), float(rtot[0])

    # ---- trajectory state + events (2DGrowth.py:64-166,229-257,317-329)
    def state_set(self, xy):
        k, p, mem = _in(xy)
        n = _count(k) // 2 if k is not None else 0
        check(self.lib.kmc_state_set(self.handle, p, n, mem if mem is not None else KMC_HOST))

    def state_move(self, xy):
        """new positions for the same atoms; keeps the incremental rate table's history (kmc_state_move)"""
        k, p, mem = _in(xy)
        check(self.lib.kmc_state_move(self.handle, p, _count(k) // 2, mem if mem is not None else KMC_HOST))

    def state_get(self):
        n = int(self.lib.kmc_state_count(self.handle))
        out = np.empty(2 * n)
        nn = C.c_int64(0)
        check(self.lib.kmc_state_get(self.handle, _ptr(out), n, C.byref(nn), KMC_HOST))
        return out

    def state_count(self):
        return int(self.lib.kmc_state_count(self.handle))

    def relax(self, sbins, scale=16, threshold=1e-4, max_line_searches=0):
        st = KmcRelaxStats()
        rc = self.lib.kmc_relax(self.handle, sbins.handle, int(scale), float(threshold), int(max_line_searches), C.byref(st))
        if rc != 0 and rc != _lib.KMC_E_ITERCAP:
            check(rc)
        return st

    def set_rate_mode(self, mode, threshold=0.0):
        """0: every event rebuilds the whole rate table (the reference, 2DGrowth.py:318); 1: kmc_event keeps an incremental
        table and recomputes only neighbourhoods in which an atom moved by more than `threshold` Angstrom (0 = exact)"""
        check(self.lib.kmc_set_rate_mode(self.handle, int(mode), float(threshold)))

    def rates_update(self, sbins, threshold=0.0):
        """incremental rate table of the resident state -> dict(rate, is_surface, delta_u, recomputed)"""
        n = self.state_count()
        r = dict(rate=np.zeros(n), is_surface=np.zeros(n, np.uint8), delta_u=np.zeros(n))
        redone = C.c_int64(0)
        check(self.lib.kmc_rates_update(self.handle, sbins.handle, float(threshold), _ptr(r["rate"]), _ptr(r["is_surface"]),
                                        _ptr(r["delta_u"]), C.byref(redone), _lib.KMC_HOST))
        r["recomputed"] = int(redone.value)
        return r

    def set_select_mode(self, mode):
        check(self.lib.kmc_set_select_mode(self.handle, int(mode)))

    def set_relax_mode(self, mode):
        check(self.lib.kmc_set_relax_mode(self.handle, int(mode)))

    def deposit_place(self, sbins, u):
        check(self.lib.kmc_deposit_place(self.handle, sbins.handle, float(u)))

    def hop_place(self, sbins, k, u):
        check(self.lib.kmc_hop_place(self.handle, sbins.handle, int(k), float(u)))

    # ---- spatial domain decomposition (local configuration = own atoms first, then halo copies)
    def domain_sweep(self, xy, n_own, sbins, forces=True, energies=True, rates=True):
        """-> dict(F[2*n_own], e_aa[n_own], e_as[n_own], rate[n_own], is_surface[n_own]) of the own atoms (one cell build)"""
        k, p, mem = _in(xy)
        n = _count(k) // 2
        o = _Out(k)
        out = {"F": o.new((2 * n_own,)) if forces else None,
               "e_aa": o.new((n_own,)) if energies else None, "e_as": o.new((n_own,)) if energies else None,
               "rate": o.new((n_own,)) if rates else None, "is_surface": o.new((n_own,), np.uint8) if rates else None}
        check(self.lib.kmc_domain_sweep(self.handle, p, n, int(n_own), sbins.handle, _ptr(out["F"]), _ptr(out["e_aa"]), _ptr(out["e_as"]),
                                        _ptr(out["rate"]), _ptr(out["is_surface"]), mem))
        return out

    def domain_line_energies(self, xy, direction, n_own, sbins, scale=16, K=20):
        """this rank's share of the K candidate energies (all-reduce over the ranks gives RelaxEnergy of every candidate)"""
        k, p, mem = _in(xy)
        kd, pd, _ = _in(direction)
        n = _count(k) // 2
        out = _Out(k).new((K,))
        check(self.lib.kmc_domain_line_energies(self.handle, p, pd, n, int(n_own), int(scale), int(K), sbins.handle, _ptr(out), mem))
        return out

    def fp64_peak(self, repeats=5):
        """measured non-tensor FP64 FMA rate of the device -> (TFLOP/s, ms of the best run)"""
        tf, ms = C.c_double(0), C.c_double(0)
        check(self.lib.kmc_fp64_peak(self.handle, int(repeats), C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def relax_diag(self):
        """counters of the last relaxation: dict(resorts, movers, work_items), plus rates_recomputed = the number of
        rate-table entries the last kmc_rates_update / thresholded kmc_event recomputed"""
        out = (C.c_int64 * 4)()
        check(self.lib.kmc_relax_diag(self.handle, out))
        return {"resorts": int(out[0]), "movers": int(out[1]), "work_items": int(out[2]), "rates_recomputed": int(out[3])}

    def barrier_bench(self, variant=0, cluster=1, iters=20000):
        """-> (us per device-wide barrier, stale reads seen): variant 0 = k_relax2's barrier, 1 = cluster-hierarchical arrival"""
        us, stale = C.c_double(0), C.c_int32(0)
        check(self.lib.kmc_barrier_bench(self.handle, int(variant), int(cluster), int(iters), C.byref(us), C.byref(stale)))
        return us.value, stale.value

    def cells_stats(self):
        """persistent adatom cell list: dict(full_builds, in_place_updates, last_edits, last_check)"""
        out = (C.c_int64 * 4)()
        check(self.lib.kmc_cells_stats(self.handle, out))
        return {"full_builds": int(out[0]), "in_place_updates": int(out[1]), "last_edits": int(out[2]), "last_check": int(out[3])}

    def cells_check(self):
        """words in which the in-place edited cell list differs from a rebuilt one (0 by construction)"""
        d = C.c_int64(0)
        check(self.lib.kmc_cells_check(self.handle, C.byref(d)))
        return int(d.value)

    def set_event_mode(self, mode):
        """1 (default): a KMC event is queued as one chain of kernels with a single host synchronisation;
        0: first-generation host-stepped placement (same results, cross-check)"""
        check(self.lib.kmc_set_event_mode(self.handle, int(mode)))

    def event_enqueue(self, sbins, u0, u1, max_line_searches=0):
        """queue one KMC event on the context's stream; event_finish() waits for it and returns its outcome"""
        check(self.lib.kmc_event_enqueue(self.handle, sbins.handle, float(u0), float(u1), int(max_line_searches)))

    def event_place(self, sbins, u0, u1):
        """rate table -> selection -> placement, no relaxation -> (kind, hop_k, Rtot)"""
        kind, hk, rtot = C.c_int32(-1), C.c_int64(-1), C.c_double(0)
        check(self.lib.kmc_event_place(self.handle, sbins.handle, float(u0), float(u1), C.byref(kind), C.byref(hk), C.byref(rtot)))
        return kind.value, hk.value, rtot.value

    def event_ready(self):
        """True when the enqueued event has completed on the device (event_finish will not block)"""
        rc = self.lib.kmc_event_ready(self.handle)
        if rc < 0:
            check(rc)
        return rc == 1

    def event_finish(self):
        kind, hk, rtot, st = C.c_int32(-1), C.c_int64(-1), C.c_double(0), KmcRelaxStats()
        rc = self.lib.kmc_event_finish(self.handle, C.byref(kind), C.byref(hk), C.byref(rtot), C.byref(st))
        if rc != 0 and rc != _lib.KMC_E_ITERCAP:
            check(rc)
        return kind.value, hk.value, rtot.value, st

    def event(self, sbins, u0, u1, max_line_searches=0):
        """one KMC event -> (kind 0 deposit / 1 hop, hop_k, Rtot, stats)"""
        kind, hk, rtot, st = C.c_int32(-1), C.c_int64(-1), C.c_double(0), KmcRelaxStats()
        rc = self.lib.kmc_event(self.handle, sbins.handle, float(u0), float(u1), int(max_line_searches), C.byref(kind),
                                C.byref(hk), C.byref(rtot), C.byref(st))
        if rc != 0 and rc != _lib.KMC_E_ITERCAP:
            check(rc)
        return kind.value, hk.value, rtot.value, st


def events_batch(ctxs, sbins, u0, u1, max_line_searches=0):
    """One KMC event of each of R replicas (one Context + substrate cell list each, all on the same device or not):
    every event is queued before the first is waited for (kmc_events_batch).  -> list of (kind, hop_k, Rtot, stats, rc)"""
    R = len(ctxs)
    if R == 0:
        return []
    lib = ctxs[0].lib
    hc = (C.c_void_p * R)(*[c.handle for c in ctxs])
    hb = (C.c_void_p * R)(*[b.handle for b in sbins])
    a0 = (C.c_double * R)(*[float(v) for v in u0])
    a1 = (C.c_double * R)(*[float(v) for v in u1])
    kinds, hks, rtots = (C.c_int32 * R)(), (C.c_int64 * R)(), (C.c_double * R)()
    stats, rcs = (KmcRelaxStats * R)(), (C.c_int32 * R)()
    rc = lib.kmc_events_batch(hc, hb, R, a0, a1, int(max_line_searches), kinds, hks, rtots, stats, rcs)
    if rc != 0:
        check(rc)
    return [(kinds[r], hks[r], rtots[r], stats[r], rcs[r]) for r in range(R)]


def _count(a):
    if a is None:
        return 0
    return int(a.numel()) if _is_torch(a) else int(a.size)


# ------------------------------------------------------------------ per-process default context
_CTX = None


def default_device():
    return int(os.environ.get("KMC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def context():
    """Lazy per-process context bound to the current values of the Constants module."""
    global _CTX
    p = params_from(constants())
    if _CTX is None or _CTX.pid != os.getpid() or _CTX.handle is None:
        _CTX = Context(p, default_device())
    elif _CTX.params.key() != p.key():
        _CTX.set_params(p)
    return _CTX


def reset():
    global _CTX
    if _CTX is not None and _CTX.pid == os.getpid():
        _CTX.close()
    _CTX = None
