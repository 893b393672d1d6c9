"""1-D spatial domain decomposition of the force / energy / rate sweeps for very large substrates
(BASELINE.json configs[4]): each rank owns a contiguous range of bin COLUMNS (x is the
only periodic axis, Bins.py:54-70) and exchanges the adatoms of its boundary columns with the ranks whose
stencils reach them -- one column on each side, plus the seam pair (column 0 <-> column nbins_x-2,
Bins.py:66-69).  The cell grid is the GLOBAL one (same L, bin_size, nbins), so every owned atom sees
exactly the neighbours it sees on one GPU; only the summation order inside a cell can differ.

Collectives (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests):
  * halo exchange: counts by all_gather, payload by batch_isend_irecv (KB-sized, latency bound);
  * energy: one all_reduce of a scalar;
  * event selection: all_gather of the dense rate slices (8 B per adatom), then the SAME sequential-order
    selection kernel as on one GPU, so the selected index stays bit-exact.
The substrate is static: every rank bins the substrate atoms of its own columns + halo once.

The KMC *sequence* of events is serial; whole trajectories scale through
replicas (replicas.py), this module scales the sweeps of ONE large configuration.
"""
import numpy as np


def column_of(x, gv):
    """nx of Bins.BinIndex (Bins.py:20) vectorised: IEEE add, divide, truncation toward zero."""
    return ((np.asarray(x, dtype=np.float64) + gv.L / 2) / gv.bin_size).astype(np.int64)


def stencil_columns(cx, nbx):
    """columns an atom in column cx reads (Bins.py:54-69), duplicates irrelevant here"""
    cols = set()
    for v in (cx - 1, cx, cx + 1):
        if v < 0:
            v += nbx
        if v >= nbx:
            v -= nbx
        cols.add(v)
    if cx == 0:
        cols.add(nbx - 2)
    if cx == nbx - 2:
        cols.add(0)
    return cols


class Decomposition(object):
    """column ranges [c0, c1) per rank and the halo relations between ranks"""

    def __init__(self, gv, world):
        self.gv, self.world = gv, world
        nbx = gv.nbins_x
        self.bounds = [(r * nbx // world, (r + 1) * nbx // world) for r in range(world)]
        self.owner = np.empty(nbx, dtype=np.int64)
        for r, (a, b) in enumerate(self.bounds):
            self.owner[a:b] = r
        # needs[r] = columns rank r must see but does not own
        self.needs = []
        for r, (a, b) in enumerate(self.bounds):
            want = set()
            for cx in range(a, b):
                want |= stencil_columns(cx, nbx)
            self.needs.append(sorted(c for c in want if not (a <= c < b)))
        # sends[r][d] = columns of rank r that rank d needs
        self.sends = [[[c for c in self.needs[d] if self.owner[c] == r] for d in range(world)] for r in range(world)]

    def owner_of(self, x):
        """owning rank of atoms at x; atoms outside the addressable columns go to the nearest end rank
        (they see others through their wrapped stencil but are seen by nobody)"""
        cx = column_of(x, self.gv)
        return self.owner[np.clip(cx, 0, self.gv.nbins_x - 1)]


class OracleBackend(object):
    """local sweeps on the CPU oracle -- TEST infrastructure only (tests/test_domain_gloo.py)"""

    def __init__(self, oracle, substrate):
        self.o = oracle
        self.sb = oracle.PutInBins(substrate)

    def forces(self, xy):
        return self.o.AdatomAdatomForces(xy) + self.o.AdatomSubstrateForces(xy, self.sb)

    def site_energies(self, xy, own):
        # per owned atom: e_aa (stencil / all-pairs sum incl. the zero self term) and e_as
        ua, ul, ui, du, na, ns = self.o.Barriers(xy, self.sb)
        eas = np.array([self.o.AdatomSurfaceEnergy(xy[2 * i:2 * i + 2], self.sb) for i in range(xy.size // 2)])
        return (ua - eas)[own], eas[own]

    def rates(self, xy):
        rate, surf, du = self.o.HoppingRatesDense(xy, self.sb)
        return rate, surf.astype(np.uint8)

    def select(self, rate, surf, u):
        return self.o.Select(rate, u)


class GpuBackend(object):
    """local sweeps on this rank's B200 through libkmcb200"""

    def __init__(self, gv, substrate, device):
        from . import runtime
        self.ctx = runtime.Context(runtime.params_from(gv), device)
        self.sb = self.ctx.bins_create(substrate)

    def forces(self, xy):
        return self.ctx.forces(xy, self.sb, 3)

    def site_energies(self, xy, own):
        probes = np.ascontiguousarray(xy.reshape(-1, 2)[own]).reshape(-1)
        e_aa, e_as, _, _ = self.ctx.probe(probes, xy, self.sb)
        return e_aa, e_as

    def rates(self, xy):
        r = self.ctx.rates(xy, self.sb)
        return r["rate"], r["is_surface"]

    def select(self, rate, surf, u):
        return self.ctx.select(rate, surf, u)

    def close(self):
        self.sb.close()
        self.ctx.close()


class DomainSweeps(object):
    """Force / energy / rate sweeps of one configuration spread over the ranks of a process group."""

    def __init__(self, gv, substrate, adatoms, rank, world, backend_factory, dist=None, device=None):
        if int(getattr(gv, "aa_mode", 0)) != 1:
            # all-pairs adatom sums (aa_mode 0, Energy.py:62-67) need every adatom on every rank: a one-column halo would
            # silently drop pairs.  The decomposition is defined for the stencil mode only.
            raise ValueError("domain decomposition needs aa_mode == 1 (bins3x3 adatom sums); got aa_mode=%r" % getattr(gv, "aa_mode", 0))
        self.gv, self.rank, self.world, self.dist, self.device = gv, rank, world, dist, device
        self.dec = Decomposition(gv, world)
        # static substrate: own columns + halo columns
        sx = substrate[0::2]
        scol = np.clip(column_of(sx, gv), 0, gv.nbins_x - 1)
        a, b = self.dec.bounds[rank]
        keep = ((scol >= a) & (scol < b)) | np.isin(scol, self.dec.needs[rank])
        # atoms the reference stores outside the grid (x == L/2 with integral L/bin_size) are invisible anyway
        self.backend = backend_factory(np.ascontiguousarray(substrate.reshape(-1, 2)[keep]).reshape(-1))
        ad = np.asarray(adatoms, dtype=np.float64).reshape(-1, 2)
        self.n_global = ad.shape[0]
        mine = self.dec.owner_of(ad[:, 0]) == rank
        self.gid = np.nonzero(mine)[0].astype(np.int64)          # global indices of the owned adatoms (ascending)
        self.xy = ad[mine].copy()
        self.halo_bytes = 0

    # ---- halo exchange ---------------------------------------------------------------------
    def _exchange(self):
        """-> (local_xy [m,2], local_gid [m], owned_mask [m]) sorted by global index"""
        if self.world == 1:
            return self.xy, self.gid, np.ones(len(self.gid), dtype=bool)
        import torch
        dist = self.dist
        col = np.clip(column_of(self.xy[:, 0], self.gv), 0, self.gv.nbins_x - 1)
        out = []
        for d in range(self.world):
            cols = self.dec.sends[self.rank][d]
            sel = np.isin(col, cols) if (cols and d != self.rank) else np.zeros(len(col), dtype=bool)
            out.append(np.column_stack([self.gid[sel].astype(np.float64), self.xy[sel]]) if sel.any() else np.zeros((0, 3)))
        dev = self.device if self.device is not None else "cpu"
        counts = torch.tensor([len(o) for o in out], dtype=torch.int64, device=dev)
        allc = [torch.zeros_like(counts) for _ in range(self.world)]
        dist.all_gather(allc, counts)
        recv_counts = [int(allc[s][self.rank]) for s in range(self.world)]
        ops, recv_bufs, keep = [], {}, []
        for d in range(self.world):
            if d == self.rank:
                continue
            if len(out[d]):
                t = torch.from_numpy(np.ascontiguousarray(out[d])).to(dev)
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, d))
                self.halo_bytes += t.numel() * 8
            if recv_counts[d]:
                recv_bufs[d] = torch.empty((recv_counts[d], 3), dtype=torch.float64, device=dev)
                ops.append(dist.P2POp(dist.irecv, recv_bufs[d], d))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        halos = [recv_bufs[d].cpu().numpy() for d in sorted(recv_bufs)]
        if halos:
            h = np.concatenate(halos)
            # a seam column can be requested twice (as neighbour and as seam partner): keep one copy
            hg, first = np.unique(h[:, 0].astype(np.int64), return_index=True)
            h = h[first]
            gid = np.concatenate([self.gid, hg])
            xy = np.concatenate([self.xy, h[:, 1:3]])
            own = np.concatenate([np.ones(len(self.gid), dtype=bool), np.zeros(len(hg), dtype=bool)])
        else:
            gid, xy, own = self.gid, self.xy, np.ones(len(self.gid), dtype=bool)
        order = np.argsort(gid, kind="stable")
        return xy[order], gid[order], own[order]

    def _allreduce_sum(self, v):
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t)
        return float(t[0])

    def _gather_dense(self, local_vals, dtype):
        """dense [n_global] array assembled from every rank's owned slice"""
        full = np.zeros(self.n_global, dtype=dtype)
        full[self.gid] = local_vals
        if self.world == 1:
            return full
        import torch
        dev = self.device if self.device is not None else "cpu"
        t = torch.from_numpy(full.astype(np.float64)).to(dev)
        self.dist.all_reduce(t)          # slices are disjoint and zero elsewhere: the sum is the assembly, exactly
        return t.cpu().numpy().astype(dtype)

    # ---- sweeps ------------------------------------------------------------------------------
    def forces(self):
        """Energy.AdatomAdatomForces + AdatomSubstrateForces for the owned adatoms -> [n_owned, 2]"""
        xy, gid, own = self._exchange()
        F = np.asarray(self.backend.forces(xy.reshape(-1))).reshape(-1, 2)
        return F[own]

    def energy(self):
        """Energy.RelaxEnergy of the whole configuration (every rank gets the same value)"""
        xy, gid, own = self._exchange()
        e_aa, e_as = self.backend.site_energies(xy.reshape(-1), own)
        local = float(0.5 * np.sum(np.asarray(e_aa)) + np.sum(np.asarray(e_as)))
        return self._allreduce_sum(local)

    def rates(self):
        """dense rate table and surface flags of ALL adatoms, on every rank (2DGrowth.py:201-215)"""
        xy, gid, own = self._exchange()
        rate, surf = self.backend.rates(xy.reshape(-1))
        rate, surf = np.asarray(rate)[own], np.asarray(surf)[own]
        return self._gather_dense(rate, np.float64), self._gather_dense(surf, np.float64).astype(np.uint8)

    def select(self, u):
        """2DGrowth.py:319-329 on the gathered table: same kernel, same order, same index as on one GPU"""
        rate, surf = self.rates()
        return self.backend.select(rate, surf, u)

    def gather_owned(self, local_rows):
        """assemble per-owned-atom rows ([n_owned, k]) into the global order on every rank"""
        local_rows = np.asarray(local_rows, dtype=np.float64)
        cols = [self._gather_dense(local_rows[:, k], np.float64) for k in range(local_rows.shape[1])]
        return np.column_stack(cols)
