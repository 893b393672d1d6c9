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


# ================================================================================================================
# Device-resident decomposition (the product path for BASELINE configs[4]): positions, directions, halos and every
# partial result stay in HBM; the ranks talk through NCCL point-to-point / small collectives on device tensors
# (NVLink), the host only launches.  torch is used for buffers, a few elementwise operations on them and
# torch.distributed -- the pair sweeps are the library's kernels (kmc_domain_sweep, kmc_domain_line_energies).
# ================================================================================================================
DUMMY_Y = 1.0e9          # a halo slot that carries no atom: far outside the cell grid (seen by nobody, evaluated by nobody)


class DeviceDomain(object):
    """One large configuration spread over the ranks of a process group by bin columns.

    Layout of a rank's LOCAL configuration (the argument of the library's domain entry points):
        [ own atoms (n_own) | slots for rank s1's atoms | slots for rank s2's atoms | ... ]
    A neighbour sends ALL its own atoms with those outside the requested boundary columns replaced by a dummy position
    (fixed message size: no compaction, no count exchange, no host synchronisation; a message is n_own x 16 or 32 bytes,
    i.e. KBs -- the exchange is latency bound, not bandwidth bound, so the padding is free).  Ownership is fixed by the
    columns at partition time (once per event); halo membership follows the current columns at every exchange.
    The substrate is static and replicated (268 MB for the 4096x4096 substrate of configs[4]: 0.15 % of one B200's HBM)."""

    def __init__(self, gv, substrate, adatoms, rank=0, world=1, device=0, dist=None, ctx=None, sbins=None):
        import torch
        from . import runtime
        if int(getattr(gv, "aa_mode", 0)) != 1:
            raise ValueError("domain decomposition needs aa_mode == 1 (bins3x3 adatom sums); got aa_mode=%r" % getattr(gv, "aa_mode", 0))
        self.torch, self.gv, self.rank, self.world, self.dist = torch, gv, rank, world, dist
        self.dev = torch.device("cuda", device)
        self.dec = Decomposition(gv, world)
        self.ctx = ctx if ctx is not None else runtime.Context(runtime.params_from(gv), device)
        self.own_ctx = ctx is None
        with torch.cuda.device(self.dev):
            self.ctx.use_torch_stream()          # the library's kernels and torch's elementwise / NCCL work share one stream order
        self.sb = sbins if sbins is not None else self.ctx.bins_create(np.asarray(substrate, dtype=np.float64))
        self.own_sb = sbins is None
        self.srcs = [s for s in range(world) if s != rank and self.dec.sends[s][rank]]
        self.dsts = [d for d in range(world) if d != rank and self.dec.sends[rank][d]]
        self.cols_for = {d: torch.tensor(self.dec.sends[rank][d], dtype=torch.int64, device=self.dev) for d in self.dsts}
        self.halo_bytes = 0
        self.collectives = 0
        self.partition(adatoms)

    # ---- ownership -------------------------------------------------------------------------------------------
    def partition(self, adatoms):
        """(re)assign every adatom to the rank that owns its bin column; every rank computes the same assignment"""
        torch = self.torch
        ad = np.asarray(adatoms, dtype=np.float64).reshape(-1, 2)
        self.n_global = ad.shape[0]
        owner = self.dec.owner_of(ad[:, 0]) if ad.shape[0] else np.zeros(0, dtype=np.int64)
        self.gids = [np.nonzero(owner == r)[0].astype(np.int64) for r in range(self.world)]
        self.n_own = int(len(self.gids[self.rank]))
        self.gid = torch.from_numpy(self.gids[self.rank]).to(self.dev)
        self.x = torch.from_numpy(np.ascontiguousarray(ad[self.gids[self.rank]])).to(self.dev)
        off = self.n_own
        self.recv_slice = {}
        for s in self.srcs:
            self.recv_slice[s] = (off, off + len(self.gids[s]))
            off += len(self.gids[s])
        self.n_local = off
        self.loc4 = torch.empty((self.n_local, 4), dtype=torch.float64, device=self.dev)        # x, y, sx, sy
        self.loc4[:, 0] = 0.0
        self.loc4[:, 1] = DUMMY_Y
        self.loc4[:, 2:] = 0.0

    # ---- halo exchange ---------------------------------------------------------------------------------------
    def exchange(self, x, s=None):
        """-> (local positions [n_local, 2], local directions [n_local, 2] or None), contiguous device tensors.
        Own atoms first; the neighbours' boundary columns behind them (dummies elsewhere)."""
        torch, dist = self.torch, self.dist
        w = 4 if s is not None else 2
        own = torch.cat([x, s], dim=1) if s is not None else x
        self.loc4[:self.n_own, :w] = own
        if self.world > 1 and (self.srcs or self.dsts):
            col = ((x[:, 0] + self.gv.L / 2) / self.gv.bin_size).to(torch.int64).clamp_(0, self.gv.nbins_x - 1)
            dummy = torch.tensor([0.0, DUMMY_Y, 0.0, 0.0][:w], dtype=torch.float64, device=self.dev)
            ops, keep = [], []
            for d in self.dsts:
                cols = self.dec.sends[self.rank][d]              # one to three columns: plain comparisons (torch.isin sorts)
                mask = col == cols[0]
                for cnum in cols[1:]:
                    mask = mask | (col == cnum)
                payload = torch.where(mask[:, None], own, dummy).contiguous()
                keep.append(payload)
                ops.append(dist.P2POp(dist.isend, payload, d))
                self.halo_bytes += payload.numel() * 8
            bufs = {}
            for src in self.srcs:
                a, b = self.recv_slice[src]
                bufs[src] = torch.empty((b - a, w), dtype=torch.float64, device=self.dev)
                ops.append(dist.P2POp(dist.irecv, bufs[src], src))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                self.collectives += 1
            for src in self.srcs:
                a, b = self.recv_slice[src]
                self.loc4[a:b, :w] = bufs[src]
        lx = self.loc4[:, 0:2].contiguous()
        ls = self.loc4[:, 2:4].contiguous() if s is not None else None
        return lx, ls

    def _allreduce(self, t, op="sum"):
        if self.world > 1:
            d = self.dist
            d.all_reduce(t, op={"sum": d.ReduceOp.SUM, "max": d.ReduceOp.MAX, "min": d.ReduceOp.MIN}[op])
            self.collectives += 1
        return t

    # ---- sweeps ----------------------------------------------------------------------------------------------
    def sweep(self, x=None, forces=True, energies=True, rates=True):
        """one decomposed sweep of the configuration: dict(F [n_own, 2], energy (0-dim, all ranks), rate_dense, surf_dense)"""
        torch = self.torch
        x = self.x if x is None else x
        lx, _ = self.exchange(x)
        r = self.ctx.domain_sweep(lx.view(-1), self.n_own, self.sb, forces=forces, energies=energies, rates=rates)
        out = {}
        if forces:
            out["F"] = r["F"].view(-1, 2)
        if energies:
            out["energy"] = self._allreduce((0.5 * r["e_aa"].sum() + r["e_as"].sum()).reshape(1))[0]
        if rates:
            dense = torch.zeros(self.n_global, dtype=torch.float64, device=self.dev)
            dsurf = torch.zeros(self.n_global, dtype=torch.float64, device=self.dev)
            dense.index_copy_(0, self.gid, r["rate"])
            dsurf.index_copy_(0, self.gid, r["is_surface"].to(torch.float64))
            both = torch.stack([dense, dsurf])
            self._allreduce(both)            # slices are disjoint and zero elsewhere: the sum is the assembly, exactly
            out["rate_dense"], out["surf_dense"] = both[0].contiguous(), both[1].to(torch.uint8).contiguous()
        return out

    def select(self, u, sw=None):
        """2DGrowth.py:319-329 on the assembled table: the same sequential-order kernel as on one GPU (bit-exact index)"""
        sw = self.sweep(forces=False, energies=False, rates=True) if sw is None else sw
        return self.ctx.select(sw["rate_dense"], sw["surf_dense"], u)

    def gather_positions(self, x=None):
        """global positions [n_global, 2] on every rank (device)"""
        torch = self.torch
        x = self.x if x is None else x
        g = torch.zeros((self.n_global, 2), dtype=torch.float64, device=self.dev)
        g.index_copy_(0, self.gid, x)
        return self._allreduce(g)

    # ---- Relaxation (2DGrowth.py:97-166) over the ranks ------------------------------------------------------------
    def _line_search(self, xn, sn, scale):
        """one 20-candidate line search + forces at the winner; -> (xnp, F, maxF, top, bot, miny) with host floats"""
        torch = self.torch
        lx, ls = self.exchange(xn, sn)                                              # halo positions AND directions
        e = self.ctx.domain_line_energies(lx.view(-1), ls.view(-1), self.n_own, self.sb, scale, 20)
        self._allreduce(e)                                                          # RelaxEnergy of the 20 candidates (:122,153)
        amin = torch.min(e, dim=0).indices.to(torch.float64)                        # first minimum (:123,155)
        lxnp = lx + (amin * ls) / 20 / scale                                        # x + a*s/20/scale, own and halo slots (:121,152)
        r = self.ctx.domain_sweep(lxnp.contiguous().view(-1), self.n_own, self.sb, forces=True, energies=False, rates=False)
        F = r["F"].view(-1, 2)
        xnp = lxnp[:self.n_own]
        big = torch.finfo(torch.float64).max
        if self.n_own:
            part = torch.stack([(F * F).sum(dim=1).max(), (xnp * (xnp - xn)).sum(), (xn * xn).sum(), -xnp[:, 1].min()])
        else:
            part = torch.tensor([0.0, 0.0, 0.0, -big], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            allp = [torch.empty_like(part) for _ in range(self.world)]
            self.dist.all_gather(allp, part)
            self.collectives += 1
            allp = torch.stack(allp)
        else:
            allp = part[None]
        red = torch.stack([allp[:, 0].max(), allp[:, 1].sum(), allp[:, 2].sum(), -allp[:, 3].max()]).cpu()       # the ONE host sync of a line search
        return xnp, F, float(red[0]), float(red[1]), float(red[2]), float(red[3])

    def relax(self, scale0=16, threshold0=1e-4, max_line_searches=0):
        """the reference's Relaxation with its restart rules, host-driven, every sum reduced over the ranks -> dict(stats)"""
        x0 = self.x.clone()
        scale, threshold = int(scale0), float(threshold0)
        ls_count = restarts = 0
        status = 0
        maxF = 0.0
        xnp = x0
        while True:
            if scale > 1024:                                                        # :111-113
                scale, threshold = 16, threshold * 10
            xn = x0
            sn = self.sweep(xn, forces=True, energies=False, rates=False)["F"].clone()      # dx0, :120,125
            xnp, F, maxF, top, bot, miny = self._line_search(xn, sn, scale)         # :121-129
            ls_count += 1
            lastmaxF = maxF
            restart = False
            while maxF > threshold:                                                 # :131
                if max_line_searches > 0 and ls_count >= max_line_searches:
                    status = -6
                    break
                if miny > self.gv.Dmax:                                             # :137-140
                    restart = True
                    break
                beta = top / bot                                                    # :143-145
                sn = F + beta * sn                                                  # :147
                xn = xnp                                                            # :149
                xnp, F, maxF, top, bot, miny = self._line_search(xn, sn, scale)     # :152-160
                ls_count += 1
                if abs(lastmaxF - maxF) < 1e-8:                                     # :162-164
                    restart = True
                    break
                lastmaxF = maxF
            if status != 0 or not restart:
                break
            restarts += 1
            scale *= 2
        self.x = xnp.contiguous().clone()
        self.ctx.put_all_in_box(self.x.view(-1))                                    # :166
        return {"line_searches": ls_count, "restarts": restarts, "maxF": maxF, "status": status, "final_scale": scale,
                "final_threshold": threshold}

    # ---- one KMC event (2DGrowth.py:317-329) ----------------------------------------------------------------------
    def event(self, u0, u1, max_line_searches=0):
        """rate table -> selection -> placement on the full state (every rank, identically: the kernels of the single-GPU
        event chain), then the decomposed relaxation."""
        g = self.gather_positions().contiguous()
        self.ctx.state_set(g.view(-1))
        kind, hk, rtot = self.ctx.event_place(self.sb, u0, u1)
        self.partition(self.ctx.state_get())
        st = self.relax(16, 1e-4, max_line_searches)
        st.update(kind=kind, hop_k=hk, rtot=rtot)
        return st

    def close(self):
        if self.own_sb:
            self.sb.close()
        if self.own_ctx:
            self.ctx.close()
