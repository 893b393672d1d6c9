"""Replica-parallel temperature x flux sweeps (BASELINE.json configs[3]).

The only way this path shards without a data-path collective: independent trajectories.  Every replica carries its own
constants (beta from its temperature, Rd from its flux -- the reference hard-codes both, Constants.py:45-46,
2DGrowth.py:43) and its own uniform stream (seed 1000 + r).  There is no communication during the run.

One GPU advances MANY replicas at once (`ReplicaBatch`): every replica owns a context and a stream, a KMC event is one
chain of kernels queued without a host round trip (kmc_event_enqueue), and a replica's next event is queued the moment its
previous one has completed (kmc_event_ready) -- no lockstep between replicas, no host thread per replica.  The relaxations
of small systems are single-block / single-cluster launches (<= 8 SMs), so a dozen and more overlap on the 148 SMs.

Across GPUs (one process per GPU) replicas are handed out DYNAMICALLY: a shared counter (the process group's store)
gives the next replica id to whichever rank has a free slot, so a rank stuck in one long relaxation (a single event can
take 4e4 line searches) does not hold other replicas hostage; `assign` (static round robin) is kept for reference.
The per-replica summaries are gathered once at the end (all_gather_object on whatever backend the group uses:
NCCL on the GPU box, gloo in the CPU tests).
"""
import time

import numpy as np

from . import workloads


def sweep_grid(n_temperatures=8, n_fluxes=8, t_lo=250.0, t_hi=600.0, f_lo=-2.0, f_hi=1.0):
    """The 8 x 8 grid of SURVEY.md §8d item 4: T in linspace(250, 600), flux in logspace(-2, 1)."""
    Ts = np.linspace(t_lo, t_hi, n_temperatures)
    Fs = np.logspace(f_lo, f_hi, n_fluxes)
    return [(float(T), float(F)) for T in Ts for F in Fs]


def assign(n_replicas, world_size, rank):
    """static assignment (replica r -> rank r mod world_size); the sweeps use the dynamic queue below"""
    return [r for r in range(n_replicas) if r % world_size == rank]


def replica_constants(r, grid, W=18, Hs=10, Ha=10, aa_mode=0):
    T, flux = grid[r]
    return workloads.constants_for(W, Hs, Ha, temperature=T, deposition_rate=flux, aa_mode=aa_mode,
                                   hop_index_mode=1, deposit_mode=1)


def replica_uniforms(r, n_events):
    return np.random.default_rng(1000 + r).uniform(0, 1, (n_events, 2))


def summarize(r, grid, n_events, log, ad):
    return {"replica": r, "temperature": grid[r][0], "flux": grid[r][1], "events": n_events,
            "hops": int(sum(1 for k in log if k[0] == 1)), "deposits": int(sum(1 for k in log if k[0] == 0)),
            "line_searches": int(sum(k[2] for k in log)), "n_adatoms": int(ad.size // 2),
            "mean_height": float(ad[1::2].mean()) if ad.size else 0.0}


def run_replica_gpu(r, grid, n_events, device, max_line_searches=0, W=18, Hs=10, Ha=10):
    """One trajectory of `n_events` KMC events on `device`, alone; returns a small summary dict."""
    from . import Growth, runtime
    gv = replica_constants(r, grid, W, Hs, Ha)
    sim = Growth.Simulation(adatoms=None, substrate=workloads.substrate(gv), device=device, params=runtime.params_from(gv))
    us = replica_uniforms(r, n_events)
    log = []
    for t in range(n_events):
        rec = sim.step(us[t, 0], us[t, 1], max_line_searches)
        log.append((rec["kind"], rec["hop_k"], rec["line_searches"]))
    ad = sim.adatoms
    sim.close()
    return summarize(r, grid, n_events, log, ad)


class LocalQueue(object):
    """next replica id for a single process (thread safe: several engine threads may share it)"""

    def __init__(self, ids):
        import threading
        self.ids, self.pos, self.lock = list(ids), 0, threading.Lock()

    def __call__(self):
        with self.lock:
            if self.pos >= len(self.ids):
                return None
            self.pos += 1
            return self.ids[self.pos - 1]


class StoreQueue(object):
    """next replica id across the ranks of a torch.distributed process group: an atomic counter in the group's store
    (no data-path collective; works on every backend)"""

    def __init__(self, n_replicas, key="kmc_replica_queue", order=None):
        import torch.distributed as dist
        self.store = dist.distributed_c10d._get_default_store()
        self.n, self.key = n_replicas, key
        self.order = list(order) if order is not None else list(range(n_replicas))

    def __call__(self):
        k = int(self.store.add(self.key, 1)) - 1
        return self.order[k] if k < self.n else None


class _Slot(object):
    __slots__ = ("r", "ctx", "us", "t", "log", "n_events")


class ReplicaBatch(object):
    """Up to `width` replicas in flight on one GPU, fed from a queue of replica ids."""

    def __init__(self, ids, grid, device=0, width=16, W=18, Hs=10, Ha=10, max_line_searches=0, queue=None, runtime=None):
        if runtime is None:         # (tests of the host logic inject a stand-in for the CUDA binding)
            from . import runtime
        self.grid, self.device, self.width = grid, device, max(1, int(width))
        self.geom = (W, Hs, Ha)
        self.max_ls = int(max_line_searches)
        self.queue = queue if queue is not None else LocalQueue(ids)
        gv0 = replica_constants(0, grid, W, Hs, Ha)
        self._runtime = runtime
        self.pool = []              # idle contexts (same cell grid for every replica: only beta and Rd differ)
        self.owner = runtime.Context(runtime.params_from(gv0), device)
        self.sbins = self.owner.bins_create(workloads.substrate(gv0))        # static substrate, shared by every replica on this device

    def _ctx_for(self, r):
        p = self._runtime.params_from(replica_constants(r, self.grid, *self.geom))
        if self.pool:
            ctx = self.pool.pop()
            ctx.set_params(p)
        else:
            ctx = self._runtime.Context(p, self.device)
        ctx.state_set(np.zeros(0))
        return ctx

    def _start(self, r, n_events):
        s = _Slot()
        s.r, s.ctx, s.us, s.t, s.log, s.n_events = r, self._ctx_for(r), replica_uniforms(r, n_events), 0, [], n_events
        s.ctx.event_enqueue(self.sbins, s.us[0, 0], s.us[0, 1], self.max_ls)
        return s

    def run(self, n_events):
        """-> dict(log = {r: [(kind, hop_k, line_searches), ...]}, adatoms = {r: final positions}, summaries = [...])"""
        out = {"log": {}, "adatoms": {}, "summaries": []}
        active = []

        def refill():
            while len(active) < self.width:
                r = self.queue()
                if r is None:
                    return
                active.append(self._start(r, n_events))

        prof = out["host_seconds"] = {"start": 0.0, "poll": 0.0, "finish": 0.0, "enqueue": 0.0, "idle_sweeps": 0}
        t0 = time.perf_counter()
        refill()
        prof["start"] += time.perf_counter() - t0
        while active:
            progressed = False
            for s in list(active):
                t0 = time.perf_counter()
                ready = s.ctx.event_ready()
                t1 = time.perf_counter()
                prof["poll"] += t1 - t0
                if not ready:
                    continue
                progressed = True
                kind, hk, rtot, st = s.ctx.event_finish()
                t2 = time.perf_counter()
                prof["finish"] += t2 - t1
                s.log.append((kind, hk, int(st.line_searches)))
                s.t += 1
                if s.t < s.n_events:
                    s.ctx.event_enqueue(self.sbins, s.us[s.t, 0], s.us[s.t, 1], self.max_ls)
                    prof["enqueue"] += time.perf_counter() - t2
                else:
                    ad = s.ctx.state_get()
                    out["log"][s.r], out["adatoms"][s.r] = s.log, ad
                    out["summaries"].append(summarize(s.r, self.grid, n_events, s.log, ad))
                    active.remove(s)
                    self.pool.append(s.ctx)
                    refill()
            if not progressed:
                prof["idle_sweeps"] += 1
                time.sleep(0)           # yield the GIL / CPU while every replica is busy on the device
        return out

    def close(self):
        for c in self.pool:
            c.close()
        self.pool = []
        self.sbins.close()
        self.owner.close()


def run_sweep(grid, n_events, runner=run_replica_gpu, world_size=1, rank=0, device=0, gather=None, threads=1, **kw):
    """Static round-robin sweep with a pluggable per-replica runner (kept for the CPU tests of the host logic and as the
    baseline the batched engine is compared with).  Returns the summaries of ALL replicas, ordered by replica id."""
    ids = assign(len(grid), world_size, rank)
    if threads > 1 and len(ids) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            mine = list(ex.map(lambda r: runner(r, grid, n_events, device, **kw), ids))
    else:
        mine = [runner(r, grid, n_events, device, **kw) for r in ids]
    allr = [mine] if gather is None else gather(mine)
    return sorted((s for part in allr for s in part), key=lambda s: s["replica"])


def run_sweep_batched(grid, n_events, device=0, width=16, queue=None, gather=None, threads=1, **kw):
    """The sweep on the batched engine: this process keeps `width` replicas in flight on `device` PER ENGINE THREAD, taking
    replica ids from `queue` (LocalQueue over all replicas by default, StoreQueue for a multi-process run).  `threads` > 1
    runs that many engines concurrently (the ctypes calls release the GIL): a host thread that the driver holds up while it
    submits work does not stall the other engines' replicas.  -> summaries of all replicas."""
    if queue is None:
        queue = LocalQueue(range(len(grid)))
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor

        def one(_):
            b = ReplicaBatch((), grid, device=device, width=width, queue=queue, **kw)
            try:
                return b.run(n_events)
            finally:
                b.close()

        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(one, range(threads)))
        mine = [s for p in parts for s in p["summaries"]]
        run_sweep_batched.last_host_seconds = parts[0]["host_seconds"]
        allr = [mine] if gather is None else gather(mine)
        return sorted((s for part in allr for s in part), key=lambda s: s["replica"])
    b = ReplicaBatch(range(len(grid)), grid, device=device, width=width, queue=queue, **kw)
    try:
        res = b.run(n_events)
        mine = res["summaries"]
        run_sweep_batched.last_host_seconds = res["host_seconds"]
    finally:
        b.close()
    allr = [mine] if gather is None else gather(mine)
    return sorted((s for part in allr for s in part), key=lambda s: s["replica"])


def dist_gather(local):
    """all_gather_object over the default process group"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out
