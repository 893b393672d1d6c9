"""Replica-parallel temperature x flux sweeps (BASELINE.json configs[3]).

The only way this path shards without a data-path collective: independent trajectories.  Replica r
runs on rank r mod world_size (one process per GPU); every replica carries its own constants
(beta from its temperature, Rd from its flux -- the reference hard-codes both, Constants.py:45-46,
2DGrowth.py:43) and its own uniform stream (seed 1000 + r).  There is no communication during the
run; the per-replica summaries are gathered once at the end (all_gather_object on whatever backend
the process group uses: NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

from . import workloads


def sweep_grid(n_temperatures=8, n_fluxes=8, t_lo=250.0, t_hi=600.0, f_lo=-2.0, f_hi=1.0):
    """The 8 x 8 grid of SURVEY.md §8d item 4: T in linspace(250, 600), flux in logspace(-2, 1)."""
    Ts = np.linspace(t_lo, t_hi, n_temperatures)
    Fs = np.logspace(f_lo, f_hi, n_fluxes)
    return [(float(T), float(F)) for T in Ts for F in Fs]


def assign(n_replicas, world_size, rank):
    """replica ids owned by `rank` (round robin: replica r -> rank r mod world_size)"""
    return [r for r in range(n_replicas) if r % world_size == rank]


def replica_constants(r, grid, W=18, Hs=10, Ha=10, aa_mode=0):
    T, flux = grid[r]
    return workloads.constants_for(W, Hs, Ha, temperature=T, deposition_rate=flux, aa_mode=aa_mode,
                                   hop_index_mode=1, deposit_mode=1)


def run_replica_gpu(r, grid, n_events, device, max_line_searches=0, W=18, Hs=10, Ha=10):
    """One trajectory of `n_events` KMC events on `device`; returns a small summary dict."""
    from . import Growth, runtime
    gv = replica_constants(r, grid, W, Hs, Ha)
    sim = Growth.Simulation(adatoms=None, substrate=workloads.substrate(gv), device=device, params=runtime.params_from(gv))
    rng = np.random.default_rng(1000 + r)
    us = rng.uniform(0, 1, (n_events, 2))
    hops = deposits = ls = 0
    for t in range(n_events):
        rec = sim.step(us[t, 0], us[t, 1], max_line_searches)
        hops += rec["kind"] == 1
        deposits += rec["kind"] == 0
        ls += rec["line_searches"]
    ad = sim.adatoms
    sim.close()
    return {"replica": r, "temperature": grid[r][0], "flux": grid[r][1], "events": n_events, "hops": int(hops),
            "deposits": int(deposits), "line_searches": int(ls), "n_adatoms": int(ad.size // 2),
            "mean_height": float(ad[1::2].mean()) if ad.size else 0.0}


def run_sweep(grid, n_events, runner=run_replica_gpu, world_size=1, rank=0, device=0, gather=None, threads=1, **kw):
    """Run this rank's share of the sweep; `gather(local_list) -> list of all ranks' lists` joins the
    summaries (None: single process).  Returns the summaries of ALL replicas, ordered by replica id.

    threads > 1 runs that many replicas of this rank CONCURRENTLY on its GPU: every replica owns a context and
    a stream, its relaxations are single-block or single-cluster launches (<= 8 SMs), so several fit on the 148
    SMs at once; the ctypes calls release the GIL while they wait."""
    ids = assign(len(grid), world_size, rank)
    if threads > 1 and len(ids) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as ex:
            mine = list(ex.map(lambda r: runner(r, grid, n_events, device, **kw), ids))
    else:
        mine = [runner(r, grid, n_events, device, **kw) for r in ids]
    if gather is None:
        allr = [mine]
    else:
        allr = gather(mine)
    flat = sorted((s for part in allr for s in part), key=lambda s: s["replica"])
    return flat


def dist_gather(local):
    """all_gather_object over the default process group"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out
