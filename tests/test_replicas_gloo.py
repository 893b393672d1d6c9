"""N > 1 host logic on CPU: replica sharding over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from kmcislandgrowth_b200 import replicas


def _fake_runner(r, grid, n_events, device, **kw):
    T, flux = grid[r]
    return {"replica": r, "temperature": T, "flux": flux, "events": n_events, "rank_device": device}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = replicas.sweep_grid()
    res = replicas.run_sweep(grid, 5, runner=_fake_runner, world_size=world, rank=rank, device=rank, gather=replicas.dist_gather)
    dist.barrier()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_assignment_partitions_the_sweep():
    grid = replicas.sweep_grid()
    assert len(grid) == 64 and grid[0] == (250.0, 0.01) and abs(grid[-1][1] - 10.0) < 1e-12
    for world in (1, 2, 4, 8):
        owned = [replicas.assign(64, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(64))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) == 0
    gv = replicas.replica_constants(9, grid)
    assert gv.beta == 1.0 / grid[9][0] / 8.617e-5 / 4 and gv.deposition_rate == grid[9][1]      # Constants.py:45-46


def test_sweep_over_two_gloo_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [s["replica"] for s in res] == list(range(64))
    assert all(s["rank_device"] == s["replica"] % 2 for s in res)
    single = replicas.run_sweep(replicas.sweep_grid(), 5, runner=_fake_runner)
    assert [(s["replica"], s["temperature"], s["flux"]) for s in single] == [(s["replica"], s["temperature"], s["flux"]) for s in res]
