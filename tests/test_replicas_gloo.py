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


# ---- dynamic assignment + the batched engine's host logic (no GPU: a stand-in for the CUDA binding) ---------------------
class _FakeCtx(object):
    """mimics runtime.Context for ReplicaBatch: an enqueued event becomes ready after a replica-dependent number of polls"""
    polls = 0

    def __init__(self, params, device):
        self.params, self.n, self.wait, self.pending = params, 0, 0, None

    def set_params(self, p):
        self.params = p

    def state_set(self, xy):
        self.n = 0

    def bins_create(self, sub):
        class B(object):
            def close(self_):
                pass
        return B()

    def event_enqueue(self, sb, u0, u1, cap):
        assert self.pending is None
        self.pending = (u0, u1)
        self.wait = 1 + int(u0 * 5)

    def event_ready(self):
        _FakeCtx.polls += 1
        self.wait -= 1
        return self.wait <= 0

    def event_finish(self):
        import types
        u0, u1 = self.pending
        self.pending = None
        self.n += 1
        return (0 if u0 < 0.5 else 1), (-1 if u0 < 0.5 else 0), 1.0, types.SimpleNamespace(line_searches=int(u1 * 100))

    def state_get(self):
        return np.arange(2.0 * self.n)

    def close(self):
        pass


class _FakeRuntime(object):
    Context = _FakeCtx

    @staticmethod
    def params_from(gv):
        return gv


def test_batched_engine_host_logic_keeps_per_replica_order():
    grid = replicas.sweep_grid(3, 3)
    b = replicas.ReplicaBatch(range(len(grid)), grid, device=0, width=4, runtime=_FakeRuntime)
    out = b.run(7)
    b.close()
    assert sorted(out["log"]) == list(range(9)) and len(out["summaries"]) == 9
    for r in range(9):
        us = replicas.replica_uniforms(r, 7)
        want = [((0 if u0 < 0.5 else 1), (-1 if u0 < 0.5 else 0), int(u1 * 100)) for u0, u1 in us]
        assert out["log"][r] == want                     # every replica saw ITS uniforms in order, whatever the interleaving
        assert out["adatoms"][r].size == 14
    assert len(b.pool) == 0


def _queue_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = replicas.sweep_grid(4, 5)
    q = replicas.StoreQueue(len(grid))
    res = replicas.run_sweep_batched(grid, 3, device=rank, width=3, queue=q, gather=replicas.dist_gather, runtime=_FakeRuntime)
    dist.barrier()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_dynamic_queue_over_two_gloo_ranks():
    """the shared counter hands every replica to exactly one rank; the gathered summaries cover the whole sweep"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [s["replica"] for s in res] == list(range(20))
    assert all(s["events"] == 3 for s in res)
