"""Spatial domain decomposition on CPU: world_size 2 and 3 gloo groups, local sweeps on the
oracle, compared with the single-domain oracle on the same configuration (forces, energy, rates, selection)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from kmcislandgrowth_b200 import domain, workloads
from oracle import oracle as orc

W, HS, HA, N = 48, 12, 10, 400


def _setup():
    gv = workloads.constants_for(W, HS, HA, aa_mode=1)
    sub, ad = workloads.substrate(gv), workloads.film(gv, N, seed=77)
    p = orc.make_params(W=W, Hs=HS, Ha=HA, aa_mode=1)
    return gv, sub, ad, p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.set_num_threads(1)
    gv, sub, ad, p = _setup()
    o = orc.Oracle(p)
    ds = domain.DomainSweeps(gv, sub, ad, rank, world, lambda s: domain.OracleBackend(o, s), dist=dist)
    F = ds.gather_owned(ds.forces())
    E = ds.energy()
    rate, surf = ds.rates()
    sel = ds.select(0.4321)
    dist.barrier()
    if rank == 0:
        q.put((F, E, rate, surf, sel, ds.halo_bytes, [len(n) for n in ds.dec.needs]))
    dist.destroy_process_group()


def test_decomposition_covers_every_stencil():
    gv = workloads.constants_for(256, 256, 64)
    for world in (1, 2, 4, 8):
        dec = domain.Decomposition(gv, world)
        assert dec.bounds[0][0] == 0 and dec.bounds[-1][1] == gv.nbins_x
        for r, (a, b) in enumerate(dec.bounds):
            have = set(range(a, b)) | set(dec.needs[r])
            for cx in range(a, b):
                assert domain.stencil_columns(cx, gv.nbins_x) <= have
        if world > 1:
            # the seam: rank 0 needs the last TWO columns, the last rank needs column 0 (Bins.py:66-69)
            assert gv.nbins_x - 1 in dec.needs[0] and gv.nbins_x - 2 in dec.needs[0]
            assert 0 in dec.needs[world - 1]


@pytest.mark.parametrize("world", [2, 3])
def test_domain_sweeps_match_single_domain(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    F, E, rate, surf, sel, halo_bytes, needs = q.get(timeout=300)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    gv, sub, ad, p = _setup()
    o = orc.Oracle(p)
    sb = o.PutInBins(sub)
    F0 = (o.AdatomAdatomForces(ad) + o.AdatomSubstrateForces(ad, sb)).reshape(-1, 2)
    assert np.max(np.abs(F - F0)) < 1e-11 * max(1.0, np.max(np.abs(F0)))
    E0 = o.RelaxEnergy((ad, sb))
    assert abs(E - E0) < 1e-11 * abs(E0)
    rate0, surf0, _ = o.HoppingRatesDense(ad, sb)
    assert np.array_equal(surf.astype(bool), surf0)
    assert np.allclose(rate, rate0, rtol=1e-10)
    assert sel[:2] == o.Select(rate0, 0.4321)[:2]
    assert halo_bytes > 0 and all(n >= 2 for n in needs)
