"""Domain-decomposed sweeps of BASELINE config 5 (4096x4096 substrate, 8192 adatoms, bins3x3) over the ranks
of a torchrun job (NCCL halo exchange over NVLink).  Prints one JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_domain.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from kmcislandgrowth_b200 import domain, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_4096x4096"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gv, sub, ad = workloads.make(name)
ds = domain.DomainSweeps(gv, sub, ad, rank, world, lambda s: domain.GpuBackend(gv, s, local), dist=dist if world > 1 else None, device=dev)
del sub
def sweep():
    F = ds.forces()
    E = ds.energy()
    sel = ds.select(0.37)
    return F, E, sel
F, E, sel = sweep(); sweep()
if world > 1: dist.barrier()
torch.cuda.synchronize()
h0 = ds.halo_bytes
t0 = time.perf_counter()
for _ in range(reps):
    F, E, sel = sweep()
if world > 1: dist.barrier()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tt = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
Fg = ds.gather_owned(F)
if rank == 0:
    print(json.dumps({"workload": name, "n_gpus": world, "sweeps_per_s": reps / float(tt[0]), "ms_per_sweep": float(tt[0]) / reps * 1e3,
                      "sweep": "halo exchange + forces + energy + rate table + selection", "energy": E, "select": list(sel[:2]),
                      "force_checksum": float(np.abs(Fg).sum()), "halo_bytes_per_sweep_rank0": (ds.halo_bytes - h0) // reps,
                      "owned_adatoms_rank0": int(len(ds.gid)), "substrate_atoms_rank0": int(ds.backend.sb.n)}))
if world > 1: dist.destroy_process_group()
