"""BASELINE configs[3]: 64 independent replicas (8 temperatures x 8 fluxes, default geometry, from the empty surface),
replica r on rank r mod N, each rank running its share with 4 concurrent host threads (one context + stream per replica).

    python scripts/bench_sweep64.py [events_per_replica]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... scripts/bench_sweep64.py

No data-path collective: one all_gather_object of the summaries at the end.  Time = max over ranks of the wall clock between
two barriers (every replica synchronises its stream after every event, so the host clock brackets the device work)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from kmcislandgrowth_b200 import replicas

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
nev = int(sys.argv[1]) if len(sys.argv) > 1 else 30
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
grid = replicas.sweep_grid()
replicas.run_sweep(grid[:4 * world], 3, world_size=world, rank=rank, device=local, threads=4, max_line_searches=20000)     # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = replicas.run_sweep(grid, nev, world_size=world, rank=rank, device=local, threads=4, max_line_searches=20000,
                         gather=replicas.dist_gather if world > 1 else None)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "sweep64_T8xF8_default_geometry", "n_gpus": world, "replicas": len(grid), "events_each": nev,
                      "seconds": float(dt[0]), "events_per_s": len(grid) * nev / float(dt[0]), "threads_per_gpu": 4,
                      "line_searches": sum(r["line_searches"] for r in res), "adatoms_mean": sum(r["n_adatoms"] for r in res) / len(res),
                      "checksum_hops": sum(r["hops"] for r in res)}))
if world > 1:
    dist.destroy_process_group()
