"""BASELINE configs[3]: 64 independent replicas (8 temperatures x 8 fluxes, default geometry, from the empty surface) on the
batched replica engine: every rank keeps `width` replicas in flight on its GPU (one context + stream each, events queued
without host round trips) and takes replica ids from a shared counter (dynamic assignment; no data-path collective).

    python scripts/bench_sweep64.py [events_per_replica] [width] [threads]                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... scripts/bench_sweep64.py

Time = max over ranks of the wall clock between two barriers (the engine synchronises every replica's stream before it
returns, so the host clock brackets the device work).  `--static` uses the round-robin assignment of round 1 for comparison."""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # one hardware queue per replica stream (default 8: streams share queues)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from kmcislandgrowth_b200 import replicas


def run(nev=30, width=8, threads=1, static=False, tag="timed"):
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    grid = replicas.sweep_grid()

    def queue():
        if static:
            return replicas.LocalQueue(replicas.assign(len(grid), world, rank))
        return replicas.StoreQueue(len(grid), key="kmc_queue_" + tag) if world > 1 else None

    replicas.run_sweep_batched(grid[:8], 3, device=local, width=width, threads=threads, max_line_searches=20000)      # warm-up (module load, contexts)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = replicas.run_sweep_batched(grid, nev, device=local, width=width, threads=threads, queue=queue(), max_line_searches=20000,
                                     gather=replicas.dist_gather if world > 1 else None)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        longest = max(res, key=lambda r: r["line_searches"])
        out = {"config": "sweep64_T8xF8_default_geometry", "n_gpus": world, "replicas": len(grid), "events_each": nev,
               "seconds": float(dt[0]), "events_per_s": len(grid) * nev / float(dt[0]), "width": width, "threads": threads,
               "assignment": "static" if static else "dynamic",
               "line_searches": sum(r["line_searches"] for r in res), "adatoms_mean": sum(r["n_adatoms"] for r in res) / len(res),
               "longest_replica": {"replica": longest["replica"], "line_searches": longest["line_searches"],
                                   "share_of_all_line_searches": longest["line_searches"] / max(1, sum(r["line_searches"] for r in res))},
               "note": "a trajectory cannot be split: the wall clock is bounded below by the longest replica on ANY number of GPUs",
               "host_seconds_rank0": getattr(replicas.run_sweep_batched, "last_host_seconds", None),
               "checksum_hops": sum(r["hops"] for r in res)}
    if own_group:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    out = run(int(argv[0]) if len(argv) > 0 else 30, int(argv[1]) if len(argv) > 1 else 16, int(argv[2]) if len(argv) > 2 else 1,
              static="--static" in sys.argv)
    if out is not None:
        print(json.dumps(out))
