"""BASELINE configs[4]: 4096x4096 substrate, 8192 adatoms, bin-column domain decomposition over the ranks of a torchrun job,
device-resident (kmcislandgrowth_b200.domain.DeviceDomain: NCCL point-to-point halo exchange + small collectives over NVLink).

    python scripts/bench_domain.py [workload] [sweeps] [events]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_domain.py ...

Prints one JSON line from rank 0: ms per sweep (halo exchange + forces + energy + rate table + selection), the time of the
exchange alone, us per line search and events/s of the decomposed relaxation loop, and the agreement with the
single-domain result computed on rank 0's GPU by the plain (single-GPU) entry points."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from kmcislandgrowth_b200 import domain, runtime, workloads


def run(name="cfg5_4096x4096", reps=20, n_events=6, max_ls=150):
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    gv, sub, ad = workloads.make(name)
    dd = domain.DeviceDomain(gv, sub, ad, rank, world, local, dist=dist if world > 1 else None)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(dt):
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- agreement with the single-domain sweeps (same GPU, plain entry points)
    sw = dd.sweep()
    sel = dd.select(0.37, sw)
    Fg = torch.zeros((dd.n_global, 2), dtype=torch.float64, device=dev)
    Fg.index_copy_(0, dd.gid, sw["F"])
    dd._allreduce(Fg)
    check = {}
    if rank == 0:
        ref = runtime.Context(runtime.params_from(gv), local)
        F1 = np.asarray(ref.forces(ad, dd.sb, 3)).reshape(-1, 2)
        e1 = float(ref.relax_energy(ad, dd.sb)[0])
        r1 = ref.rates(ad, dd.sb)
        s1 = ref.select(r1["rate"], r1["is_surface"], 0.37)
        scale = float(np.max(np.abs(F1))) or 1.0
        check = {"force_max_abs_diff_over_scale": float(np.max(np.abs(Fg.cpu().numpy() - F1)) / scale),
                 "energy_rel_diff": abs(float(sw["energy"]) - e1) / abs(e1),
                 "rate_max_rel_diff": float(np.max(np.abs(sw["rate_dense"].cpu().numpy() - r1["rate"]) / np.maximum(np.abs(r1["rate"]), 1e-300))),
                 "selection_equal": [int(sel[0]), int(sel[1])] == [int(s1[0]), int(s1[1])], "selected": [int(sel[0]), int(sel[1])]}
        ref.close()
    # ---- sweeps
    for _ in range(3):
        dd.select(0.37, dd.sweep())
    sync()
    c0, h0 = dd.collectives, dd.halo_bytes
    t0 = time.perf_counter()
    for _ in range(reps):
        sw = dd.sweep()
        sel = dd.select(0.37, sw)
    sync()
    t_sweep = tmax(time.perf_counter() - t0) / reps
    coll_per_sweep = (dd.collectives - c0) / reps
    halo_bytes = (dd.halo_bytes - h0) // reps
    # ---- the halo exchange alone
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        dd.exchange(dd.x)
    sync()
    t_halo = tmax(time.perf_counter() - t0) / reps
    # ---- decomposed relaxation + events
    us = workloads.bench_uniforms(n_events + 2)
    dd.event(us[0, 0], us[0, 1], max_ls)              # warm-up
    sync()
    t0 = time.perf_counter()
    nls = 0
    kinds = []
    for s in range(n_events):
        st = dd.event(us[1 + s, 0], us[1 + s, 1], max_ls)
        nls += st["line_searches"]
        kinds.append(st["kind"])
    sync()
    t_ev = tmax(time.perf_counter() - t0)
    # ---- the same events on ONE GPU through the persistent relaxation kernel (rank 0; what the decomposition competes with)
    single = None
    if rank == 0:
        one = runtime.Context(runtime.params_from(gv), local)
        one.state_set(ad)
        one.event(dd.sb, us[0, 0], us[0, 1], max_ls)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nls1, kinds1 = 0, []
        for s in range(n_events):
            kind, _, _, st1 = one.event(dd.sb, us[1 + s, 0], us[1 + s, 1], max_ls)
            nls1 += int(st1.line_searches)
            kinds1.append(int(kind))
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t1
        single = {"events_per_s": n_events / dt1, "us_per_line_search": dt1 / max(nls1, 1) * 1e6, "line_searches": nls1,
                  "same_event_kinds_and_line_searches": kinds1 == [int(k) for k in kinds] and nls1 == nls}
        one.close()
    out = None
    if rank == 0:
        out = {"workload": name, "n_gpus": world, "ms_per_sweep": t_sweep * 1e3, "sweeps_per_s": 1.0 / t_sweep,
               "sweep": "halo exchange + forces + energy (all-reduce) + rate table (dense all-reduce) + sequential-order selection",
               "halo_us": t_halo * 1e6, "halo_bytes_per_sweep_rank0": int(halo_bytes), "collectives_per_sweep": coll_per_sweep,
               "events_per_s": n_events / t_ev, "events": n_events, "line_search_cap_per_event": max_ls,
               "us_per_line_search": t_ev / max(nls, 1) * 1e6, "line_searches": nls, "event_kinds": kinds,
               "owned_adatoms_rank0": dd.n_own, "local_atoms_rank0": dd.n_local, "agreement_with_single_domain": check,
               "single_gpu_persistent_kernel": single,
               "limiting_step": "per line search: 1 NCCL send/recv group (halo x and s), 1 all-reduce (20 candidate energies), 1 all-gather "
                                "(max|F|^2, top, bot, min y) and one host read of 4 doubles; the messages are KBs, the cost is their latency"}
    dd.close()
    if own_group:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_4096x4096"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    nev = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    res = run(name, reps, nev)
    if res is not None:
        print(json.dumps(res))
