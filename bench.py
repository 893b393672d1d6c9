#!/usr/bin/env python3
"""bench.py -- KMC events/s (and pair-energy evaluations/s) of the hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A *step* is one KMC event of the reference's loop (2DGrowth.py:317-329): rate table -> selection ->
hop/deposit -> global relaxation, on BASELINE.json's configs[1] (256x256 periodic substrate, 10k
adatoms, bins3x3 adatom sums) -- synthetic, seeded (kmcislandgrowth_b200/workloads.py).
N > 1: independent replica trajectories, one per rank (weak scaling, no data-path collective).

value : events/s with the trajectory resident in HBM (max over ranks of CUDA-event time).
e2e   : events/s through the host API: every step uploads the adatom array from pinned host memory,
        runs the event, and reads the new array back.
roofline : dominant kernel (the 20-candidate configuration-energy sweep), ALGORITHMIC bytes per launch
         / average CUDA-event launch duration, against MEASURED_PEAKS.json's HBM copy rate.
cpu_baseline : the oracle (C port of the reference, oracle/kmc_oracle.c) on the host cores, bounded sample.

--impl reference : the same metric from the CPU oracle alone (the reference itself is Python 2 and
cannot run >= 1000 adatoms, Energy.py:118; its port is timed with all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # the replica sweep of the N > 1 extras keeps one hardware queue per stream

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kmc_events_per_s"
UNIT = "events/s"


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(object):
    """SM clocks and throttle reasons sampled DURING the timed region through NVML (nvidia_ml_py): the counters of
    the recipe's nvidia-smi line.  Samples are taken by the MAIN thread between events of the timed loop (every few
    steps): a concurrent poller -- nvidia-smi -lms or an NVML thread -- contends with the CUDA API calls of the event
    loop for the driver lock and showed up as 5-8 ms outliers per event in 2 runs out of 5."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows = []
        self.h = None
        self.how = "unavailable"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"
        except Exception:
            self.index = index

    def sample(self):
        if self.h is not None:
            nv = self.nv
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.rows.append((sm, self.mx, rs))
            return
        try:        # fallback: one nvidia-smi query
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
            c = [v.strip() for v in out.strip().split(",")]
            rs = sum(bit for (name, bit), v in zip(self.REASONS, c[2:6]) if v.lower().startswith("active"))
            self.rows.append((float(c[0]), float(c[1]), rs))
            self.how = "nvidia-smi"
        except Exception:
            pass

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": [name for name, bit in self.REASONS if bits & bit], "samples": len(self.rows), "via": self.how}


def oracle_for(gv):
    from oracle import oracle as orc
    p = orc.make_params(gv.E_a, gv.E_s, gv.r_a, gv.r_s, gv.W, gv.Hs, gv.Ha, gv.temperature, gv.deposition_rate,
                        aa_mode=gv.aa_mode, hop_index_mode=gv.hop_index_mode, deposit_mode=gv.deposit_mode)
    return orc, orc.Oracle(p)


def cpu_sample(gv, sub, ad, us, ls_cap, ls_per_event):
    """Bounded CPU sample of one event on the oracle (all host threads, set explicitly: torchrun exports OMP_NUM_THREADS=1).
    Timed separately: (1) rate table + selection, (2) placement, (3) the relaxation capped at 4 and at 4 + ls_cap line
    searches -- their difference / ls_cap is the MARGINAL time of one line search (cell build, first force sweep and the
    allocation are in both runs and cancel).  The event time is fixed part + set-up + marginal time x `ls_per_event`, the
    line searches the same seeded event takes (tests/golden/bench_*.json; the first entries are confirmed by the oracle
    itself in tests/test_gpu_parity.py::test_headline_events_follow_the_oracle_uncapped)."""
    orc, o = oracle_for(gv)
    orc.set_num_threads(os.cpu_count() or 1)
    threads = orc.num_threads()
    sb = o.PutInBins(sub)
    t0 = time.perf_counter()
    rate, surf, du = o.HoppingRatesDense(ad, sb)
    k, dense, rtot = o.Select(rate, us[0])
    t_rates = time.perf_counter() - t0
    t0 = time.perf_counter()
    placed = o.HopPlace(dense if gv.hop_index_mode else k, ad, sb, us[1]) if k >= 0 else o.DepositPlace(ad, sb, us[1])
    t_place = time.perf_counter() - t0
    lo = 4
    t0 = time.perf_counter()
    _, st_lo, _ = o.Relaxation(placed, sb, max_line_searches=lo)
    t_lo = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, st_hi, rc = o.Relaxation(placed, sb, max_line_searches=lo + ls_cap)
    t_hi = time.perf_counter() - t0
    n_lo, n_hi = int(st_lo.line_searches), int(st_hi.line_searches)
    if n_hi > n_lo:
        t_ls = max(t_hi - t_lo, 1e-9) / (n_hi - n_lo)
        t_setup = max(t_lo - n_lo * t_ls, 0.0)
        t_event = t_rates + t_place + t_setup + t_ls * ls_per_event
        how = ("rate table + selection + placement timed whole; relaxation: marginal time of %d line searches (runs capped at %d and %d) "
               "x %.0f line searches of the same seeded event (extrapolated)" % (n_hi - n_lo, n_lo, n_hi, ls_per_event))
    else:               # the relaxation converged inside the low cap: a complete event was timed
        t_ls = t_lo / max(n_lo, 1)
        t_event = t_rates + t_place + t_lo
        how = "complete event (%d line searches)" % n_lo
    pairs_per_ls = st_hi.pair_evals / max(n_hi, 1)
    return {"t_event": t_event, "t_ls": t_ls, "t_rates": t_rates, "t_place": t_place, "threads": threads, "how": how,
            "pair_evals_per_s": pairs_per_ls / t_ls if t_ls > 0 else 0.0, "line_searches": n_hi}


def algorithmic_bytes_line_energy(gv, sub, ad, K=20):
    """SURVEY.md §8d, batched energy: read 2p*N_a positions + 2p*N_a direction + 2p*N_s,active +
    8*n_cells_touched, write 8*K (p = 8).  N_s,active = substrate atoms in cells some adatom stencil touches."""
    n = ad.size // 2
    nx = ((ad[0::2] + gv.L / 2) / gv.bin_size).astype(np.int64)
    ny = ((ad[1::2] - gv.Dmin) / gv.bin_size).astype(np.int64)
    occ = np.zeros((gv.nbins_x, gv.nbins_y), dtype=bool)
    ok = (nx >= 0) & (nx < gv.nbins_x) & (ny >= 0) & (ny < gv.nbins_y)
    occ[nx[ok], ny[ok]] = True
    touched = np.zeros_like(occ)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            touched |= np.roll(np.roll(occ, dx, axis=0), dy, axis=1)
    sx = ((sub[0::2] + gv.L / 2) / gv.bin_size).astype(np.int64)
    sy = ((sub[1::2] - gv.Dmin) / gv.bin_size).astype(np.int64)
    oks = (sx >= 0) & (sx < gv.nbins_x) & (sy >= 0) & (sy < gv.nbins_y)
    ns_act = int(touched[sx[oks], sy[oks]].sum())
    ncells = int(touched.sum())
    return 16 * n + 16 * n + 16 * ns_act + 8 * ncells + 8 * K, {"n_adatoms": n, "n_substrate_active": ns_act, "cells_touched": ncells}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kmcislandgrowth_b200 import workloads
    gv, sub, ad = workloads.make(args.workload)
    us_all = workloads.bench_uniforms(args.warmup + args.steps)
    fx = load_fixture(args.workload)
    ls_list = fx.get("line_searches", [])
    vals = []
    info = None
    for s in range(args.warmup + args.steps):
        ls_event = float(ls_list[s]) if s < len(ls_list) else fx["line_searches_per_event"]
        info = cpu_sample(gv, sub, ad, us_all[s], args.ref_line_searches, ls_event)
        if s >= args.warmup:
            vals.append(info["t_event"])
    t = float(np.mean(vals))
    v = 1.0 / t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, gv, ad),
        "pair_evals_per_s": info["pair_evals_per_s"], "ms_per_line_search": info["t_ls"] * 1e3,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["threads"], "host_cpus": os.cpu_count(), "kind": "port", "sample": info["how"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "one CPU process on rank 0 regardless of --gpus (the reference has no multi-GPU path); value is NOT multiplied by n_gpus",
    }
    print(json.dumps(line))


def workload_config(name, gv, ad):
    return {"workload": name, "W": int(gv.W), "Hs": int(gv.Hs), "Ha": int(gv.Ha), "adatoms": int(ad.size // 2),
            "substrate_atoms": int(gv.W * gv.Hs), "aa_mode": "bins3x3" if gv.aa_mode else "all_pairs",
            "event": "rates+select+hop/deposit+global relaxation (2DGrowth.py:317-329)",
            "replicas": "one trajectory per GPU, identical seeded uniform stream on every rank (fixed per-GPU work)",
            "l2_flush": "256 MB buffer rewritten between events (inside the timed region); the working set (<2 MB) is L2 resident within an event by construction of the workload"}


def load_fixture(name, first=0, count=None):
    """line searches per event of the seeded bench trajectory, recorded on the GPU with --record-fixture
    (a property of the workload: the CPU oracle follows the same trajectory, tests/test_gpu_parity.py)"""
    path = os.path.join(ROOT, "tests", "golden", "bench_%s.json" % name)
    if os.path.exists(path):
        fx = json.load(open(path))
        ls = fx.get("line_searches", [])
        sel = ls[first:first + count] if count else ls[first:]
        if not sel:
            sel = ls
        fx["line_searches_per_event"] = float(np.mean(sel)) if sel else 400.0
        return fx
    return {"line_searches_per_event": 400.0, "note": "default; no fixture recorded yet"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from kmcislandgrowth_b200 import Growth, runtime, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    gv, sub, ad0 = workloads.make(args.workload)
    p = runtime.params_from(gv)
    sim = Growth.Simulation(adatoms=ad0, substrate=sub, device=local, params=p)
    ctx = sim.ctx
    ctx.use_torch_stream()      # the library's kernels run on torch's current stream, so torch.cuda.Event sees them
    # every rank runs its own replica of the SAME seeded trajectory, so that per-GPU work is identical as N grows
    # (weak scaling measures the hardware, not the variance of relaxation lengths between random streams)
    cap = args.max_line_searches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def progress(msg):
        if rank == 0:
            print("[bench %.1fs] %s" % (time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    t_start = time.perf_counter()
    sampler = ClockSampler(local)
    # untimed set-up: the synthetic film is relaxed to the reference's own convergence criterion (2DGrowth.py:131)
    t_pre = time.perf_counter()
    st, prerelax_ls = workloads.bench_prerelax(ctx, sim.sbins, args.prerelax)
    t_pre = time.perf_counter() - t_pre
    progress("set-up relaxation: %d line searches, status %d" % (prerelax_ls, int(st.status)))
    fp64_peak_tflops, _ = ctx.fp64_peak(5)            # measured non-tensor FP64 FMA rate of this device
    progress("fp64 peak %.1f TFLOP/s" % fp64_peak_tflops)
    nsteps = args.warmup + args.steps
    us = workloads.bench_uniforms(max(nsteps, 64))
    if args.record_fixture and rank == 0:
        nrec = max(nsteps, 64)
        recs = [sim.step(us[s, 0], us[s, 1], cap) for s in range(nrec)]
        out = {"workload": args.workload, "prerelax_line_searches": int(prerelax_ls), "prerelax_status": int(st.status),
               "line_searches": [int(r["line_searches"]) for r in recs],
               "kinds": [int(r["kind"]) for r in recs], "restarts": [int(r["restarts"]) for r in recs],
               "note": "recorded by bench.py --record-fixture on a B200; the first entries are re-derived by the CPU oracle in tests/test_gpu_parity.py"}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_%s.json" % args.workload), "w"))
        print(json.dumps(out))
        sim.close()
        return
    # L2 hygiene: the working set (< 2 MB) is far below the 126 MB L2, so a 256 MB buffer is rewritten between events
    # (inside the timed regions, ~0.05 ms per step); within an event the persistent kernel necessarily runs from L2.
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    for s in range(args.warmup):          # warm-up events, with the flush (its first call costs tens of ms: lazy module load)
        flush.zero_()
        sim.step(us[s, 0], us[s, 1], cap)
    sampler.sample()
    progress("warm-up done")

    # ---- timed region 1: device-resident trajectory
    state0 = sim.adatoms.copy()           # both timed regions run the SAME events from this state
    ctx.timing_reset()
    ctx.timing(True)
    l0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    t_begin1 = t0
    e0.record()
    recs = []
    step_wall = []
    for s in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        _ts = time.perf_counter()
        recs.append(sim.step(us[args.warmup + s, 0], us[args.warmup + s, 1], cap))
        step_wall.append(time.perf_counter() - _ts)
        if s % max(1, args.steps // 5) == 0 or s == args.steps - 1:
            sampler.sample()          # inside the timed region, between two events
    e1.record()
    barrier()
    t_end = time.perf_counter()
    wall_dev = t_end - t0
    t_dev = e0.elapsed_time(e1) * 1e-3        # CUDA events on the launching stream
    launches = ctx.launch_count() - l0
    ctx.timing(False)
    fam = {}
    for f in ("relax_persistent", "sweep_energy", "sweep_forces", "sweep_rates", "bin_", "line_", "misc_", "select", "probe", "hop_", "event_"):
        ms, cnt = ctx.timing_read(f)
        fam[f] = {"ms": ms, "launches": cnt}
    ctx.timing_reset()

    progress("timed region 1 done: %.1f ms per event" % (t_dev / max(args.steps, 1) * 1e3))
    # ---- timed region 2: the same events through the host API (upload from pinned memory, event, read back)
    host = torch.from_numpy(state0.copy()).pin_memory()
    n_at = host.numel() // 2
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for s in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        ctx.state_set(host.numpy())
        h2d += host.numel() * 8
        rec = sim.step(us[args.warmup + s, 0], us[args.warmup + s, 1], cap)
        out = ctx.state_get()
        d2h += out.size * 8
        if out.size == host.numel():
            host.numpy()[:] = out
        else:
            host = torch.from_numpy(out.copy()).pin_memory()
    barrier()
    t_e2e = time.perf_counter() - t0
    progress("timed region 2 done")
    clocks = sampler.summary()

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])

    sweep64 = domain_cfg5 = None
    if world > 1 and not args.no_sharded:
        sweep64, domain_cfg5 = sharded_configs(args, world, rank, progress)
    if rank == 0:
        peaks, peak_src = load_peaks()
        ls = [r["line_searches"] for r in recs]
        ls_mean = float(np.mean(ls)) if ls else 0.0
        # dominant kernel: the persistent relaxation kernel; one launch = one Relaxation = `ls` line searches, each
        # the 20-candidate energy (32 B/adatom + 16 B/active substrate atom + 8 B/cell + 160 B) plus one force
        # sweep (40 B/adatom + 16 B/active substrate atom + 8 B/cell)  -- SURVEY.md 8d
        abytes_line, binfo = algorithmic_bytes_line_energy(gv, sub, ad0)
        abytes_force = 40 * binfo["n_adatoms"] + 16 * binfo["n_substrate_active"] + 8 * binfo["cells_touched"]
        ek = fam["relax_persistent"]
        k_ms = ek["ms"] / max(ek["launches"], 1)
        abytes = (abytes_line + abytes_force) * ls_mean
        achieved = abytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        total_kernel_ms = sum(v["ms"] for v in fam.values())
        value = world * args.steps / t_dev
        e2e = world * args.steps / t_e2e
        cpu = None
        if not args.no_cpu:
            info = cpu_sample(gv, sub, ad0, us[args.warmup], args.ref_line_searches, ls_mean if ls_mean > 0 else 400.0)
            cpu = {"value": 1.0 / info["t_event"], "unit": UNIT, "cores": info["threads"], "host_cpus": os.cpu_count(), "kind": "port",
                   "sample": info["how"], "pair_evals_per_s": info["pair_evals_per_s"], "ms_per_line_search": info["t_ls"] * 1e3}
        # pair evaluations: per line search 20 energy sweeps + 1 force sweep
        pr = pairs_per_sweep(ctx, sim)
        pair_evals = sum(ls) * (20 * pr["energy"] + pr["forces"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, gv, ad0),
            "pair_evals_per_s": world * pair_evals / t_dev, "line_searches_per_event": ls_mean,
            "us_per_line_search": t_dev / max(sum(ls), 1) * 1e6,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps},
            "gpu_launches": int(launches), "wall_s_timed_region": wall_dev,
            "step_wall_ms": [round(1e3 * v, 2) for v in step_wall],
            "clocks": clocks,
            "roofline": roofline_block(peaks, peak_src, fp64_peak_tflops, k_ms, ek, ls_mean, pr, abytes, abytes_line + abytes_force, binfo,
                                       total_kernel_ms),
            "kernel_families_ms": {k: round(v["ms"], 3) for k, v in fam.items()},
            "cpu_baseline": cpu,
            "prerelax": {"line_searches": int(prerelax_ls), "status": int(st.status), "seconds": round(t_pre, 3)},
            "fp64_peak_tflops_measured": fp64_peak_tflops,
            "sweep64": sweep64, "domain_cfg5": domain_cfg5,
        }
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def sharded_configs(args, world, rank, progress):
    """N > 1: the two configurations of BASELINE.json that shard (after the headline measurement, untimed for `value`):
    configs[3] -- the 64-replica temperature x flux sweep on the batched engine with dynamic assignment; configs[4] -- the
    4096x4096 substrate with bin-column domain decomposition (NCCL halo exchange).  -> (sweep64, domain_cfg5) on rank 0."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    sweep64 = domain_cfg5 = None
    try:
        import bench_sweep64
        sweep64 = bench_sweep64.run(nev=args.sweep_events, width=8, tag="bench")
        progress("sweep64 done")
    except Exception as e:                      # a failure here must not take the headline line down
        sweep64 = {"error": repr(e)[:300]}
    try:
        import bench_domain
        domain_cfg5 = bench_domain.run("cfg5_4096x4096", reps=20, n_events=args.domain_events, max_ls=150)
        progress("domain_cfg5 done")
    except Exception as e:
        domain_cfg5 = {"error": repr(e)[:300]}
    return sweep64, domain_cfg5


NCU_CAPTURE = os.path.join(ROOT, "profiles", "relax2_ncu_latest.json")


def roofline_block(peaks, peak_src, fp64_peak, k_ms, ek, ls_mean, pr, abytes, abytes_per_ls, binfo, total_kernel_ms):
    """The dominant kernel is k_relax2 (one launch = one Relaxation).  It is FP64-pipe bound by construction (about 80 flop per
    algorithmic byte; its working set is L2 resident), so the primary roof is the MEASURED non-tensor FP64 FMA rate
    of this device (kmc_fp64_peak: register-resident DFMA chains at full occupancy); the algorithmic-bytes / HBM figure is kept
    beside it.  FP64 work model per line search: the energy phase forms ONE Taylor polynomial per unique pair (about 91 FP64
    instructions, FMA = 2 flop; the reference's formulation would need (10 + 14 x 20) flop for the same pair), the force phase
    about 26 FP64 instructions per ordered pair (reference formulation: 24 flop)."""
    sec = k_ms * 1e-3
    issued = (2 * 91 * pr["energy"] + 2 * 26 * pr["forces"]) * ls_mean / sec / 1e12 if sec > 0 else 0.0
    ref_form = ((10 + 14 * 20) * pr["energy"] + 24 * pr["forces"]) * ls_mean / sec / 1e12 if sec > 0 else 0.0
    hbm = abytes / sec / 1e9 if sec > 0 else 0.0
    traffic, traffic_src = None, None
    try:                # dram__bytes_read.sum + dram__bytes_write.sum of the NEWEST committed ncu --set full capture of this kernel
        cap = json.load(open(NCU_CAPTURE))
        traffic = cap["dram_bytes_per_line_search"] * ls_mean
        traffic_src = "%s (%s; %.1f kB per line search x this run's line searches per launch)" % (
            os.path.relpath(NCU_CAPTURE, ROOT), cap.get("revision", "?"), cap["dram_bytes_per_line_search"] / 1e3)
    except Exception:
        pass
    return {"bound": "fp64", "kernel": "k_relax2 (persistent: one launch = one Relaxation; per line search 20 candidate energies + 1 force sweep)",
            "achieved": issued, "peak": fp64_peak, "unit": "TFLOP/s", "frac": issued / fp64_peak if fp64_peak else None,
            "peak_source": "measured in this run (kmc_fp64_peak, DFMA chains)",
            "achieved_reference_formulation": ref_form, "frac_reference_formulation": ref_form / fp64_peak if fp64_peak else None,
            "hbm": {"achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": abytes, "algorithmic_bytes_per_line_search": abytes_per_ls, "bytes_detail": binfo},
            "traffic": traffic, "traffic_source": traffic_src,
            "avg_launch_us": k_ms * 1e3, "launches": ek["launches"], "us_per_line_search_in_kernel": k_ms * 1e3 / ls_mean if ls_mean else None,
            "kernel_share_of_gpu_time": ek["ms"] / total_kernel_ms if total_kernel_ms else None}


def pairs_per_sweep(ctx, sim):
    """ordered pair visits of one energy sweep / one force sweep on the current state"""
    ad = sim.adatoms
    n = ad.size // 2
    ab = ctx.bins_create(ad)
    _, offs_a = ctx.nearby_atoms(ab, ad)
    _, offs_s = ctx.nearby_atoms(sim.sbins, ad)
    ab.close()
    aa = int(offs_a[n]) if ctx.params.aa_mode else n * n
    asub = int(offs_s[n])
    return {"energy": (aa // 2 if ctx.params.aa_mode else n * (n - 1) // 2) + asub, "forces": aa + asub}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_256x256_10k")
    ap.add_argument("--max-line-searches", type=int, default=0, help="0 = the reference's unbounded relaxation")
    ap.add_argument("--prerelax", type=int, default=200000, help="line-search bound of the untimed set-up relaxation (it converges before)")
    ap.add_argument("--ref-line-searches", type=int, default=30, help="line searches of the marginal-time CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sweep64 / domain_cfg5 extra measurements")
    ap.add_argument("--sweep-events", type=int, default=30, help="events per replica of the sweep64 extra measurement")
    ap.add_argument("--domain-events", type=int, default=6, help="events of the domain_cfg5 extra measurement")
    ap.add_argument("--no-flush", action="store_true", help="skip the 256 MB L2 flush between events")
    ap.add_argument("--record-fixture", action="store_true", help="record line searches per event of the seeded trajectory")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
