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
    """Bounded CPU sample of one event on the oracle: the rate table, the selection, the placement and
    the first `ls_cap` line searches of the relaxation; the event time is completed with the measured
    per-line-search time x the line searches per event of this workload (`ls_per_event`)."""
    orc, o = oracle_for(gv)
    threads = orc.num_threads()
    sb = o.PutInBins(sub)
    t0 = time.perf_counter()
    ad2, kind, hk, rtot, st, rc = o.Event(ad, sb, us, max_line_searches=ls_cap)
    t_total = time.perf_counter() - t0
    # time the non-relaxation part separately (rate table + select + placement)
    t1 = time.perf_counter()
    rate, surf, du = o.HoppingRatesDense(ad, sb)
    o.Select(rate, us[0])
    t_rates = time.perf_counter() - t1
    nls = max(int(st.line_searches), 1)
    t_ls = max(t_total - t_rates, 1e-9) / nls
    if rc == 0:       # converged inside the cap: a complete event
        t_event, how = t_total, "complete event (%d line searches)" % nls
    else:
        t_event = t_rates + t_ls * ls_per_event
        how = ("first %d line searches of one event + its rate table; event time extrapolated to %.0f line searches/event "
               "(mean of this workload on the GPU)" % (nls, ls_per_event))
    pairs_per_ls = st.pair_evals / nls if st.line_searches else 0.0
    return {"t_event": t_event, "t_ls": t_ls, "t_rates": t_rates, "threads": threads, "how": how,
            "pair_evals_per_s": pairs_per_ls / t_ls if t_ls > 0 else 0.0, "line_searches": nls}


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
    rng = np.random.default_rng(1234)
    fx = load_fixture(args.workload, args.warmup, args.steps)
    vals = []
    info = None
    for s in range(args.warmup + args.steps):
        us = rng.uniform(0, 1, 2)
        info = cpu_sample(gv, sub, ad, us, args.ref_line_searches, fx["line_searches_per_event"])
        if s >= args.warmup:
            vals.append(info["t_event"])
    t = float(np.mean(vals))
    v = 1.0 / t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, gv, ad),
        "pair_evals_per_s": info["pair_evals_per_s"],
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": info["how"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        fx["line_searches_per_event"] = float(np.mean(sel)) if sel else 1500.0
        return fx
    return {"line_searches_per_event": 1500.0, "note": "default; no fixture recorded yet"}


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
    rng = np.random.default_rng(1234)
    cap = args.max_line_searches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    # untimed set-up: bring the synthetic film to a relaxed state (bounded)
    st = ctx.relax(sim.sbins, 16, 1e-4, args.prerelax)
    nsteps = args.warmup + args.steps
    us = rng.uniform(0, 1, (2 * nsteps, 2))
    if args.record_fixture and rank == 0:
        recs = [sim.step(us[s, 0], us[s, 1], cap) for s in range(nsteps)]
        out = {"workload": args.workload, "prerelax": args.prerelax, "line_searches": [int(r["line_searches"]) for r in recs],
               "kinds": [int(r["kind"]) for r in recs], "note": "recorded by bench.py --record-fixture on a B200"}
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
    for f in ("relax_persistent", "sweep_energy", "sweep_forces", "sweep_rates", "bin_", "line_", "misc_", "select", "probe", "hop_"):
        ms, cnt = ctx.timing_read(f)
        fam[f] = {"ms": ms, "launches": cnt}
    ctx.timing_reset()

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
    clocks = sampler.summary()

    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tt[0]), float(tt[1])

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
        if world == 1 and not args.no_cpu:
            info = cpu_sample(gv, sub, ad0, us[0], args.ref_line_searches, ls_mean if ls_mean > 0 else 1500.0)
            cpu = {"value": 1.0 / info["t_event"], "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": info["how"],
                   "pair_evals_per_s": info["pair_evals_per_s"], "ms_per_line_search": info["t_ls"] * 1e3}
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
            "roofline": {"bound": "hbm", "kernel": "k_relax2 (persistent: one launch = one Relaxation; per line search 20 candidate energies + 1 force sweep)",
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of k_relax2 from the committed ncu --set full capture
                         # (profiles/r1m_relax2_ncu.md, revision p: 1.017 MB for a 200-line-search launch), scaled to this run's
                         # line searches per launch: the working set is L2 resident, DRAM sees first touches only
                         "traffic": 1016832.0 / 200.0 * ls_mean, "traffic_source": "profiles/r1m_relax2_ncu.md, revision p (5.1 kB per line search)",
                         "algorithmic_bytes_per_launch": abytes, "bytes_detail": binfo,
                         "avg_launch_us": k_ms * 1e3, "launches": ek["launches"],
                         "kernel_share_of_gpu_time": ek["ms"] / total_kernel_ms if total_kernel_ms else None,
                         "binding_roof": "fp64 pipe (about 80 flop per algorithmic byte, SURVEY.md §8d)",
                         "algorithmic_bytes_per_line_search": abytes_line + abytes_force,
                         # flops the reference's formulation spends on the same work (10 + 14 per candidate per energy pair, 24 per
                         # force pair) / kernel time, and the flops this kernel actually issues: the 20 candidates of a
                         # pair are ONE degree-6 polynomial (~91 FP64 instructions, FMA = 2 flops) instead of 20 evaluations
                         "fp64_gflops_reference_formulation": ((10 + 14 * 20) * pr["energy"] + 24 * pr["forces"]) * ls_mean / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0,
                         "fp64_gflops_issued": (2 * 91 * pr["energy"] + 2 * 26 * pr["forces"]) * ls_mean / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0},
            "kernel_families_ms": {k: round(v["ms"], 3) for k, v in fam.items()},
            "cpu_baseline": cpu,
            "prerelax": {"line_searches": int(st.line_searches), "status": int(st.status)},
        }
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--prerelax", type=int, default=4000)
    ap.add_argument("--ref-line-searches", type=int, default=6, help="line searches per bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the 256 MB L2 flush between events")
    ap.add_argument("--record-fixture", action="store_true", help="record line searches per event of the seeded trajectory")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
