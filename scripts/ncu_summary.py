"""Turn the two ncu artefacts of scripts/gpu_check.sh into the committed summaries under profiles/.

    python scripts/ncu_summary.py TAG [line_searches_in_captured_launch]

reads  gpurun_out/TAG_launches.csv           (ncu --metrics gpu__time_duration.sum ... python bench.py)
       gpurun_out/prof_relax2_TAG.ncu-rep    (ncu --set full -k regex:k_relax2 -s 1 -c 1 python scripts/profile_relax.py 4000 200)
writes profiles/TAG_launches.csv (copy), profiles/TAG_relax2_ncu_summary.md, profiles/relax2_ncu_latest.json (what bench.py's
roofline.traffic reads).  Needs the `ncu` binary (present in the build container) to decode the report."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
n_ls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list
src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ix = {name: k for k, name in enumerate(rows[h])}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) < len(rows[h]) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(r[ix["Metric Unit"]], 1)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
shutil.copy(src, os.path.join(out_dir, tag + "_launches.csv"))

# ---- full capture
rep = os.path.join(ROOT, "gpurun_out", "prof_relax2_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[2]


def m(name, scale=1.0):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1.0, "ns": 1e-6}.get(u.split("/")[0], 1.0)
    return v * mult * scale


dur_ms = m("gpu__time_duration.sum")
dram = m("dram__bytes_read.sum") + m("dram__bytes_write.sum")
stalls = sorted(((float(vals[i]), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                 for i, n in enumerate(hdr) if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")
                 and "not_issued" not in n), reverse=True)
lines = ["# %s: launch list of bench.py and one full ncu capture of `k_relax2`" % tag, "",
         "Produced by `scripts/gpu_check.sh %s` on a B200 (every ncu pass after the same command exited 0 without ncu) and" % tag,
         "`scripts/ncu_summary.py %s`.  Times under ncu are cold-cache and serialised: shares are comparable, absolute values are not bench values." % tag, "",
         "## Launch list (`profiles/%s_launches.csv`, %d launches, %.1f ms)" % (tag, sum(cnt.values()), T / 1e6), "",
         "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for name, v in tot.most_common(12):
    lines.append("| `%s` | %d | %.3f | %.2f %% |" % (name[:70], cnt[name], v / 1e6, 100 * v / T))
lines += ["", "## Full capture: one launch = %d line searches on the 256 x 256 geometry" % n_ls, "", "| metric | value |", "|---|---:|",
          "| duration | %.3f ms (%.2f us per line search) |" % (dur_ms, dur_ms * 1e3 / n_ls),
          "| grid x block, registers, dynamic smem | %d x %d, %d regs/thread, %.1f KB/block |" % (
              m("launch__grid_size"), m("launch__block_size"), m("launch__registers_per_thread"), m("launch__shared_mem_per_block_dynamic") / 1e3),
          "| DRAM read + written | %.2f MB per launch = %.1f KB per line search |" % (dram / 1e6, dram / n_ls / 1e3),
          "| L2 / L1 hit rate | %.1f %% / %.1f %% |" % (m("lts__t_sector_hit_rate.pct"), m("l1tex__t_sector_hit_rate.pct")),
          "| warp instructions | %.3g (%.3g per line search) |" % (m("smsp__inst_executed.sum"), m("smsp__inst_executed.sum") / n_ls),
          "| issue slot utilisation | %.1f %% |" % m("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "| FP64 pipe | %.1f %% of peak |" % m("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
          "| ALU / LSU / FMA pipe | %.1f / %.1f / %.1f %% |" % (m("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                                              m("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                                                              m("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")),
          "| warps per scheduler: active / eligible | %.2f / %.2f |" % (m("smsp__warps_active.avg.per_cycle_active"), m("smsp__warps_eligible.avg.per_cycle_active")),
          "", "Stall reasons (warps stalled per issued instruction): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:9]) + "."]
open(os.path.join(out_dir, tag + "_relax2_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump({"revision": tag, "source": "ncu --set full --clock-control none --import-source on -k regex:k_relax2 -s 1 -c 1 python scripts/profile_relax.py 4000 %d" % n_ls,
           "line_searches_in_launch": n_ls, "gpu_time_ms": dur_ms, "dram_bytes_read": m("dram__bytes_read.sum"), "dram_bytes_written": m("dram__bytes_write.sum"),
           "dram_bytes_per_line_search": dram / n_ls,
           "note": "the whole working set (< 4 MB) lives in the 126 MB L2; DRAM traffic is first-touch only"},
          open(os.path.join(out_dir, "relax2_ncu_latest.json"), "w"), indent=1)
print("\n".join(lines))
