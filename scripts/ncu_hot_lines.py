"""Top source lines of k_relax2 by executed instructions and by warp-stall samples, from the full ncu capture of scripts/gpu_check.sh.

    python scripts/ncu_hot_lines.py TAG [n_lines]      ->  profiles/TAG_relax2_hot_lines.md
(reads gpurun_out/prof_relax2_TAG.ncu-rep with `ncu --page source --print-source cuda,sass`; needs -lineinfo builds and
--import-source on captures, which is what build.py and gpu_check.sh do)"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rep = os.path.join(ROOT, "gpurun_out", "prof_relax2_%s.ncu-rep" % tag)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows, cur, hdr = [], None, None
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0] not in ("", "Function Name") and len(r) > 8:
        try:
            rows.append((int(r[7]), int(r[6]), cur, r[0], " ".join(r[1].split())[:110]))
        except ValueError:
            pass
ti, ts = sum(r[0] for r in rows), sum(r[1] for r in rows)
byfile = {}
for r in rows:
    byfile[r[2]] = byfile.get(r[2], 0) + r[0]
out = ["# %s: where `k_relax2` spends instructions and stall samples (source view of the full ncu capture)" % tag, "",
       "%.3g warp instructions, %d stall samples in one launch (200 line searches).  By file: %s." % (
           ti, ts, ", ".join("%s %.1f %%" % (k, 100.0 * v / ti) for k, v in sorted(byfile.items(), key=lambda kv: -kv[1]) if v > 0.001 * ti)), "",
       "## Top %d lines by executed instructions" % top, "", "| % instr | % samples | where | source |", "|---:|---:|---|---|"]
for r in sorted(rows, reverse=True)[:top]:
    out.append("| %.2f | %.2f | %s:%s | `%s` |" % (100.0 * r[0] / ti, 100.0 * r[1] / ts, r[2], r[3], r[4].replace("|", "\\|")))
out += ["", "## Top %d lines by stall samples" % (top // 2), "", "| % samples | % instr | where | source |", "|---:|---:|---|---|"]
for r in sorted(rows, key=lambda r: -r[1])[:top // 2]:
    out.append("| %.2f | %.2f | %s:%s | `%s` |" % (100.0 * r[1] / ts, 100.0 * r[0] / ti, r[2], r[3], r[4].replace("|", "\\|")))
path = os.path.join(ROOT, "profiles", "%s_relax2_hot_lines.md" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12] + out[-16:]))
