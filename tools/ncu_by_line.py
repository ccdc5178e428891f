"""Per-source-line summary of one kernel of an ncu report: joins `ncu --page source --csv` (SASS view: instruction counts and
stall samples per instruction) with `nvdisasm -g` (source line of every instruction) of the cubin inside libipp.so.

    python tools/ncu_by_line.py report.ncu-rep k_visibility_bins visibility_bins [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kernel, unit = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
so = os.path.join(ROOT, "inspection_path_planning_b200", "libipp.so")

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", unit + ".sm_100a.cubin", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# offset -> (file, line) inside the kernel's section
line_of = {}
cur = None
inside = False
for ln in sass.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
body = rows[hdr_i + 1:]
base = int(body[0][0], 16)
by_line = defaultdict(lambda: defaultdict(float))
total = defaultdict(float)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for r in body:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    off = int(r[0], 16) - base
    key, _ = line_of.get(off, ((("?", 0)), ""))
    key = key or ("?", 0)
    inst = float(r[col["Instructions Executed"]] or 0)
    smp = float(r[col["# Samples"]] or 0)
    by_line[key]["inst"] += inst
    by_line[key]["samples"] += smp
    total["inst"] += inst
    total["samples"] += smp
    for s in stalls:
        v = float(r[col[s]] or 0)
        by_line[key][s] += v
        total[s] += v

print("kernel %s: %.4g warp instructions, %d samples" % (kernel, total["inst"], total["samples"]))
print("stall mix: " + ", ".join("%s %.1f%%" % (s[6:], 100 * total[s] / max(1, total["samples"])) for s in sorted(stalls, key=lambda s: -total[s])[:8]))
src_cache = {}


def src(f, n):
    for d in ("inspection_path_planning_b200/csrc", "include"):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""


print("%-22s %7s %7s  %-24s %s" % ("line", "inst%", "smpl%", "top stalls", "source"))
for key, v in sorted(by_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    ts = sorted(stalls, key=lambda s: -v[s])[:2]
    print("%-22s %6.2f%% %6.2f%%  %-24s %s" % ("%s:%d" % key, 100 * v["inst"] / total["inst"], 100 * v["samples"] / max(1, total["samples"]),
                                               " ".join("%s:%d%%" % (s[6:], 100 * v[s] / max(1, v["samples"])) for s in ts), src(*key)))
# regions of interest, by file
by_file = defaultdict(lambda: defaultdict(float))
for key, v in by_line.items():
    by_file[key[0]]["inst"] += v["inst"]
    by_file[key[0]]["samples"] += v["samples"]
print("by file: " + ", ".join("%s inst %.1f%% samples %.1f%%" % (f, 100 * v["inst"] / total["inst"], 100 * v["samples"] / max(1, total["samples"]))
                              for f, v in by_file.items()))
if len(sys.argv) > 5:  # dump every line, in source order, for region sums
    for key, v in sorted(by_line.items()):
        print("ALL %s:%d inst %.3f%% samples %.3f%%" % (key[0], key[1], 100 * v["inst"] / total["inst"], 100 * v["samples"] / max(1, total["samples"])))
