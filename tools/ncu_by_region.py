"""Per-region summary of k_visibility_bins from an ncu report taken with --import-source on: instructions issued and warp stall
samples of `ncu --page source --print-source cuda,sass`, summed over the regions of raster.cuh / visibility_bins.cu that
profiles/README.md quotes (pixel steps, block ballot, fp64 set-up, ...), and the stall reasons of the whole kernel.

    ncu -i report.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_visibility_bins > /tmp/cs.csv
    python tools/ncu_by_region.py /tmp/cs.csv
"""
import csv
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
cur = hdr = curline = None
agg = defaultdict(lambda: defaultdict(float))
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        curline = int(r[0])
    if r[2] == "":
        continue  # a source line without SASS rows of its own
    for i, h in enumerate(hdr):
        if h in ("Instructions Executed", "# Samples") or (h.startswith("stall_") and "Not Issued" not in h):
            try:
                agg[(cur, curline)][h] += float(r[i])
            except ValueError:
                pass
tot = sum(v["Instructions Executed"] for v in agg.values())
tots = sum(v["# Samples"] for v in agg.values())
stalls = defaultdict(float)
for v in agg.values():
    for k, x in v.items():
        if k.startswith("stall_"):
            stalls[k] += x
print("stall samples:", "  ".join("%s %.1f %%" % (k[6:], 100 * x / tots) for k, x in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]))

text = open(os.path.join(ROOT, "inspection_path_planning_b200", "csrc", "raster.cuh")).read().splitlines()


def line_of(pattern):
    for i, l in enumerate(text):
        if pattern in l:
            return i + 1
    return -1


L = {k: line_of(p) for k, p in {
    "inner0": "while (todo) {", "inner1": "return dirty;", "staged": "__device__ __forceinline__ bool raster_staged",
    "refresh": "void refresh_depth_bounds", "output": "void tile_output_row", "find": "int find_snapshot",
    "setup": "template <bool kPoseOnly>", "stage": "bool stage_record", "geom": "void tile_geometry",
    "exact0": "#if IPP_EXACT_TILE_MASK", "exact1": "*tilesOut = tiles;"}.items()}


def region(f, l):
    if f == "raster.cuh":
        if L["inner0"] <= l < L["inner1"]:
            return "pixel steps of the rasteriser"
        if L["staged"] - 1 <= l < L["inner0"]:
            return "block ballot per live triangle"
        if L["refresh"] <= l < L["output"]:
            return "depth-bound refresh"
        if L["output"] <= l < L["find"]:
            return "winner output"
        if L["exact0"] <= l <= L["exact1"]:
            return "exact tile masks"
        if L["setup"] <= l < L["stage"] or 28 < l < 60:
            return "fp64 pose set-up"
        if L["stage"] <= l < L["staged"] - 1:
            return "record staging"
        if L["geom"] <= l < L["setup"] - 20:
            return "tile geometry + hit-buffer init"
        return "raster.cuh, other"
    if f == "visibility_bins.cu":
        return "visibility_bins.cu (table filter, header scan, queues, work cursor)"
    return "intrinsics headers (shuffles, ballots)"


by = defaultdict(lambda: [0.0, 0.0])
for (f, l), v in agg.items():
    k = region(f, l)
    by[k][0] += v["Instructions Executed"]
    by[k][1] += v["# Samples"]
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print("%-72s %5.1f %% of the instructions  %5.1f %% of the stall samples" % (k, 100 * v[0] / tot, 100 * v[1] / tots))
