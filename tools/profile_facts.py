"""profiles/traffic.json from an `ncu --set full` capture of the K4 kernels on the bench workload: what only a profiler can count, per
32x32 tile the kernel rasterised (bench.py multiplies by the tiles of its own timed launches).

    python tools/profile_facts.py profiles/r2_ncu_full_bins_raw.csv gpurun_out/r2_diag_ncu.log "<what the capture is of>"
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, log, what = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
tiles = int(re.search(r"\(tiles (\d+)\)", open(log).read()).group(1))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def val(r, name):
    return float(r[col[name]]) * SCALE.get(units[col[name]], 1.0)


out = {"source": "%s: %s" % (os.path.relpath(raw, ROOT), what), "tiles_in_capture": tiles}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    k = "k_visibility_bins" if "k_visibility_bins" in name else ("k_leaf_table" if "k_leaf_table" in name else None)
    if not k:
        continue
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    inst = val(r, "smsp__inst_executed.sum")
    lts = val(r, "lts__t_sectors.sum") * 32.0
    out[k + "_dram_bytes_per_launch"] = dram
    out[k + "_warp_inst_per_launch"] = inst
    out[k + "_lts_bytes_per_launch"] = lts
    out[k + "_ms_in_capture"] = float(r[col["gpu__time_duration.sum"]])
    out[k + "_issue_active_pct"] = float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
    out[k + "_warps_active_pct"] = float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])
    if k == "k_visibility_bins":
        out[k + "_warp_inst_per_tile"] = inst / tiles
        out[k + "_lts_bytes_per_tile"] = lts / tiles
        out[k + "_dram_bytes_per_tile"] = dram / tiles
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
