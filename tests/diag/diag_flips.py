"""Per-edge flip diagnostics: GPU seen set vs oracle nominal/strict/lenient sets."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bit_ids, mesh_path, popcount, specs  # noqa: E402
from inspection_path_planning_b200 import cpp_binding  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
o = Oracle(mesh_path(mesh), specs(H))
if len(sys.argv) > 5:
    o.set_margins(float(sys.argv[4]), float(sys.argv[5]))
rng = np.random.default_rng(int(os.environ.get("SEED", "31")))
if os.environ.get("VARIANT"):
    print("variant", g.visibility_variant(os.environ["VARIANT"] == "bins"))
edges = [tuple(int(x) for x in e.split("-")) for e in os.environ.get("EDGES", "").split(",") if e]
N = o.getNumberOfBoxes()
area = o.areas()
tot = dict(seen=0, flips=0, outside=0, darea=0.0)
for k in range(max(n, len(edges))):
    prev, curr = edges[k] if k < len(edges) else (int(x) for x in rng.choice(N, 2, replace=False))
    if edges and k >= len(edges): break
    got = g.edge_bitmap(prev, curr)
    seen, strict, lenient = o.edge_bitmap(prev, curr, classify=True)
    extra, missing = bit_ids(got & ~seen), bit_ids(seen & ~got)
    out_extra, out_missing = bit_ids(got & ~lenient), bit_ids(strict & ~got)
    tot["seen"] += popcount(seen)
    tot["flips"] += extra.size + missing.size
    tot["outside"] += out_extra.size + out_missing.size
    tot["darea"] += area[extra].sum() - area[missing].sum()
    print("edge %d->%d seen %d strict %d lenient %d | gpu extra %d missing %d | outside: extra %s missing %s" % (
        prev, curr, popcount(seen), popcount(strict), popcount(lenient), extra.size, missing.size, out_extra.tolist(), out_missing.tolist()), flush=True)
print(tot, "T", o.T, "total area", area.sum())
