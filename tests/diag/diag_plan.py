"""Per-edge flips of the plans of a seeded population against the oracle classes, both K4 variants."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bit_ids, mesh_path, popcount, random_plans, specs
from inspection_path_planning_b200 import cpp_binding
from oracle.oracle import Oracle
mesh, H, seed, pop = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lo, hi = int(sys.argv[5]), int(sys.argv[6])
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
o = Oracle(mesh_path(mesh), specs(H))
rng = np.random.default_rng(seed)
plans = random_plans(rng, o.getNumberOfBoxes(), pop, (lo, hi))
area = o.areas(); tot = area.sum()
for pi, plan in enumerate(plans):
    ids = [int(x[0]) for x in plan]
    for a, b in zip(ids[:-1], ids[1:]):
        seen, strict, lenient = o.edge_bitmap(a, b, classify=True)
        for var in (False, True):
            name = g.visibility_variant(var)
            got = g.edge_bitmap(a, b)
            extra, missing = bit_ids(got & ~seen), bit_ids(seen & ~got)
            oe, om = bit_ids(got & ~lenient), bit_ids(strict & ~got)
            if extra.size or missing.size:
                print("plan %d edge %d->%d %s: extra %s missing %s (area fractions %s %s) outside band: %s %s" % (
                    pi, a, b, name, extra.tolist(), missing.tolist(), (area[extra] / tot).round(7).tolist(), (area[missing] / tot).round(7).tolist(), oe.tolist(), om.tolist()), flush=True)
