"""Do two builds of libipp.so see the same triangles?  Edge bitmaps of the same random plans from both; the triangles they disagree
on are classified by the oracle (strict: must be seen, lenient: may be seen).

    python tools/ab_bitmaps.py <mesh> <a.so> <b.so> [plans]
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if sys.argv[1] == "--child":
    from conftest import mesh_path, random_plans, specs
    from inspection_path_planning_b200 import cpp_binding
    mesh, out, n = sys.argv[2], sys.argv[3], int(sys.argv[4])
    g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(1024))
    plans = random_plans(np.random.default_rng(4), g.getNumberOfBoxes(), n, (4, 8))
    g.evaluatePopulation(plans, True, True)
    edges = sorted({(p[k][0], p[k + 1][0]) for p in plans for k in range(len(p) - 1)})
    np.savez(out, edges=np.array(edges), maps=np.stack([g.edge_bitmap(a, b) for a, b in edges]))
    sys.exit(0)

mesh, so_a, so_b = sys.argv[1], sys.argv[2], sys.argv[3]
n = sys.argv[4] if len(sys.argv) > 4 else "40"
tree = os.path.join(ROOT, "inspection_path_planning_b200", "libipp.so")
saved = open(tree, "rb").read()
res = []
try:
    for k, so in enumerate((so_a, so_b)):
        open(tree, "wb").write(open(so, "rb").read())
        subprocess.check_call([sys.executable, __file__, "--child", mesh, "/tmp/ab_%d.npz" % k, n])
        res.append(np.load("/tmp/ab_%d.npz" % k))
finally:
    open(tree, "wb").write(saved)
A, B = res
assert np.array_equal(A["edges"], B["edges"])
diff = A["maps"] ^ B["maps"]
per_edge = np.array([int(np.unpackbits(d.view(np.uint8)).sum()) for d in diff])
seen = int(np.unpackbits(A["maps"].view(np.uint8)).sum())
print("%d edges, %d seen bits, %d bits differ on %d edges" % (len(per_edge), seen, per_edge.sum(), int((per_edge > 0).sum())))
if per_edge.sum():
    from conftest import mesh_path, specs
    from oracle.oracle import Oracle
    o = Oracle(mesh_path(mesh), specs(1024))
    bad = 0
    for e in np.nonzero(per_edge)[0][:6]:
        a, b = (int(x) for x in A["edges"][e])
        s_seen, strict, lenient = o.edge_bitmap(a, b, classify=True)
        d = diff[e]
        in_strict = int(np.unpackbits((d & strict).view(np.uint8)).sum())
        outside = int(np.unpackbits((d & ~lenient).view(np.uint8)).sum())
        print("edge %d-%d: %d differ, %d of them strict (one build misses a sure triangle), %d outside lenient (one build sees an impossible one); "
              "A vs oracle %d flips, B vs oracle %d flips" % (a, b, per_edge[e], in_strict, outside,
              int(np.unpackbits((A["maps"][e] ^ s_seen).view(np.uint8)).sum()), int(np.unpackbits((B["maps"][e] ^ s_seen).view(np.uint8)).sum())))
        bad += in_strict + outside
    print("outside the edge-epsilon band:", bad)
