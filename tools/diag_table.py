"""Dump of one pose's leaf table (binned visibility kernel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs
from inspection_path_planning_b200 import cpp_binding
mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
c, _bb = g.scene_info()
eye = c + np.array([6.0, 1.0, 1.0])
t = g.leaf_table(eye, c)
n = t["count"]
print("T", g.T, "entries", n)
zq = t["zq"]
print("sorted:", bool(np.all(np.diff(zq) >= 0)), "zq range", zq.min() if n else None, zq.max() if n else None)
w = t["tx1"] - t["tx0"] + 1; h = t["ty1"] - t["ty0"] + 1
print("rect width tiles: mean %.2f max %d; height mean %.2f max %d; triangles/leaf mean %.2f" % (w.mean(), w.max(), h.mean(), h.max(), t["cnt"].mean()))
print("hist of w*h:", np.bincount(np.minimum(w * h, 40))[:41])
print("first 10:", list(zip(zq[:10], t["tx0"][:10], t["tx1"][:10], t["ty0"][:10], t["ty1"][:10], t["first"][:10], t["cnt"][:10])))
print("tile mask rows:", [hex(int(x)) for x in t["tile_mask"]])
img = g.render_ids(eye, c, (0, 0, 1))
print("covered pixels", int((img >= 0).sum()))
