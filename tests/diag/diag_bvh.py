import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs
from inspection_path_planning_b200 import cpp_binding
from oracle.oracle import Oracle
mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(64))
o = Oracle(mesh_path(mesh), specs(8), post_processing=True)
V = o.vertices().astype(np.float64)
rng = np.random.default_rng(5)
c, bb = g.scene_info()
n = 20000
org = c + rng.normal(size=(n, 3)) * (bb[3:] - bb[:3]).max()
tgt = bb[:3] + rng.random((n, 3)) * (bb[3:] - bb[:3])
d = tgt - org
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, d], axis=1).astype(np.float32)
ids, brute, t, mism = g.bvh_selftest_ex(rays)
print("mismatches", mism)
bad = np.nonzero(ids != brute)[0]
def isect(r, tri):
    o_, d_ = r[:3].astype(np.float64), r[3:].astype(np.float64)
    a, b, c_ = tri
    n_ = np.cross(b - a, c_ - a)
    den = n_ @ d_
    tt = (n_ @ (a - o_)) / den if den != 0 else np.inf
    p = o_ + tt * d_
    # barycentric
    def e(p0, p1): return np.cross(p1 - p0, p - p0) @ n_
    return tt, (e(a, b), e(b, c_), e(c_, a)), n_
for i in bad[:12]:
    print("ray", i, rays[i], "bvh id", ids[i], "t", t[i, 0], "| brute id", brute[i], "t", t[i, 1])
    for nm, k in (("bvh", ids[i]), ("brute", brute[i])):
        if k >= 0:
            tt, es, n_ = isect(rays[i], V[k])
            lo, hi = V[k].min(0), V[k].max(0)
            print("   ", nm, k, "t64 %.9g" % tt, "edge fn", ["%.3e" % x for x in es], "|n| %.3e" % np.linalg.norm(n_), "tri bbox", lo, hi)
