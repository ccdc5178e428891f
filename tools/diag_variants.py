"""A/B of the two edge-visibility kernels (leaf bins vs BVH walk per tile): per-edge bitmaps, fitness, kernel times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, random_plans, specs
from inspection_path_planning_b200 import cpp_binding
mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pop = int(sys.argv[3]) if len(sys.argv) > 3 else 20
V = int(sys.argv[4]) if len(sys.argv) > 4 else 20
nedges = int(sys.argv[5]) if len(sys.argv) > 5 else 40
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
N = g.getNumberOfBoxes()
print("T=%d N=%d default variant %s" % (g.T, N, g.visibility_variant()), flush=True)
rng = np.random.default_rng(20261019)
plans = cpp_binding.pack_population(random_plans(rng, N, pop, V))
names = ["reduce", "visibility", "collision", "pack", "plan", "table"]
res = {}
for variant in (False, True, False, True):
    name = g.visibility_variant(variant)
    g.clear_memo()
    g.counters(reset=True); g.visibility_stats(reset=True)
    for k in range(6): g.kernel_time(k, reset=True)
    t = time.time(); fit = g.evaluatePopulation(plans, True, True); dt = time.time() - t
    c = g.counters(); vs = g.visibility_stats()
    kt = {names[k]: round(g.kernel_time(k)[0], 3) for k in range(6)}
    print("%-4s wall %.4fs  kernel ms %s" % (name, dt, kt), flush=True)
    print("     snapshots %d rays %.3e tiles %d  per tile: accel-64B %.1f triangles %.1f" % (
        c["snapshots"], c["rays"], vs["tiles"], vs["nodes"] / max(1, vs["tiles"]), vs["triangles"] / max(1, vs["tiles"])), flush=True)
    res[name] = fit
print("fitness max |bins - bvh|:", np.abs(res["bins"] - res["bvh"]).max(axis=0))
# per-edge bitmaps
tot = 0; diff = 0
for _ in range(nedges):
    a, b = rng.integers(0, N, size=2)
    if a == b: continue
    g.visibility_variant(False); r0 = g.edge_bitmap(int(a), int(b))
    g.visibility_variant(True); r1 = g.edge_bitmap(int(a), int(b))
    g.visibility_variant(True); r2 = g.edge_bitmap(int(a), int(b))
    x = np.bitwise_xor(r0, r1)
    d = int(np.unpackbits(x.view(np.uint8)).sum())
    assert np.array_equal(r1, r2), "bins variant is not deterministic"
    tot += int(np.unpackbits(r0.view(np.uint8)).sum()); diff += d
    if d: print("edge %d->%d: %d differing triangles" % (a, b, d))
print("edges compared: seen bits %d, differing %d" % (tot, diff))
