"""Latency of the small batched calls an NSGA-II generation makes (optimize_coverage.py:141,191): ~20 % of a population of 40..1000
plans re-evaluated per generation, ~10 % of their edges new.  Wall time per call against the kernel-group times of the call.

    python tools/diag_small.py <mesh> <plans per call> <calls>
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs  # noqa: E402
from inspection_path_planning_b200 import cpp_binding  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "luis4"
per_call = int(sys.argv[2]) if len(sys.argv) > 2 else 8
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(1024))
N = g.getNumberOfBoxes()
rng = np.random.default_rng(5)
pool = []
for _ in range(200):
    L = int(rng.integers(2, 16))
    p = [int(rng.integers(0, N))]
    while len(p) < L:
        x = int(rng.integers(0, N))
        if x != p[-1]:
            p.append(x)
    pool.append(p)
g.evaluatePopulation([[[x] for x in p] for p in pool], True, True)
names = ["reduce", "visibility", "collision", "pack", "plan", "table"]
walls, kern, new_edges = [], np.zeros(6), 0
for c in range(calls):
    batch = []
    for _ in range(per_call):
        p = list(pool[int(rng.integers(0, len(pool)))])
        k = int(rng.integers(0, len(p)))
        x = int(rng.integers(0, N))
        if (k == 0 or p[k - 1] != x) and (k == len(p) - 1 or p[k + 1] != x):
            p[k] = x  # one mutated viewpoint: up to two new edges
        batch.append(p)
    plans = [[[x] for x in p] for p in batch]
    for k in range(6):
        g.kernel_time(k, reset=True)
    c0 = g.counters()["edges_rendered"]
    t = time.perf_counter()
    g.evaluatePopulation(plans, True, False)
    walls.append(time.perf_counter() - t)
    new_edges += g.counters()["edges_rendered"] - c0
    kern += np.array([g.kernel_time(k)[0] for k in range(6)])
    pool.extend(batch)
w = np.array(walls) * 1e3
print("%s: %d calls of %d plans, %.1f new edges per call" % (mesh, calls, per_call, new_edges / calls))
print("wall per call: median %.3f ms, mean %.3f ms, p90 %.3f ms" % (np.median(w), w.mean(), np.percentile(w, 90)))
print("kernel groups per call (ms):", {n: round(v / calls, 3) for n, v in zip(names, kern)}, "sum %.3f" % (kern.sum() / calls))
t = time.perf_counter()
for _ in range(20):
    cpp_binding.pack_population(plans)
print("pack_population of the batch: %.3f ms" % ((time.perf_counter() - t) / 20 * 1e3))
t = time.perf_counter()
for _ in range(20):
    g.evaluatePopulation(plans, True, False)
print("warm call (no new edges): %.3f ms" % ((time.perf_counter() - t) / 20 * 1e3))
g.close()
