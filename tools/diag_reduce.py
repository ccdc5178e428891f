"""plan_reduce (K6) streaming test: warm re-evaluation of a large population; kernel time from CUDA events."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs
from inspection_path_planning_b200 import cpp_binding
mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pops = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1000, 10000, 100000]
V = int(sys.argv[4]) if len(sys.argv) > 4 else 20
nedges = int(sys.argv[5]) if len(sys.argv) > 5 else 20000
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
N = g.getNumberOfBoxes(); T = g.T
rng = np.random.default_rng(1)
# restrict plans to a pool of `nedges` distinct edges so that the cold render stays short: walk over a random subgraph
keys = rng.choice(N * N, size=nedges, replace=False)
a, b = keys // N, keys % N
ok = a != b
a, b = a[ok], b[ok]
succ = {}
for x, y in zip(a, b):
    succ.setdefault(int(x), []).append(int(y))
starts = [k for k in succ]
def walk(L):
    while True:
        p = [starts[rng.integers(len(starts))]]
        while len(p) < L:
            nx = succ.get(p[-1])
            if not nx: break
            p.append(nx[rng.integers(len(nx))])
        if len(p) == L: return p
for pop in pops:
    ids = np.array([walk(V) for _ in range(pop)], dtype=np.int32).reshape(-1)
    off = np.arange(pop + 1, dtype=np.int64) * V
    t = time.time(); fit = g.evaluatePopulation((ids, off), True, True); cold = time.time() - t
    g.kernel_time(0, reset=True)
    t = time.time()
    for _ in range(5): fit2 = g.evaluatePopulation((ids, off), True, True)
    warm = (time.time() - t) / 5
    ms, n = g.kernel_time(0)
    words = (T + 31) // 32
    bplan = (V - 1) * 4 * words + 4 * (V + 1) + 32
    per = ms / n * 1e-3
    print("pop %7d: cold %.2fs, warm call %.3f ms (%.2f M plans/s e2e), kernel %.1f us -> %.2f M plans/s, %.0f GB/s (%.1f%% of 6548), memo %d" % (
        pop, cold, warm * 1e3, pop / warm / 1e6, per * 1e6, pop / per / 1e6, pop * bplan / per / 1e9, 100 * pop * bplan / per / 1e9 / 6548.2, g.memo_size()), flush=True)
    assert np.array_equal(fit, fit2)
