"""plan_reduce (K6) streaming test: warm re-evaluation of a large population over a pool of `nedges` memoised edges, every variant
of the streaming kernel on the same population; kernel time from CUDA events, L2 flushed before every launch.

    python tools/diag_reduce.py <mesh> <image height> <pop,pop,...> <viewpoints> <nedges> [variants, e.g. 0,1,2]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs
from inspection_path_planning_b200 import _lib, cpp_binding
mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pops = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1000, 10000, 100000]
V = int(sys.argv[4]) if len(sys.argv) > 4 else 20
nedges = int(sys.argv[5]) if len(sys.argv) > 5 else 20000
variants = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0]
path = mesh_path(mesh)
if mesh.startswith("ico"):  # the displaced icosphere of BASELINE config 5 (ico8: 1 310 720 triangles)
    import tempfile
    from inspection_path_planning_b200 import synth_meshes
    path = os.path.join(tempfile.mkdtemp(), mesh + ".ippm")
    synth_meshes.write_ippm(path, synth_meshes.icosphere(level=int(mesh[3:])))
g = cpp_binding.PlanCoverageEstimator(path, specs(H))
L = _lib.load()
N = g.getNumberOfBoxes(); T = g.T
rng = np.random.default_rng(1)
# plans are random walks over a random subgraph of `nedges` directed edges, so that the cold render stays short
keys = rng.choice(N * N, size=min(nedges, N * N), replace=False)
a, b = keys // N, keys % N
ok = a != b
a, b = a[ok], b[ok]
order = np.argsort(a, kind="stable")
a, b = a[order], b[order]
start = np.searchsorted(a, np.arange(N + 1))
deg = np.diff(start)
have = np.nonzero(deg > 0)[0]


def walks(pop, Lw):
    """vectorised: every step picks a random successor; walkers that reach a dead end restart from a fresh source"""
    out = np.empty((pop, Lw), dtype=np.int64)
    cur = have[rng.integers(0, len(have), pop)]
    out[:, 0] = cur
    for k in range(1, Lw):
        d = deg[cur]
        dead = d == 0
        nxt = b[np.minimum(start[cur] + (rng.random(pop) * np.maximum(d, 1)).astype(np.int64), len(b) - 1)]
        # a dead end: jump along any pooled edge out of a fresh source (the jump itself is then not a pooled edge: rendered once)
        nxt[dead] = have[rng.integers(0, len(have), int(dead.sum()))]
        same = nxt == cur
        nxt[same] = have[rng.integers(0, len(have), int(same.sum()))]
        out[:, k] = nxt
        cur = nxt
    return out


res = []
for pop in pops:
    ids = walks(pop, V).astype(np.int32).reshape(-1)
    off = np.arange(pop + 1, dtype=np.int64) * V
    t = time.time(); fit = g.evaluatePopulation((ids, off), True, True); cold = time.time() - t
    words = (T + 31) // 32
    bplan = (V - 1) * 4 * words + 4 * (V + 1) + 32
    for v in variants:
        name = L.ipp_plan_reduce_variant(v).decode()
        fit2 = g.evaluatePopulation((ids, off), True, True)
        assert np.array_equal(fit, fit2), ("variant %d differs" % v, np.abs(fit - fit2).max())
        g.kernel_time(0, reset=True)
        for _ in range(5):
            _lib.check(L.ipp_flush_l2(g._h))
            g.evaluatePopulation((ids, off), True, True)
        ms, n = g.kernel_time(0)
        per = ms / n * 1e-3
        r = {"mesh": mesh, "pop": pop, "viewpoints": V, "variant": v, "name": name, "us_per_launch": per * 1e6, "plans_per_s": pop / per,
             "bytes_per_plan": bplan, "gbs": pop * bplan / per / 1e9, "frac_of_6548": pop * bplan / per / 1e9 / 6548.2, "memo": g.memo_size(),
             "pool_mb": g.memo_size() * ((words + 3) // 4 * 16) / 1e6}
        res.append(r)
        print("pop %7d variant %d (%s): kernel %.1f us -> %.2f M plans/s, %.0f GB/s (%.1f%% of 6548), memo %d rows (%.0f MB), cold %.2fs" % (
            pop, v, name, per * 1e6, pop / per / 1e6, r["gbs"], 100 * r["frac_of_6548"], r["memo"], r["pool_mb"], cold), flush=True)
    L.ipp_plan_reduce_variant(-1)
if os.environ.get("IPP_DIAG_JSON"):
    json.dump(res, open(os.environ["IPP_DIAG_JSON"], "w"), indent=1)
