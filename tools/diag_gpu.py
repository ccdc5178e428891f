"""Quick GPU diagnostics (not a benchmark): setup time, cold/warm population timing, kernel-group times."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, random_plans, specs  # noqa: E402
from inspection_path_planning_b200 import cpp_binding  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "mockup_only"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pop = int(sys.argv[3]) if len(sys.argv) > 3 else 200
V = int(sys.argv[4]) if len(sys.argv) > 4 else 20

path = mesh_path(mesh)
if mesh.startswith("ico"):  # the displaced icosphere of BASELINE config 5 (ico8: 1 310 720 triangles)
    import tempfile
    from inspection_path_planning_b200 import synth_meshes
    path = os.path.join(tempfile.mkdtemp(), mesh + ".ippm")
    synth_meshes.write_ippm(path, synth_meshes.icosphere(level=int(mesh[3:])))
t = time.time()
g = cpp_binding.PlanCoverageEstimator(path, specs(H))
print("setup %.3fs T=%d N=%d maxE=%.6f" % (time.time() - t, g.T, g.getNumberOfBoxes(), g.getMaxAllowedEnergy()), flush=True)
rng = np.random.default_rng(20261019)
plans = cpp_binding.pack_population(random_plans(rng, g.getNumberOfBoxes(), pop, V))
names = ["reduce", "visibility", "collision", "pack", "plan", "table"]
for it in range(3):
    g.counters(reset=True)
    g.visibility_stats(reset=True)
    for k in range(6):
        g.kernel_time(k, reset=True)
    t = time.time()
    fit = g.evaluatePopulation(plans, True, True)
    dt = time.time() - t
    c = g.counters()
    vs = g.visibility_stats()
    if vs["tiles"]:
        print("   per tile: nodes %.1f triangles %.1f (tiles %d)" % (vs["nodes"] / vs["tiles"], vs["triangles"] / vs["tiles"], vs["tiles"]), flush=True)
        if vs["live_triangles"]:  # library built with -DIPP_VIS_DIAG (tools/build_variant.py)
            print("   per tile: records staged %.1f, live after the tile test %.1f, 8x4 block tests %.1f, %.1f %% of the tiles stay open" % (
                vs["supertile_splits"] / vs["tiles"], vs["live_triangles"] / vs["tiles"], vs["supertiles"] / vs["tiles"],
                100.0 * vs["open_tiles"] / vs["tiles"]), flush=True)
    kt = {names[k]: g.kernel_time(k) for k in range(6)}
    import hashlib
    print("iter %d: %.4fs  %.1f plans/s  cov mean %.4f  fitness sha1 %s  counters %s" % (
        it, dt, pop / dt, (1 - fit[:, 0]).mean(), hashlib.sha1(np.ascontiguousarray(fit).tobytes()).hexdigest()[:12], c), flush=True)
    print("   kernel ms:", {k: (round(v[0], 3), v[1]) for k, v in kt.items()}, flush=True)
    if c["rays"]:
        print("   rays/s (visibility kernel) %.3e" % (c["rays"] / (kt["visibility"][0] * 1e-3)), flush=True)
print("memo", g.memo_size())
