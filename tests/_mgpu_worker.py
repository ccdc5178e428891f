"""Worker of tests/test_multi_gpu.py: one process per GPU, NCCL id handed over through a file."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, random_plans, specs  # noqa: E402
from inspection_path_planning_b200 import _lib, cpp_binding  # noqa: E402

rank, world, work = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
mesh, height, pop = sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
L = _lib.load()
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(height), device=rank)
idfile = os.path.join(work, "nccl_id.bin")
idbuf = np.zeros(128, dtype=np.uint8)
if rank == 0:
    _lib.check(L.ipp_comm_unique_id(idbuf.ctypes.data_as(_lib.c_u8p)))
    with open(idfile + ".tmp", "wb") as f:
        f.write(idbuf.tobytes())
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id")
        time.sleep(0.05)
    idbuf = np.frombuffer(open(idfile, "rb").read(), dtype=np.uint8).copy()
_lib.check(L.ipp_comm_init(g._h, rank, world, idbuf.ctypes.data_as(_lib.c_u8p)))
rng = np.random.default_rng(71)
plans = random_plans(rng, g.getNumberOfBoxes(), pop, (2, 9)) + [[]]
out = {}
for name, limit_off in (("nolimit", True), ("limit", False)):
    g.clear_memo()
    g.counters(reset=True)
    fit = g.evaluatePopulationSharded(plans, True, limit_off)
    c = g.counters()
    out[name] = fit
    out[name + "_edges_rendered"] = c["edges_rendered"]
    out[name + "_memo"] = g.memo_size()
    fit2 = g.evaluatePopulationSharded(plans, True, limit_off)  # warm: nothing new, same rows
    assert np.array_equal(fit, fit2)
    assert g.counters()["edges_rendered"] == c["edges_rendered"]

# ---- [id, angleOffset] genes through the sharded call (every rank renders what its own shard needs; their memo is keyed on the host) ----
g.clear_memo()
ang = [[[gene[0], 0.25 * ((gene[0] % 5) - 2)] for gene in p] for p in plans]
out["angles"] = g.evaluatePopulationSharded(ang, True, True)

# ---- a bad box id in the last rank's shard: every rank raises, nobody waits in a collective, and the evaluators stay usable ----
N = g.getNumberOfBoxes()
bad = [list(p) for p in plans[:-1]] + [[[0], [N + 3], [1]]]
try:
    g.evaluatePopulationSharded(bad, True, True)
    out["badid_raised"] = 0
except (IndexError, RuntimeError):
    out["badid_raised"] = 1
g.clear_memo()
out["after_badid"] = g.evaluatePopulationSharded(plans, True, True)
g.close()

# ---- a row pool far too small for the population: the shared rendering gives way, every rank evaluates its shard in pieces ----
os.environ["IPP_CACHE_GB"] = "0.000004"
g2 = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(height), device=rank)
del os.environ["IPP_CACHE_GB"]
idfile2 = os.path.join(work, "nccl_id2.bin")
idbuf2 = np.zeros(128, dtype=np.uint8)
if rank == 0:
    _lib.check(L.ipp_comm_unique_id(idbuf2.ctypes.data_as(_lib.c_u8p)))
    with open(idfile2 + ".tmp", "wb") as f:
        f.write(idbuf2.tobytes())
    os.replace(idfile2 + ".tmp", idfile2)
else:
    t0 = time.time()
    while not os.path.exists(idfile2):
        if time.time() - t0 > 120:
            raise SystemExit("no NCCL id")
        time.sleep(0.05)
    idbuf2 = np.frombuffer(open(idfile2, "rb").read(), dtype=np.uint8).copy()
_lib.check(L.ipp_comm_init(g2._h, rank, world, idbuf2.ctypes.data_as(_lib.c_u8p)))
out["smallpool"] = g2.evaluatePopulationSharded(plans, True, True)
out["smallpool_again"] = g2.evaluatePopulationSharded(plans, True, True)
out["smallpool_evictions"] = g2.evictions()
g2.close()
np.savez(os.path.join(work, "rank%d.npz" % rank), **out)
