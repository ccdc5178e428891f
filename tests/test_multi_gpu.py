"""Two GPUs, one process each: the population is sharded by individual, the edges of the whole population are rendered once
across the ranks and their bitmap rows exchanged (NCCL), the fitness rows all-gathered.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import mesh_path, random_plans, specs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import ctypes
    try:
        L = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    return n.value if L.cuInit(0) == 0 and L.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("share", ["1", "0"])
def test_two_ranks_share_the_edge_rendering(tmp_path, share):
    from inspection_path_planning_b200 import cpp_binding
    mesh, height, pop = "sphere", 64, 60
    env = dict(os.environ, IPP_SHARE_EDGES=share)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_mgpu_worker.py"), str(r), "2", str(tmp_path), mesh, str(height), str(pop)],
                              env=env) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    res = [np.load(os.path.join(tmp_path, "rank%d.npz" % r)) for r in range(2)]
    g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(height))
    try:
        rng = np.random.default_rng(71)
        plans = random_plans(rng, g.getNumberOfBoxes(), pop, (2, 9)) + [[]]
        for name, limit_off in (("nolimit", True), ("limit", False)):
            g.clear_memo()
            g.counters(reset=True)
            want = g.evaluatePopulation(plans, True, limit_off)
            distinct = g.counters()["edges_rendered"]
            for r in range(2):
                assert np.array_equal(res[r][name], want), (name, r)  # every rank holds every row, bit for bit
            rendered = [int(res[r][name + "_edges_rendered"]) for r in range(2)]
            if share == "1":
                # each edge rendered exactly once across the ranks, both memos complete
                assert sum(rendered) == distinct and abs(rendered[0] - rendered[1]) <= 1
                assert all(int(res[r][name + "_memo"]) == distinct for r in range(2))
            else:
                assert all(0 < x <= distinct for x in rendered) or distinct == 0
        # [id, angleOffset] genes, the error path and a row pool that is far too small: same rows as one GPU with room for everything
        g.clear_memo()
        ang = [[[gene[0], 0.25 * ((gene[0] % 5) - 2)] for gene in p] for p in plans]
        want_ang = g.evaluatePopulation(ang, True, True)
        g.clear_memo()
        want = g.evaluatePopulation(plans, True, True)
        for r in range(2):
            assert np.array_equal(res[r]["angles"], want_ang), r
            assert int(res[r]["badid_raised"]) == 1, r
            assert np.array_equal(res[r]["after_badid"], want), r
            assert np.array_equal(res[r]["smallpool"], want) and np.array_equal(res[r]["smallpool_again"], want), r
            assert int(res[r]["smallpool_evictions"]) > 0
    finally:
        g.close()
