"""BASELINE.json configs other than the bench workload, on the CUDA evaluator (timings only; parity is in tests/).

    python tools/run_configs.py c1        cube 12 288 triangles, NSGA-II pop 100 x 20 generations (optimize_coverage.py flow)
    python tools/run_configs.py c1cpu     the same run with the CPU oracle as the evaluator (the reference-side leg of config 1)
    python tools/run_configs.py c3 [pop]  luis4.obj (stand-in for oil_manifold.stl), pop x 30-viewpoint plans, cold + warm
    python tools/run_configs.py c4 [pop]  luis1..5 one after the other (multiRun.py style), pop plans of 2..15 viewpoints each, setup included
    python tools/run_configs.py c5 [lvl]  displaced icosphere (20 * 4^lvl triangles), 1 000 x 50-viewpoint plans (BVH-walk variant above 16 384 leaves)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inspection_path_planning_b200 import cpp_binding, optimize_coverage as oc, synth_meshes  # noqa: E402

MESHES = os.path.join(ROOT, "tests", "golden", "meshes")
SPECS = [2.0, 46, 46, 1024, 0.1, 10, 1, 1]


def population(rng, n_boxes, pop, vmin, vmax):
    plans = []
    for _ in range(pop):
        L = int(rng.integers(vmin, vmax + 1))
        ids = rng.integers(0, n_boxes, size=L)
        for k in range(1, L):
            while ids[k] == ids[k - 1]:
                ids[k] = rng.integers(0, n_boxes)
        plans.append(ids)
    offsets = np.zeros(pop + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(p) for p in plans])
    return np.concatenate(plans).astype(np.int32), offsets


def timed_population(g, ids, off, label):
    out = {}
    for phase in ("cold", "warm"):
        g.counters(reset=True)
        t = time.perf_counter()
        fit = g.evaluatePopulation((ids, off), True, True)
        dt = time.perf_counter() - t
        c = g.counters()
        out[phase] = {"s": dt, "plans_per_s": (len(off) - 1) / dt, "edges_rendered": c["edges_rendered"], "snapshots": c["snapshots"],
                      "rays": c["rays"], "mean_coverage": float((1 - fit[:, 0]).mean())}
    print(json.dumps({"config": label, "variant": g.visibility_variant(), "triangles": g.T, "viewpoints": g.getNumberOfBoxes(),
                      "plans": len(off) - 1, **out}), flush=True)


def c1():
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "cube.ippm")
        synth_meshes.write_ippm(path, synth_meshes.cube())
        params = oc.Parameters(NUM_INDIVS=100, NUM_GENERATIONS=20)
        t = time.perf_counter()
        g = oc.make_evaluator(path, params)
        setup = time.perf_counter() - t
        pop, pf, logbook, stats = oc.run_nsga2(g, params, seed=20261019)
        print(json.dumps({"config": "c1 cube 12288 triangles, NSGA-II pop 100 x 20 generations, seeding on, memoisation on", "setup_s": setup,
                          "total_s": stats["total_s"], "evaluator_s": stats["eval_s"], "evaluations": stats["evaluations"],
                          "batched_calls": stats["calls"], "pareto_front": len(pf),
                          "hypervolume": oc.hypervolume_2d([i.fitness for i in pf], params.HYPERVOLUME_REF_POINT),
                          "best_score": logbook[-1]["fit_min"][0], "memoised_edges": stats["memoised_edges"]}), flush=True)
        g.close()


def c1cpu():
    """config 1 as the reference ran it: the same seeded NSGA-II loop with the CPU restatement of the reference evaluator (oracle/,
    one thread, like the reference's serial toolbox.map) -- the CPU leg beside c1"""
    from oracle import oracle as om
    om.build()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "cube.ippm")
        synth_meshes.write_ippm(path, synth_meshes.cube())
        params = oc.Parameters(NUM_INDIVS=100, NUM_GENERATIONS=20)
        t = time.perf_counter()
        o = om.Oracle(path, params.SENSOR_PARAMETERS)
        setup = time.perf_counter() - t
        pop, pf, logbook, stats = oc.run_nsga2(o, params, seed=20261019)
        print(json.dumps({"config": "c1 cube 12288 triangles, NSGA-II pop 100 x 20 generations, CPU restatement of the reference evaluator, 1 thread",
                          "setup_s": setup, "total_s": stats["total_s"], "evaluator_s": stats["eval_s"], "evaluations": stats["evaluations"],
                          "batched_calls": stats["calls"], "pareto_front": len(pf),
                          "hypervolume": oc.hypervolume_2d([i.fitness for i in pf], params.HYPERVOLUME_REF_POINT),
                          "best_score": logbook[-1]["fit_min"][0], "memoised_edges": stats["memoised_edges"]}), flush=True)


def c3(pop=2000):
    g = cpp_binding.PlanCoverageEstimator(os.path.join(MESHES, "luis4.ippm"), SPECS)
    ids, off = population(np.random.default_rng(20261019), g.getNumberOfBoxes(), pop, 30, 30)
    timed_population(g, ids, off, "c3 luis4 (stand-in for oil_manifold), %d plans x 30 viewpoints" % pop)
    g.close()


def c4(pop=1000):
    t0 = time.perf_counter()
    for name in ("luis1", "luis2", "luis3", "luis4", "luis5"):
        t = time.perf_counter()
        g = cpp_binding.PlanCoverageEstimator(os.path.join(MESHES, name + ".ippm"), SPECS)
        setup = time.perf_counter() - t
        ids, off = population(np.random.default_rng(20261019), g.getNumberOfBoxes(), pop, 2, 15)
        timed_population(g, ids, off, "c4 %s, %d plans of 2..15 viewpoints (setup %.3f s)" % (name, pop, setup))
        g.close()
    print(json.dumps({"config": "c4 total wall", "s": time.perf_counter() - t0}), flush=True)


def c5(level=6, pop=200):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ico.ippm")
        synth_meshes.write_ippm(path, synth_meshes.icosphere(level=level))
        t = time.perf_counter()
        g = cpp_binding.PlanCoverageEstimator(path, SPECS)
        setup = time.perf_counter() - t
        ids, off = population(np.random.default_rng(20261019), g.getNumberOfBoxes(), pop, 50, 50)
        timed_population(g, ids, off, "c5 icosphere level %d, %d plans x 50 viewpoints (setup %.3f s)" % (level, pop, setup))
        g.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c1"
    arg = [int(x) for x in sys.argv[2:]]
    {"c1": c1, "c1cpu": c1cpu, "c3": c3, "c4": c4, "c5": c5}[which](*arg)
