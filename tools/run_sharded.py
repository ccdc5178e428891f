"""One population evaluated by all ranks together (strong scaling): cold and warm wall time of the host-pointer sharded call.

    python tools/run_sharded.py <mesh> <pop[,pop,...]> <viewpoints>                                              (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_sharded.py <mesh> <pop[,pop,...]> <viewpoints>                                                 (N GPUs)

<mesh>: a fixture of tests/golden/meshes (luis4, mockup_only, ...) or ico<level> = the displaced icosphere of BASELINE config 5
(ico8: 1 310 720 triangles).  The populations are evaluated one after the other on the same evaluator: `cold` renders whatever
edges the memo does not hold yet, `warm` is the same population again.  IPP_SHARE_EDGES=0 switches the shared rendering off.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inspection_path_planning_b200 import _lib, cpp_binding, synth_meshes  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "luis4"
pops = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2000]
V = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
L = _lib.load()
tmp = None
if mesh.startswith("ico"):
    tmp = tempfile.TemporaryDirectory()
    path = os.path.join(tmp.name, "%s_%d.ippm" % (mesh, rank))
    synth_meshes.write_ippm(path, synth_meshes.icosphere(level=int(mesh[3:])))
else:
    path = os.path.join(ROOT, "tests", "golden", "meshes", mesh + ".ippm")
t0 = time.perf_counter()
g = cpp_binding.PlanCoverageEstimator(path, [2.0, 46, 46, 1024, 0.1, 10, 1, 1], device=local)
setup_s = time.perf_counter() - t0
if world > 1:
    import torch.distributed as dist  # gloo: only hands the 128-byte NCCL id to the ranks
    dist.init_process_group("gloo")
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _lib.check(L.ipp_comm_unique_id(idbuf.ctypes.data_as(_lib.c_u8p)))
    obj = [idbuf.tobytes()]
    dist.broadcast_object_list(obj, src=0)
    idbuf = np.frombuffer(obj[0], dtype=np.uint8).copy()
    saved = os.dup(1); os.dup2(2, 1)
    _lib.check(L.ipp_comm_init(g._h, rank, world, idbuf.ctypes.data_as(_lib.c_u8p)))
    _lib.check(L.ipp_comm_barrier(g._h))
    os.dup2(saved, 1); os.close(saved)


def barrier():
    if world > 1:
        _lib.check(L.ipp_comm_barrier(g._h))
    g.sync()


def allsum(x):
    a = np.asarray([x], dtype=np.float64)
    if world > 1:
        _lib.check(L.ipp_comm_allreduce_sum(g._h, a.ctypes.data_as(_lib.c_f64p), 1))
    return float(a[0])


N = g.getNumberOfBoxes()
words = (g.T + 31) // 32
b_plan = (V - 1) * 4 * words + 4 * (V + 1) + 32
for pop in pops:
    rng = np.random.default_rng(20261019 + pop)
    ids = rng.integers(0, N, size=(pop, V), dtype=np.int32)
    for k in range(1, V):
        dup = ids[:, k] == ids[:, k - 1]
        while dup.any():
            ids[dup, k] = rng.integers(0, N, size=int(dup.sum()), dtype=np.int32)
            dup = ids[:, k] == ids[:, k - 1]
    plans = (np.ascontiguousarray(ids.reshape(-1)), np.arange(pop + 1, dtype=np.int64) * V)
    out = {}
    for phase in ("cold", "warm"):
        barrier()
        g.counters(reset=True)
        g.kernel_time(0, reset=True)
        t = time.perf_counter()
        fit = g.evaluatePopulationSharded(plans, True, True) if world > 1 else g.evaluatePopulation(plans, True, True)
        barrier()
        dt = time.perf_counter() - t
        c = g.counters()
        ms, n = g.kernel_time(0)
        out[phase] = {"s": dt, "plans_per_s": pop / dt, "edges_rendered_all_ranks": int(allsum(c["edges_rendered"])),
                      "poses_all_ranks": int(allsum(c["snapshots"])), "poses_rendered_all_ranks": int(allsum(c["snapshots_rendered"])),
                      "plan_reduce_ms_this_rank": ms, "plan_reduce_gbs_this_rank": ((pop + world - 1) // world) * b_plan / (ms * 1e-3) / 1e9 if ms > 0 else None}
    if rank == 0:
        print(json.dumps({"mesh": mesh, "triangles": g.T, "candidate_viewpoints": N, "variant": g.visibility_variant(), "plans": pop,
                          "viewpoints": V, "world": world, "share_edges": os.environ.get("IPP_SHARE_EDGES", "1"), "setup_s": setup_s,
                          "memo_rows": g.memo_size(), "evictions": g.evictions(), "mean_coverage": float((1 - fit[:, 0]).mean()), **out}),
              flush=True)
g.close()
