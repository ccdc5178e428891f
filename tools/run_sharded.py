"""One population sharded over the ranks (torchrun, one rank per GPU): cold and warm wall time of evaluatePopulationSharded with and
without the shared edge rendering (IPP_SHARE_EDGES=0/1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_sharded.py luis4 2000 30
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inspection_path_planning_b200 import _lib, cpp_binding  # noqa: E402

mesh = sys.argv[1] if len(sys.argv) > 1 else "luis4"
pop = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
V = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
L = _lib.load()
g = cpp_binding.PlanCoverageEstimator(os.path.join(ROOT, "tests", "golden", "meshes", mesh + ".ippm"), [2.0, 46, 46, 1024, 0.1, 10, 1, 1], device=local)
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
rng = np.random.default_rng(20261019)
N = g.getNumberOfBoxes()
ids = rng.integers(0, N, size=(pop, V))
for k in range(1, V):
    dup = ids[:, k] == ids[:, k - 1]
    while dup.any():
        ids[dup, k] = rng.integers(0, N, size=int(dup.sum()))
        dup = ids[:, k] == ids[:, k - 1]
plans = (np.ascontiguousarray(ids.reshape(-1).astype(np.int32)), np.arange(pop + 1, dtype=np.int64) * V)
out = {}
for phase in ("cold", "warm"):
    _lib.check(L.ipp_comm_barrier(g._h))
    g.counters(reset=True)
    t = time.perf_counter()
    fit = g.evaluatePopulationSharded(plans, True, True)
    _lib.check(L.ipp_comm_barrier(g._h))
    out[phase] = {"s": time.perf_counter() - t, "edges_rendered_by_this_rank": g.counters()["edges_rendered"]}
if rank == 0:
    print(json.dumps({"mesh": mesh, "plans": pop, "viewpoints": V, "world": world, "share_edges": os.environ.get("IPP_SHARE_EDGES", "1"),
                      "memo": g.memo_size(), "mean_coverage": float((1 - fit[:, 0]).mean()), **out}), flush=True)
g.close()
