"""World-size-2 test of the population-sharding host logic on the CPU (gloo): contiguous shards of ceil(pop / world) plans,
per-rank evaluation, all-gather of the fitness rows in rank order -- the layout ipp_evaluate_population_sharded uses with
NCCL on the GPUs (csrc/comm.cu).  The evaluator behind each rank is the CPU oracle here."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, mesh_path, random_plans, specs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pop, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from inspection_path_planning_b200.sharding import shard_bounds, shard_csr
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle(mesh_path("sphere"), specs(32))
    rng = np.random.default_rng(77)
    plans = random_plans(rng, o.getNumberOfBoxes(), pop, (1, 6))
    ids = np.array([g[0] for p in plans for g in p], dtype=np.int32)
    offsets = np.cumsum([0] + [len(p) for p in plans]).astype(np.int64)
    per, lo, hi = shard_bounds(pop, world, rank)
    my_ids, my_off = shard_csr(ids, offsets, lo, hi)
    mine = [[[int(x)] for x in my_ids[my_off[i]:my_off[i + 1]]] for i in range(hi - lo)]
    rows = np.zeros((per, 4))
    if mine:
        rows[:hi - lo] = o.evaluatePopulation(mine, True, True)
    gathered = [torch.zeros(per, 4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(rows))
    full = torch.cat(gathered).numpy()[:pop]
    if rank == 0:
        want = o.evaluatePopulation(plans, True, True)
        np.save(os.path.join(out_dir, "got.npy"), full)
        np.save(os.path.join(out_dir, "want.npy"), want)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("pop", [7, 8, 1])
def test_sharded_evaluation_matches_unsharded(pop, tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, pop, str(tmp_path)), nprocs=2, join=True)
    got, want = np.load(tmp_path / "got.npy"), np.load(tmp_path / "want.npy")
    assert np.array_equal(got, want)


def test_shard_bounds_cover_the_population():
    from inspection_path_planning_b200.sharding import shard_bounds
    for pop in (0, 1, 7, 8, 1000, 100001):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(pop, world, r) for r in range(world)]
            per = spans[0][0]
            assert all(s[0] == per for s in spans) and per * world >= pop
            assert spans[0][1] == 0 and spans[-1][2] == pop
            for a, b in zip(spans, spans[1:]):
                assert a[2] == b[1]
            assert all(0 <= s[2] - s[1] <= per for s in spans)
