"""Process pool around the CPU oracle for the full-resolution parity tests: one oracle per worker and (mesh, sensor) pair, one
edge per task.  At 1024 x 1024 an edge with its edge-epsilon classification costs the oracle seconds (tens of seconds on the
largest meshes); the pool keeps the GPU suite inside its budget on the box's host cores."""
import multiprocessing as mp
import os

import numpy as np

_ORACLES = {}


def _oracle(mesh_file, sp, kw):
    key = (mesh_file, tuple(sp), tuple(kw))  # kw: sorted (name, value) pairs
    if key not in _ORACLES:
        from oracle.oracle import Oracle
        _ORACLES[key] = Oracle(mesh_file, list(sp), **dict(kw))
    return _ORACLES[key]


def _classify(task):
    mesh_file, sp, kw, prev, curr, angle = task
    o = _oracle(mesh_file, sp, kw)
    seen, strict, lenient = o.edge_bitmap(prev, curr, angle, classify=True)
    return seen, strict, lenient


def workers():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, n)


def classify_edges(mesh_file, sp, edges, angles=None, **kw):
    """[(prev, curr)] -> [(seen, strict, lenient)] in order, computed by the oracle in worker processes."""
    from oracle import oracle as om
    om.build()
    kwt = tuple(sorted(kw.items()))
    tasks = [(mesh_file, tuple(sp), kwt, int(a), int(b), 0.0 if angles is None else float(angles[i])) for i, (a, b) in enumerate(edges)]
    n = min(workers(), len(tasks))
    if n <= 1:
        return [_classify(t) for t in tasks]
    # longest edges first would need their lengths; the pool hands tasks out one at a time, which balances well enough
    with mp.get_context("spawn").Pool(n) as pool:
        return pool.map(_classify, tasks, chunksize=1)
