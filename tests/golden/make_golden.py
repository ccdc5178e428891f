"""Regenerates tests/golden/golden.json: outputs of the CPU oracle on the mesh fixtures (regression vectors; the reference
itself ships none -- parity unpinned).   python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from conftest import mesh_path, random_plans, specs  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = [("sphere", 64, 8), ("luis1", 96, 4), ("mockup_only", 128, 4), ("luis4", 64, 3)]


def golden_case(mesh, height, n_plans):
    o = Oracle(mesh_path(mesh), specs(height))
    rng = np.random.default_rng(99)
    plans = random_plans(rng, o.getNumberOfBoxes(), n_plans, (2, 5))
    fit = o.evaluatePopulation(plans, True, True)
    edges = [(int(p[0][0]), int(p[1][0])) for p in plans]
    bitmaps = [o.edge_bitmap(a, b) for a, b in edges]
    c, bb = o.scene_info()
    return {
        "mesh": mesh, "height": height, "triangles": o.T, "boxes": o.getNumberOfBoxes(),
        "max_allowed_energy": o.getMaxAllowedEnergy(), "total_area": o.total_area(), "collision_penalty": o.collision_penalty(),
        "scene_center": c.tolist(), "bbox": bb.tolist(),
        "grid_sha1": hashlib.sha1(o.grid().tobytes()).hexdigest(), "grid_shape": list(o.grid().shape),
        "simple_plans": [[g[0] for g in p] for p in o.getSimplePlans()],
        "plans": [[g[0] for g in p] for p in plans], "fitness": fit.tolist(),
        "edges": edges, "edge_popcounts": [int(np.unpackbits(b.view(np.uint8)).sum()) for b in bitmaps],
        "edge_sha1": [hashlib.sha1(b.tobytes()).hexdigest() for b in bitmaps],
    }


if __name__ == "__main__":
    out = [golden_case(*c) for c in CASES]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    for g in out:
        print(g["mesh"], g["triangles"], g["boxes"], g["max_allowed_energy"], g["edge_popcounts"])
