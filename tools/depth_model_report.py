"""How much does the choice of depth model change the seen sets?  The oracle's default (planar depth, ties within a relative 2^-16,
lowest ID) against an OpenGL-style depth buffer (24-bit window depth, GL_LESS in ID order, top-left fill rule) -- what the
reference's viewer most likely did (parity is unpinned: the GL driver is not in the reference tree).

    python tools/depth_model_report.py [edges per mesh] [image height]      (CPU only; oracle in worker processes)
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _work(task):
    mesh, height, a, b = task
    from oracle.oracle import Oracle
    o = Oracle(os.path.join(ROOT, "tests", "golden", "meshes", mesh + ".ippm"), [2.0, 46, 46, height, 0.1, 10, 1, 1])
    s0 = o.edge_bitmap(a, b)
    o.set_depth_model(1)
    s1 = o.edge_bitmap(a, b)
    pc = lambda w: int(np.unpackbits(w.view(np.uint8)).sum())
    area = o.areas()
    bits = np.unpackbits((s0 ^ s1).view(np.uint8), bitorder="little")[:o.T].astype(bool)
    return {"edge": [a, b], "seen_default": pc(s0), "seen_gl24": pc(s1), "only_default": pc(s0 & ~s1), "only_gl24": pc(s1 & ~s0),
            "triangles": o.T, "coverage_difference": float(area[bits].sum() / o.total_area())}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    height = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    from oracle.oracle import Oracle
    out = {}
    for mesh in ("mockup_only", "luis4"):
        o = Oracle(os.path.join(ROOT, "tests", "golden", "meshes", mesh + ".ippm"), [2.0, 46, 46, 8, 0.1, 10, 1, 1])
        N = o.getNumberOfBoxes()
        c = o.box_centres()
        rng = np.random.default_rng(5)
        edges = []
        while len(edges) < n:
            a, b = (int(x) for x in rng.choice(N, 2, replace=False))
            if np.linalg.norm(c[a] - c[b]) <= 12.0:
                edges.append((a, b))
        with mp.get_context("spawn").Pool(min(len(edges), os.cpu_count() or 1)) as pool:
            res = pool.map(_work, [(mesh, height, a, b) for a, b in edges], chunksize=1)
        seen = sum(r["seen_default"] for r in res)
        diff = sum(r["only_default"] + r["only_gl24"] for r in res)
        out[mesh] = {"image_height": height, "edges": len(res), "seen_bits_default_model": seen, "bits_that_differ": diff,
                     "fraction_of_seen_bits": diff / max(1, seen), "worst_edge_flips": max(r["only_default"] + r["only_gl24"] for r in res),
                     "flip_budget_per_edge": int(5e-4 * res[0]["triangles"]), "worst_coverage_difference": max(r["coverage_difference"] for r in res),
                     "per_edge": res}
        print(mesh, {k: v for k, v in out[mesh].items() if k != "per_edge"}, flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_depth_model_report.json"), "w"), indent=1)
