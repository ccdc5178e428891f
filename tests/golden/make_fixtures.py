"""Regenerates the mesh fixtures under tests/golden/meshes/ from the reference's shipped assets.

Run in the build container only (needs /root/reference):  python tests/golden/make_fixtures.py

The fixtures are the reference's 3d_models/*.stl|obj after mesh ingest (file-order triangles, float32, OBJ
rotated Y-up -> Z-up) in the flat .ippm format both the oracle and libipp.so read, so that the GPU box --
which has no /root/reference -- evaluates exactly the meshes BASELINE.json's configs name
(mockup_only.stl for config 2, luis4.obj as the stand-in for the missing oil_manifold.stl in config 3).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import Oracle, write_ippm  # noqa: E402

REF = "/root/reference/3d_models"
ASSETS = {
    "mockup_only.ippm": "mockup_only.stl",
    "luis4.ippm": "luis4.obj",
    "luis1.ippm": "luis1.obj",
    "luis2.ippm": "luis2.obj",
    "luis3.ippm": "luis3.obj",
    "luis5.ippm": "luis5.obj",
    "sphere.ippm": "simple_shapes/sphere.obj",
}

if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
    for out, src in ASSETS.items():
        # tiny image: only the loader is exercised here
        o = Oracle(os.path.join(REF, src), [2.0, 46, 46, 8, 0.1, 10, 1, 1], post_processing=True)
        v = o.vertices()
        write_ippm(os.path.join(HERE, "meshes", out), v)
        print(out, v.shape[0], "triangles")
