"""Where does the GPU see triangle TRI on edge PREV->CURR, and what does the oracle see at those pixels?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import mesh_path, specs
from inspection_path_planning_b200 import cpp_binding
from oracle.oracle import Oracle
mesh, H, prev, curr, tri = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
o = Oracle(mesh_path(mesh), specs(H))
c = o.box_centres()
a, b = c[prev], c[curr]
pos = np.asarray(o.sampling_positions(a, b), dtype=np.float32).reshape(-1, 3)
hd = np.asarray(o.heading(a, b))
V = np.asarray(o.vertices(), dtype=np.float64).reshape(-1, 3, 3)
tanv = np.tan(np.radians(23.0))
def tri_t(eye, f, s, u, X, Y, k):
    A, B, Cc = (V[k] - eye)
    D = f + X * s + Y * u
    Na, Nb, Nc = np.cross(B, Cc - B), np.cross(Cc, A - Cc), np.cross(A, B - A)
    Ng = np.cross(B - A, Cc - A)
    e = np.array([D @ Na, D @ Nb, D @ Nc]); det = D @ Ng; vol = A @ Ng
    return e, det, vol, (vol / det if det != 0 else np.inf)
for p in pos:
    for cam in ("front", "down"):
        eye = p.astype(np.float64)
        if cam == "front":
            center = (p + hd.astype(np.float32)).astype(np.float64); up = np.array([0.0, 0.0, 1.0])
        else:
            center = (p - np.array([0, 0, 1], dtype=np.float32)).astype(np.float64); up = hd
        gi, oi = g.render_ids(eye, center, up), o.render_ids(eye, center, up)
        ys, xs = np.nonzero(gi == tri)
        oys, oxs = np.nonzero(oi == tri)
        if len(ys) == 0 and len(oys) == 0: continue
        print(cam, p, "gpu pixels", len(ys), "oracle pixels", len(oys), "image diffs", int((gi != oi).sum()))
        f = center - eye; f /= np.linalg.norm(f); s = np.cross(f, up); s /= np.linalg.norm(s); u = np.cross(s, f); u /= np.linalg.norm(u)
        for y, x in list(zip(ys, xs))[:6]:
            X = ((x + 0.5) / H * 2 - 1) * tanv; Y = ((y + 0.5) / H * 2 - 1) * tanv
            other = int(oi[y, x])
            print("  pixel", x, y, "oracle id", other)
            for k in (tri, other):
                if k < 0: continue
                e, det, vol, t = tri_t(eye, f, s, u, X, Y, k)
                n = np.linalg.norm(np.cross(V[k][1] - V[k][0], V[k][2] - V[k][0]))
                print("    tri %d: e/|e|max %s det %.3e vol %.3e t %.9f  cos(view,normal) %.2e" % (k, (e / np.abs(e).max()).round(9), det, vol, t, det / (n * np.linalg.norm(f + X * s + Y * u))))
