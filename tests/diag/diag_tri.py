"""Pixel-level comparison of GPU and oracle item buffers for the poses of one edge (debug aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bit_ids, mesh_path, specs
from inspection_path_planning_b200 import cpp_binding
from oracle.oracle import Oracle
mesh, H, prev, curr = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(H))
o = Oracle(mesh_path(mesh), specs(H))
got = g.edge_bitmap(prev, curr)
seen, strict, lenient = o.edge_bitmap(prev, curr, classify=True)
missing = bit_ids(seen & ~got); extra = bit_ids(got & ~seen)
print("missing", missing.tolist(), "extra", extra.tolist())
c = o.box_centres()
a, b = c[prev], c[curr]
hd = o.heading(a, b)
V = o.vertices().astype(np.float64)
area = o.areas()
for k, p in enumerate(o.sampling_positions(a, b)):
    p = p.astype(np.float64)
    hf = hd.astype(np.float32).astype(np.float64)
    for cam, (center, up) in enumerate([((p.astype(np.float32) + hf.astype(np.float32)).astype(np.float64), [0, 0, 1]), ((p.astype(np.float32) - np.array([0, 0, 1], np.float32)).astype(np.float64), hd)]):
        io = o.render_ids(p, center, up)
        ig = g.render_ids(p, center, up)
        nd = int((io != ig).sum())
        for t in list(missing) + list(extra):
            po, pg = int((io == t).sum()), int((ig == t).sum())
            if po or pg:
                ys, xs = np.nonzero((io == t) | (ig == t))
                print("sample %d cam %d: tri %d oracle px %d gpu px %d area %.4g  at x %d..%d y %d..%d | total differing px %d" % (k, cam, t, po, pg, area[t], xs.min(), xs.max(), ys.min(), ys.max(), nd))
                if pg and not po:
                    y, x = np.nonzero(ig == t); y, x = y[0], x[0]
                    other = int(io[y, x])
                    print("    gpu pixel (%d,%d): oracle has id %d there; tile (%d,%d) block (%d,%d)" % (x, y, other, x // 32, y // 32, (x % 32) // 8, (y % 32) // 4))
                    f = center - p; f /= np.linalg.norm(f)
                    sv = np.cross(f, np.asarray(up, float)); sv /= np.linalg.norm(sv); uv = np.cross(sv, f)
                    tan = np.tan(np.radians(23.0))
                    X = ((x + 0.5) / H * 2 - 1) * tan; Y = ((y + 0.5) / H * 2 - 1) * tan
                    D = f + sv * X + uv * Y
                    for tt in (t, other):
                        if tt >= 0:
                            A, B, C = (V[tt] - p)
                            es = [D @ np.cross(B, C - B), D @ np.cross(C, A - C), D @ np.cross(A, B - A)]
                            n = np.cross(B - A, C - A)
                            tdepth = (A @ n) / (D @ n)
                            sc = np.linalg.norm(A) * max(np.linalg.norm(B - A), np.linalg.norm(C - A), np.linalg.norm(C - B))
                            print("       tri %d depth %.7f edge fn / scale %s vertex depths %s cos %.4g" % (tt, tdepth, [e / sc for e in es], (V[tt] - p) @ f, (n @ D) / np.linalg.norm(n) / np.linalg.norm(D)))
                if po and not pg:
                    y, x = np.nonzero(io == t); y, x = y[0], x[0]
                    print("    oracle pixel (%d,%d): gpu has id %d there; tile (%d,%d) block (%d,%d)" % (x, y, ig[y, x], x // 32, y // 32, (x % 32) // 8, (y % 32) // 4))
                    f = center - p; f /= np.linalg.norm(f)
                    for tt in (t, int(ig[y, x])):
                        if tt >= 0:
                            z = (V[tt] - p) @ f
                            n = np.cross(V[tt][1] - V[tt][0], V[tt][2] - V[tt][0])
                            print("       tri %d vertex depths %s  normal.f/|n| %.4g" % (tt, z, (n @ f) / np.linalg.norm(n)))
