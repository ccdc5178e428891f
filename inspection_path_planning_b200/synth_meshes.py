"""Procedural inspection meshes for the configs whose assets the reference does not ship
(BASELINE.md section 4): a subdivided cube (config 1) and a displaced icosphere (config 5)."""
import numpy as np


def write_ippm(path, tris):
    """tris float32 [T, 3, 3] -> flat mesh file ("IPPM", uint32 T, 9 floats per triangle)."""
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"IPPM")
        f.write(np.uint32(tris.shape[0]).tobytes())
        f.write(tris.tobytes())


def write_binary_stl(path, tris):
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    rec = np.zeros(tris.shape[0], dtype=[("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    rec["v"] = tris
    with open(path, "wb") as f:
        f.write(b"ipp synthetic mesh".ljust(80, b" "))
        f.write(np.uint32(tris.shape[0]).tobytes())
        f.write(rec.tobytes())


def write_obj_quads(path, verts, quads):
    """Y-up OBJ (the loader rotates to Z-up) with quad faces, 1-based indices."""
    with open(path, "w") as f:
        f.write("# ipp synthetic mesh\no mesh\n")
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % tuple(v))
        for q in quads:
            f.write("f " + " ".join(str(int(i) + 1) for i in q) + "\n")


def cube(side=2.0, n=32, centre=(0.0, 0.0, 1.0)):
    """Axis-aligned cube, each face an n x n grid of quads split into two triangles: 12 n^2 triangles."""
    h = side / 2.0
    g = np.linspace(-h, h, n + 1)
    tris = []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            u, v = [a for a in range(3) if a != axis]
            for i in range(n):
                for j in range(n):
                    def P(a, b):
                        p = np.zeros(3)
                        p[axis] = sgn * h
                        p[u] = g[a]
                        p[v] = g[b]
                        return p + np.asarray(centre)
                    p00, p10, p11, p01 = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
                    tris.append([p00, p10, p11])
                    tris.append([p00, p11, p01])
    return np.asarray(tris, dtype=np.float32)


def icosphere(level=3, radius=10.0, centre=(0.0, 0.0, 10.0), bump=0.05):
    """Icosphere with 20 * 4^level triangles and radial displacement r (1 + bump sin 7 theta sin 5 phi)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    tri = v[f]  # [20, 3, 3]
    for _ in range(level):
        a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
        ab, bc, ca = a + b, b + c, c + a
        ab /= np.linalg.norm(ab, axis=1, keepdims=True)
        bc /= np.linalg.norm(bc, axis=1, keepdims=True)
        ca /= np.linalg.norm(ca, axis=1, keepdims=True)
        tri = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], 0)
    p = tri.reshape(-1, 3)
    theta = np.arccos(np.clip(p[:, 2], -1, 1))
    phi = np.arctan2(p[:, 1], p[:, 0])
    r = radius * (1.0 + bump * np.sin(7 * theta) * np.sin(5 * phi))
    p = p * r[:, None] + np.asarray(centre)
    return p.reshape(-1, 3, 3).astype(np.float32)
