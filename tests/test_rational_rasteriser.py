"""An independent second statement of the item-buffer rule, in exact integer / rational arithmetic, against the oracle (and, on a
GPU, against libipp.so's id image).

The oracle (oracle/ipp_oracle.cpp, setup_tri / frag) and the CUDA rasteriser (csrc/raster.cuh) both state coverage through
homogeneous edge functions of the pixel ray built from eye-relative vertices.  The checker below shares nothing with that
formulation: it intersects the pixel-centre ray with the triangle's plane and decides inside / outside from three same-side tests
*in the plane of the triangle*, all in exact arithmetic on the exact values of the float32 vertices and of the float64 pixel-centre
coordinates.  The rule it states (UF/CameraEstimator.cpp:99-129, UF/ImageViewerCaptureTool.cpp:22-36,59-70):

  * the ray of pixel (px, py) goes through the pixel centre X = ((px + 0.5) / W * 2 - 1) tanX, Y likewise (image y up);
  * a triangle owns a fragment there when the ray meets it (both faces, boundary included) at planar depth zNear <= z <= zFar;
  * model 0: the nearest fragment wins; fragments within a relative 2^-16 of the nearest are ties and the lowest id wins;
  * model 1 (the OpenGL depth buffer the reference's viewer most likely had): window depth f / (f - n) (1 - n / z) quantised to 24
    bits, GL_LESS with the triangles drawn in id order, top-left rule for centres exactly on an edge.

Pixels whose answer hangs on less than 1e-9 (a centre within that of an edge, two surfaces within that of the tie band's rim) are
not compared: there the oracle's float64 rounding is allowed to decide.  Scenes are generic, so those are a handful at most.
"""
import math
from fractions import Fraction

import numpy as np
import pytest

from conftest import cuda_available, specs

TIE = Fraction(1, 65536)
SLACK = Fraction(1, 10 ** 9)


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def rational_item_buffer(tris, eye, basis, W, H, fov_deg, z_near, z_far, model=0):
    """ids [H, W] (row = py, image y up; -1 = background) and a mask of the pixels whose answer is robust.
    tris: float32 [T, 3, 3]; eye: 3 floats; basis = (f, s, u) with exactly representable orthonormal components."""
    F = Fraction
    T = len(tris)
    V = [[tuple(F(float(c)) for c in v) for v in t] for t in tris]
    eye = tuple(F(float(c)) for c in eye)
    f, s, u = (tuple(F(c) for c in b) for b in basis)
    tan_y = math.tan((fov_deg * 0.5) * 3.14159265358979323846 / 180.0)
    tan_x = tan_y * (float(fov_deg) / float(fov_deg))
    ids = -np.ones((H, W), dtype=np.int32)
    robust = np.ones((H, W), dtype=bool)
    zn, zf = F(z_near), F(z_far)
    planes = []
    for t in range(T):
        A, B, C = (_sub(v, eye) for v in V[t])
        n = _cross(_sub(B, A), _sub(C, A))
        planes.append((A, B, C, n, _dot(n, A)))
    for py in range(H):
        Y = F(((py + 0.5) / H * 2.0 - 1.0) * tan_y)
        for px in range(W):
            X = F(((px + 0.5) / W * 2.0 - 1.0) * tan_x)
            D = tuple(f[k] + X * s[k] + Y * u[k] for k in range(3))  # D.f == 1: the ray parameter is the planar depth
            frags = []
            shaky = False
            for t in range(T):
                A, B, C, n, nA = planes[t]
                if n == (0, 0, 0):
                    continue  # degenerate
                nD = _dot(n, D)
                if nD == 0:
                    continue  # grazing: no fragment
                z = nA / nD
                P = tuple(z * d for d in D)
                # same side of the three edges, seen along n (in the triangle's plane)
                e = [_dot(_cross(_sub(Q, R), _sub(P, R)), n) for R, Q in ((A, B), (B, C), (C, A))]
                scale = _dot(n, n)
                lens = [_dot(_sub(Q, R), _sub(Q, R)) for R, Q in ((A, B), (B, C), (C, A))]
                # e / (|n| |edge|) is the distance of P from the edge line; compare squares, relative to the depth
                near_edge = any(ev * ev <= SLACK * SLACK * scale * ln * z * z for ev, ln in zip(e, lens))
                inside = all(ev >= 0 for ev in e)
                if near_edge and not (zn * (1 - SLACK) <= z <= zf * (1 + SLACK)):
                    continue
                if near_edge:
                    shaky = True
                if not inside:
                    continue
                if abs(z - zn) <= SLACK * zn or abs(z - zf) <= SLACK * zf:
                    shaky = True
                if not (zn <= z <= zf):
                    continue
                on_edge = [ev == 0 for ev in e]
                frags.append((z, t, on_edge, (A, B, C, n, D)))
            if shaky:
                robust[py, px] = False
            if not frags:
                continue
            if model == 0:
                zmin = min(fr[0] for fr in frags)
                band = [fr for fr in frags if fr[0] <= zmin * (1 + TIE)]
                if any(abs(fr[0] - zmin * (1 + TIE)) <= SLACK * zmin for fr in frags):
                    robust[py, px] = False
                ids[py, px] = min(fr[1] for fr in band)
            else:
                best, best_d = -1, 1 << 24
                for z, t, on_edge, geo in sorted(frags, key=lambda fr: fr[1]):
                    if any(on_edge) and not _top_left_owns(geo, on_edge, s, u):
                        continue
                    zw = zf / (zf - zn) * (1 - zn / z)
                    q = zw * ((1 << 24) - 1) + F(1, 2)
                    d = q.numerator // q.denominator
                    if abs(q - round(q)) <= SLACK:
                        robust[py, px] = False
                    if d < best_d:
                        best, best_d = t, d
                ids[py, px] = best
    return ids, robust


def _top_left_owns(geo, on_edge, s, u):
    """top-left rule in the image plane: a centre exactly on an edge belongs to the triangle when the triangle's interior lies to
    the right of the edge (a left edge) or, for a horizontal edge, below it (a top edge; image y up)"""
    A, B, C, n, D = geo
    # project the vertices along the ray direction basis: image-plane coordinates of a point P in front of the eye are
    # (P.s / P.f, P.u / P.f); f is recovered from D = f + X s + Y u with s, u orthonormal to it
    X, Y = _dot(D, s), _dot(D, u)
    fvec = tuple(D[k] - X * s[k] - Y * u[k] for k in range(3))
    pts = []
    for P in (A, B, C):
        zf_ = _dot(P, fvec)
        if zf_ <= 0:
            return True  # a vertex behind the eye: the constructed cases never do this
        pts.append((_dot(P, s) / zf_, _dot(P, u) / zf_))
    cx, cy = (pts[0][0] + pts[1][0] + pts[2][0]) / 3, (pts[0][1] + pts[1][1] + pts[2][1]) / 3
    for k, on in enumerate(on_edge):
        if not on:
            continue
        (x0, y0), (x1, y1) = pts[k], pts[(k + 1) % 3]
        # inward normal of the edge: towards the centroid
        nx, ny = (y1 - y0), -(x1 - x0)
        if nx * (cx - x0) + ny * (cy - y0) < 0:
            nx, ny = -nx, -ny
        if not (nx > 0 or (nx == 0 and ny < 0)):
            return False
    return True


def _scene(seed, n_tris):
    """triangles on a 2^-8 grid (exact in float32) in front of, around and partly behind a camera at `eye` looking along +y"""
    rng = np.random.default_rng(seed)
    tris = []
    for k in range(n_tris):
        depth = rng.uniform(0.3, 5.0)
        centre = np.array([rng.uniform(-0.5, 0.5) * depth, depth, rng.uniform(-0.5, 0.5) * depth])
        size = rng.uniform(0.2, 1.2) * depth
        t = centre + rng.uniform(-0.5, 0.5, size=(3, 3)) * size
        tris.append(np.round(t * 256) / 256)
    # one triangle that crosses the plane of the eye, one beyond the far plane, one cut by the near plane, two coplanar duplicates
    tris.append(np.array([[-1.0, -1.0, -0.25], [1.5, 3.0, 0.0], [-0.5, 2.0, 0.75]]))
    tris.append(np.array([[-8.0, 11.0, -8.0], [8.0, 11.0, -8.0], [0.0, 11.0, 8.0]]))
    tris.append(np.array([[-0.05, 0.0625, -0.04], [-0.02, 0.0625, -0.04], [-0.1, 0.4, -0.05]]))
    tris.append(np.array([[0.5, 6.0, -2.0], [3.0, 6.0, -2.0], [3.0, 6.0, 1.0]]))
    tris.append(np.array([[0.5, 6.0, -2.0], [3.0, 6.0, -2.0], [3.0, 6.0, 1.0]]))
    return np.asarray(tris, dtype=np.float32)


def _sheet(seed, n=7):
    """a bumpy n x n vertex sheet in front of the camera, two triangles per cell, every interior edge shared by exactly two triangles
    (watertightness: no pixel inside the outline may show the backdrop triangle behind it), plus that backdrop"""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-1.3, 1.3, n)[:, None] + rng.uniform(-0.12, 0.12, size=(n, n))
    zs = np.linspace(-1.3, 1.3, n)[None, :] + rng.uniform(-0.12, 0.12, size=(n, n))
    ys = 3.0 + rng.uniform(-0.4, 0.4, size=(n, n))
    v = np.stack([xs, ys, zs], axis=2)
    tris = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = v[i, j], v[i + 1, j], v[i + 1, j + 1], v[i, j + 1]
            if (i + j) % 2:
                tris += [[a, b, c], [a, c, d]]
            else:
                tris += [[a, b, d], [b, c, d]]
    tris.append([[-6.0, 8.0, -6.0], [6.0, 8.0, -6.0], [0.0, 8.0, 7.0]])  # the backdrop: id 2 (n-1)^2
    return np.asarray(tris, dtype=np.float32)


CAMERAS = {
    # f, s = f x up / |.|, u = s x f with up = (0, 0, 1): all components exactly representable
    "along_y": ((0.0, 0.0, 0.0), ((0, 1, 0), (1, 0, 0), (0, 0, 1))),
    "along_minus_x": ((0.5, 1.0, 0.25), ((-1, 0, 0), (0, 1, 0), (0, 0, 1))),
}


def _oracle_ids(oracle_mod, tmp_path, tris, cam, px, model):
    p = str(tmp_path / "scene.ippm")
    oracle_mod.write_ippm(p, tris)
    o = oracle_mod.Oracle(p, specs(px, front=1, down=0), post_processing=True)
    o.set_depth_model(model)
    eye, (f, s, u) = CAMERAS[cam]
    ids = o.render_ids(list(eye), [eye[k] + f[k] for k in range(3)], [0, 0, 1])
    o.close()
    return ids


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("cam,seed,px", [("along_y", 1, 24), ("along_y", 2, 16), ("along_minus_x", 3, 24)])
def test_oracle_item_buffer_equals_the_exact_rational_rule(oracle_mod, tmp_path, cam, seed, px, model):
    tris = _scene(seed, 14)
    if cam == "along_minus_x":  # the same scene turned so that it lies in front of this camera
        tris = np.stack([-tris[:, :, 1] + 0.5, tris[:, :, 0] + 1.0, tris[:, :, 2] + 0.25], axis=2).astype(np.float32)
    eye, basis = CAMERAS[cam]
    want, robust = rational_item_buffer(tris, eye, basis, px, px, 46, 0.1, 10.0, model)
    got = _oracle_ids(oracle_mod, tmp_path, tris, cam, px, model)
    assert robust.mean() > 0.98
    assert (want >= 0).mean() > 0.3  # the scene fills a good part of the image
    assert len(np.unique(want[want >= 0])) >= 6
    assert np.array_equal(got[robust], want[robust]), np.argwhere((got != want) & robust)[:10]


def test_shared_edges_are_watertight(oracle_mod, tmp_path):
    """A sheet of 72 triangles whose interior edges are each shared by two of them, in front of a backdrop: by the exact rule every
    pixel whose ray passes inside the sheet's outline shows a sheet triangle; the oracle must agree pixel for pixel (a crack between
    two triangles would show the backdrop's id)."""
    tris = _sheet(11)
    backdrop = len(tris) - 1
    want, robust = rational_item_buffer(tris, (0, 0, 0), CAMERAS["along_y"][1], 24, 24, 46, 0.1, 10.0, 0)
    got = _oracle_ids(oracle_mod, tmp_path, tris, "along_y", 24, 0)
    assert robust.mean() > 0.97 and (want != backdrop).mean() > 0.5
    assert np.array_equal(got[robust], want[robust])
    inner = want[6:18, 6:18]
    assert (inner != backdrop).all() and (inner >= 0).all()  # the middle of the image is all sheet


def test_the_two_depth_models_differ_only_where_they_should(oracle_mod, tmp_path):
    """Two parallel sheets 60 um apart at 6 m, the farther one with the lower id.  Model 0: 60 um is inside the relative 2^-16 tie
    band (92 um at 6 m), so the lower id wins although it is behind.  Model 1: one step of a 24-bit window depth at z = 6 m, n = 0.1,
    f = 10 is z^2 (f - n) / (f n) 2^-24 = 21 um, so GL_LESS keeps the nearer sheet.  The exact rule says the same for both."""
    near_sheet = [[-3.0, 6.0, -3.0], [3.0, 6.0, -3.0], [0.0, 6.0, 3.0]]
    far_sheet = [[-3.0, 6.00006, -3.0], [3.0, 6.00006, -3.0], [0.0, 6.00006, 3.0]]
    tris = np.asarray([far_sheet, near_sheet], dtype=np.float32)  # id 0 is the farther one
    assert float(tris[0, 0, 1]) > float(tris[1, 0, 1])
    for model, owner in ((0, 0), (1, 1)):
        got = _oracle_ids(oracle_mod, tmp_path, tris, "along_y", 16, model)
        want, robust = rational_item_buffer(tris, (0, 0, 0), CAMERAS["along_y"][1], 16, 16, 46, 0.1, 10.0, model)
        assert np.array_equal(got[robust], want[robust])
        assert set(np.unique(got[got >= 0])) == {owner}


def test_top_left_rule_on_a_shared_edge_through_pixel_centres(oracle_mod, tmp_path):
    """A quad at depth 1 split along the image diagonal, which passes through the centres of the diagonal pixels.  Model 0 draws
    those centres for both triangles and the tie goes to the lower id; model 1 gives each centre to exactly one triangle by the
    top-left rule, whatever the ids.  Both must match the exact rule."""
    t = math.tan(math.radians(23.0))
    k = 2.0 * t  # the quad is twice the image: its diagonal y = x passes through every diagonal pixel centre
    lower = [[-k, 1.0, -k], [k, 1.0, -k], [k, 1.0, k]]     # below the diagonal
    upper = [[-k, 1.0, -k], [k, 1.0, k], [-k, 1.0, k]]     # above it
    for order in ((lower, upper), (upper, lower)):
        tris = np.asarray(order, dtype=np.float32)
        for model in (0, 1):
            got = _oracle_ids(oracle_mod, tmp_path, tris, "along_y", 16, model)
            want, robust = rational_item_buffer(tris, (0, 0, 0), CAMERAS["along_y"][1], 16, 16, 46, 0.1, 10.0, model)
            off = ~np.eye(16, dtype=bool)
            assert np.array_equal(got[off & robust], want[off & robust])
            assert (got >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("cam,seed,px", [("along_y", 1, 24), ("along_minus_x", 3, 32), ("along_y", -11, 32)])
def test_gpu_item_buffer_equals_the_exact_rational_rule(oracle_mod, tmp_path, cam, seed, px):
    """libipp.so's id image (the test hook of K4: the same rasteriser the edge bitmaps come from) against the exact rule.  The
    kernel rounds its 13 coefficients to float32, so a centre may be decided differently within ~1e-6 of an edge: pixels are
    compared where the exact rule is robust to 1e-5 of the triangle size, and at most 1 % of them may be excluded."""
    from inspection_path_planning_b200 import cpp_binding
    tris = _scene(seed, 14) if seed >= 0 else _sheet(-seed)  # negative seed: the watertight sheet (shared edges, float coefficients)
    if cam == "along_minus_x":
        tris = np.stack([-tris[:, :, 1] + 0.5, tris[:, :, 0] + 1.0, tris[:, :, 2] + 0.25], axis=2).astype(np.float32)
    eye, basis = CAMERAS[cam]
    global SLACK
    saved = SLACK
    SLACK = Fraction(1, 10 ** 5)
    try:
        want, robust = rational_item_buffer(tris, eye, basis, px, px, 46, 0.1, 10.0, 0)
    finally:
        SLACK = saved
    p = str(tmp_path / "scene.ippm")
    oracle_mod.write_ippm(p, tris)
    g = cpp_binding.PlanCoverageEstimator(p, specs(px, front=1, down=0), True)
    try:
        f = basis[0]
        got = g.render_ids(list(eye), [eye[k] + f[k] for k in range(3)], [0, 0, 1])
    finally:
        g.close()
    assert robust.mean() > 0.97
    assert np.array_equal(got[robust], want[robust]), np.argwhere((got != want) & robust)[:10]
