"""ctypes wrapper around oracle/libipp_oracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED: see the header of ipp_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libipp_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ipp_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_num_triangles.restype = C.c_int64
        for f in ("oracle_max_allowed_energy", "oracle_total_area", "oracle_collision_penalty",
                  "oracle_energy_of_positions", "oracle_coverage_of_bitmap"):
            getattr(L, f).restype = C.c_double
        for name in dir(L):
            pass
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pack_plans(plans):
    """list of plans (each a sequence of genes [id] or [id, angle]) -> (genes[n,width] f64, offsets i64, width)"""
    width = 1
    for p in plans:
        for g in p:
            width = max(width, len(g))
    flat = []
    offsets = [0]
    for p in plans:
        for g in p:
            g = list(g)
            if len(g) != width:
                raise ValueError("mixed gene widths in one population")
            flat.append([float(x) for x in g])
        offsets.append(len(flat))
    genes = np.asarray(flat, dtype=np.float64).reshape(-1, width)
    return np.ascontiguousarray(genes), np.asarray(offsets, dtype=np.int64), width


class Oracle:
    """Mirrors the reference's PlanCoverageEstimator surface (cpp_binding.i:25-39) on the CPU."""

    def __init__(self, scene, sensor_specs, post_processing=False, start_location=None, plan_loops_around=False):
        L = lib()
        specs = np.asarray(sensor_specs, dtype=np.float64)
        assert specs.size == 8
        start = None if start_location is None else np.asarray(start_location, dtype=np.float64)
        h = C.c_void_p()
        rc = L.oracle_create(os.fsencode(scene), _dp(specs), int(post_processing), None if start is None else _dp(start),
                             int(plan_loops_around), C.byref(h))
        if rc:
            raise ValueError(L.oracle_last_error().decode())
        self._h = h
        self._L = L
        self.has_start = start is not None
        self.loops = bool(plan_loops_around)
        self.T = int(L.oracle_num_triangles(h))
        self.W32 = (self.T + 31) // 32

    def close(self):
        if getattr(self, "_h", None):
            self._L.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self._L.oracle_last_error().decode())

    # ---- reference API ----
    def getNumberOfBoxes(self):
        return int(self._L.oracle_num_boxes(self._h))

    def getMaxAllowedEnergy(self):
        return float(self._L.oracle_max_allowed_energy(self._h))

    def getSimplePlans(self):
        n, g = C.c_int(), C.c_int()
        self._chk(self._L.oracle_simple_plans(self._h, C.byref(n), C.byref(g), None, None))
        ids = np.zeros(max(g.value, 1), dtype=np.float64)
        off = np.zeros(n.value + 1, dtype=np.int32)
        self._chk(self._L.oracle_simple_plans(self._h, C.byref(n), C.byref(g), _dp(ids), off.ctypes.data_as(C.POINTER(C.c_int))))
        return tuple(tuple((float(ids[k]),) for k in range(off[i], off[i + 1])) for i in range(n.value))

    def evaluatePlan(self, plan, memoization, how_to_plot=3, disableEnergyLimit=False, full=False):
        genes, _, width = pack_plans([plan])
        out = np.zeros(4, dtype=np.float64)
        self._chk(self._L.oracle_evaluate_plan(self._h, _dp(genes), genes.shape[0], width, int(memoization), int(disableEnergyLimit), _dp(out)))
        return tuple(out) if full else (float(out[0]), float(out[1]))

    def evaluatePopulation(self, plans, memoization, disableEnergyLimit=False):
        genes, offsets, width = plans if isinstance(plans, tuple) else pack_plans(plans)
        pop = offsets.size - 1
        out = np.zeros((pop, 4), dtype=np.float64)
        genes = np.ascontiguousarray(genes, dtype=np.float64)
        self._chk(self._L.oracle_evaluate_population(self._h, _dp(genes), width, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     C.c_int64(pop), int(memoization), int(disableEnergyLimit), _dp(out)))
        return out

    def updateMemoisedEdges(self, plans):
        genes, offsets, width = pack_plans(plans)
        rem = C.c_int()
        self._chk(self._L.oracle_update_memoised_edges(self._h, _dp(genes), width, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                                       C.c_int64(offsets.size - 1), C.byref(rem)))
        return rem.value

    def interpretAndExportPlan(self, plan, _a=None, _b=None):
        genes, _, width = pack_plans([plan])
        n = genes.shape[0]
        pos = np.zeros((n + 1, 3))
        dirs = np.zeros((n + 1, 3))
        npos, ndir = C.c_int(), C.c_int()
        self._chk(self._L.oracle_interpret_plan(self._h, _dp(genes), n, width, _dp(pos), _dp(dirs), C.byref(npos), C.byref(ndir)))
        return [pos[:npos.value].tolist(), dirs[:ndir.value].tolist()]

    # ---- inspection hooks for the parity tests ----
    def memo_size(self):
        return int(self._L.oracle_memo_size(self._h))

    def total_area(self):
        return float(self._L.oracle_total_area(self._h))

    def collision_penalty(self):
        return float(self._L.oracle_collision_penalty(self._h))

    def scene_info(self):
        c = np.zeros(3)
        b = np.zeros(6)
        self._L.oracle_scene_info(self._h, _dp(c), _dp(b))
        return c, b

    def image_size(self):
        w, h = C.c_int(), C.c_int()
        self._L.oracle_image_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def areas(self):
        a = np.zeros(self.T)
        self._L.oracle_areas(self._h, _dp(a))
        return a

    def vertices(self):
        v = np.zeros((self.T, 3, 3), dtype=np.float32)
        self._L.oracle_vertices(self._h, v.ctypes.data_as(C.POINTER(C.c_float)))
        return v

    def box_centres(self):
        n = self.getNumberOfBoxes()
        out = np.zeros((n, 3))
        self._L.oracle_box_centres(self._h, _dp(out))
        return out

    def grid(self):
        nz, nx, ny = C.c_int(), C.c_int(), C.c_int()
        self._L.oracle_grid_dims(self._h, C.byref(nz), C.byref(nx), C.byref(ny))
        g = np.zeros((nz.value, nx.value, ny.value), dtype=np.int32)
        self._L.oracle_grid(self._h, g.ctypes.data_as(C.POINTER(C.c_int)))
        return g

    def edge_bitmap(self, prev, curr, angle=0.0, classify=False):
        seen = np.zeros(self.W32, dtype=np.uint32)
        p32 = C.POINTER(C.c_uint32)
        if classify:
            st = np.zeros(self.W32, dtype=np.uint32)
            le = np.zeros(self.W32, dtype=np.uint32)
            self._chk(self._L.oracle_edge_bitmap(self._h, int(prev), int(curr), C.c_double(angle), seen.ctypes.data_as(p32),
                                                 st.ctypes.data_as(p32), le.ctypes.data_as(p32)))
            return seen, st, le
        self._chk(self._L.oracle_edge_bitmap(self._h, int(prev), int(curr), C.c_double(angle), seen.ctypes.data_as(p32), None, None))
        return seen

    def edge_bitmap_xyz(self, a, b, angle=0.0, classify=False):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        seen = np.zeros(self.W32, dtype=np.uint32)
        p32 = C.POINTER(C.c_uint32)
        if classify:
            st = np.zeros(self.W32, dtype=np.uint32)
            le = np.zeros(self.W32, dtype=np.uint32)
            self._chk(self._L.oracle_edge_bitmap_xyz(self._h, _dp(a), _dp(b), C.c_double(angle), seen.ctypes.data_as(p32),
                                                     st.ctypes.data_as(p32), le.ctypes.data_as(p32)))
            return seen, st, le
        self._chk(self._L.oracle_edge_bitmap_xyz(self._h, _dp(a), _dp(b), C.c_double(angle), seen.ctypes.data_as(p32), None, None))
        return seen

    def render_ids(self, eye, center, up):
        w, h = self.image_size()
        ids = np.zeros((h, w), dtype=np.int32)
        e, c, u = (np.asarray(x, dtype=np.float64) for x in (eye, center, up))
        self._chk(self._L.oracle_render_ids(self._h, _dp(e), _dp(c), _dp(u), ids.ctypes.data_as(C.POINTER(C.c_int32))))
        return ids

    def edge_collides(self, a, b):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        return bool(self._L.oracle_edge_collides(self._h, _dp(a), _dp(b)))

    def box_hits_mesh(self, c, half_side):
        c = np.asarray(c, dtype=np.float64)
        return bool(self._L.oracle_box_hits_mesh(self._h, _dp(c), C.c_double(half_side)))

    def sampling_positions(self, a, b, interval=-1.0):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        out = np.zeros((4096, 3), dtype=np.float32)
        n = self._L.oracle_sampling_positions(self._h, _dp(a), _dp(b), C.c_double(interval), out.ctypes.data_as(C.POINTER(C.c_float)), 4096)
        return out[:n].copy()

    def heading(self, a, b, angle=0.0):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        out = np.zeros(3)
        self._L.oracle_heading(self._h, _dp(a), _dp(b), C.c_double(angle), _dp(out))
        return out

    def energy_of_positions(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        pl = C.c_double()
        nc = C.c_int()
        e = self._L.oracle_energy_of_positions(self._h, _dp(pos), pos.shape[0], C.byref(pl), C.byref(nc))
        return float(e), pl.value, nc.value

    def coverage_of_bitmap(self, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        return float(self._L.oracle_coverage_of_bitmap(self._h, bits.ctypes.data_as(C.POINTER(C.c_uint32))))

    def set_margins(self, eps_px, eps_depth):
        self._L.oracle_set_margins(self._h, C.c_double(eps_px), C.c_double(eps_depth))

    def set_depth_model(self, model):
        """0: planar depth, relative 2^-16 tie band, lowest ID (default).  1: OpenGL-style 24-bit window depth, GL_LESS in ID order,
        top-left fill rule (see ipp_oracle.cpp).  Clears the memo: its bitmaps belong to the model they were rendered with."""
        self._L.oracle_set_depth_model(self._h, int(model))
        self.updateMemoisedEdges([])

    def counters(self):
        out = np.zeros(3, dtype=np.uint64)
        self._L.oracle_counters(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return {"snapshots": int(out[0]), "pixels": int(out[1]), "collision_queries": int(out[2])}


def edge_prism_planes(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.zeros((6, 4))
    lib().oracle_edge_prism_planes(_dp(a), _dp(b), _dp(out))
    return out


def moore_trace(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros((8 * img.size + 16, 2), dtype=np.int32)
    n = lib().oracle_moore_trace(img.ctypes.data_as(C.POINTER(C.c_ubyte)), img.shape[0], img.shape[1],
                                 out.ctypes.data_as(C.POINTER(C.c_int)), out.shape[0])
    return [tuple(x) for x in out[:n].tolist()]


def write_ippm(path, tris):
    """tris: float32 [T,3,3] -> flat test mesh format read by both the oracle and libipp."""
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"IPPM")
        f.write(np.uint32(tris.shape[0]).tobytes())
        f.write(tris.tobytes())
