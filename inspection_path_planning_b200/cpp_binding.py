"""Drop-in for the reference's SWIG module `cpp_wrapper.cpp_binding`
(plan_optimizer/cpp_wrapper/cpp_binding.i:1-40, cpp_binding.py:512-549): same module-level names, same
class, same method names, argument order and return shapes -- backed by libipp.so on a B200.

    from inspection_path_planning_b200 import cpp_binding
    ev = cpp_binding.PlanCoverageEstimator(scene, [2.0, 46, 46, 1024, 0.1, 10, 1, 1])
    coverage, energy = ev.evaluatePlan([[3], [17], [4]], True, cpp_binding.nothing)
    fitness = ev.evaluatePopulation(list_of_plans, True)      # the one batched entry point

Differences from the SWIG build, all deliberate:
  * startLocation is copied (the reference keeps a borrowed pointer to a `Line`); any sequence of 3 numbers works.
  * C++ exceptions become Python exceptions instead of terminating the interpreter (the reference has no %exception).
  * the GL viewers (how_to_plot != nothing, storePlanImage, viewMatrixSelector) are out of scope and raise.
"""
import ctypes as C
import itertools
import os

import numpy as np

from . import _lib

# enum plotting_style, cpp_binding.i:20
normal, circulating, image, nothing = 0, 1, 2, 3


class Line(list):
    """std::vector<double> stand-in (cpp_binding.i:13)."""


class Array(list):
    """std::vector<std::vector<double>> stand-in (cpp_binding.i:14)."""


class StringVector(list):
    """std::vector<std::string> stand-in (cpp_binding.i:16)."""


def viewMatrixSelector(sceneFileName):
    """cpp_binding.i:23 -- opens an interactive OSG viewer in the reference; out of scope here."""
    raise NotImplementedError("viewMatrixSelector needs an interactive OpenGL viewer; not part of the B200 evaluator")


def _dp(a):
    return a.ctypes.data_as(_lib.c_f64p)


def pack_population(plans, with_angles=False):
    """list of plans -> (ids int32[nGenes], offsets int64[pop+1]) or, with_angles=True, (ids, angles float32[nGenes] | None, offsets).
    A gene is [id], a bare id, or [id, angleOffset] (SIMPLIFIED_VIEW_ANGLE = False individuals, individuals.py:19-22); angles is
    None when every offset is zero -- the straight-ahead heading evaluates like [id] and shares its memo.
    Gene -> id uses the reference's rounding `(size_t)(gene + 0.5)` (PlanInterpreterBoxOrder.cpp:280)."""
    offsets = np.zeros(len(plans) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(np.fromiter((len(p) for p in plans), dtype=np.int64, count=len(plans)))
    flat = ang = None
    try:
        # fast path (C-level loops): bare ids, or genes that are all sequences of length 1 or 2
        genes = list(itertools.chain.from_iterable(plans))
        if not genes:
            flat, ang = np.zeros(0), np.zeros(0)
        elif isinstance(genes[0], (list, tuple, np.ndarray)):
            lens = set(map(len, genes))
            if lens <= {1, 2}:
                flat = np.fromiter((g[0] for g in genes), dtype=np.float64, count=len(genes))
                ang = np.zeros(len(genes)) if lens == {1} else np.fromiter((g[1] if len(g) == 2 else 0.0 for g in genes), dtype=np.float64,
                                                                          count=len(genes))
            else:
                raise ValueError("a gene is [id] or [id, angleOffset]")
        else:
            try:
                flat = np.asarray(genes, dtype=np.float64)
            except ValueError:  # bare ids mixed with sequences
                flat = None
            if flat is not None and flat.ndim != 1:
                flat = None
            if flat is not None:
                ang = np.zeros(flat.size)
    except TypeError:
        flat = ang = None
    if flat is None:  # mixed gene shapes: gene by gene
        flat, ang = [], []
        for p in plans:
            for g in p:
                a = 0.0
                if isinstance(g, (list, tuple, np.ndarray)):
                    if len(g) == 2:
                        a = float(g[1])
                    elif len(g) != 1:
                        raise ValueError("a gene is [id] or [id, angleOffset]")
                    g = g[0]
                flat.append(g)
                ang.append(a)
        ang = np.asarray(ang, dtype=np.float64)
    vals = np.asarray(flat, dtype=np.float64)
    ids = np.trunc(vals + 0.5)
    ids[vals + 0.5 < 0] = 0.0  # a negative gene (the -1 cells of the seed plans) truncates to 0 like the reference
    ids = np.ascontiguousarray(ids.astype(np.int32))
    has_angles = bool(np.any(ang != 0.0))
    if not with_angles:
        if has_angles:
            raise NotImplementedError("[id, angle] genes with a non-zero offset need the host-pointer evaluation calls (evaluatePlan, "
                                      "evaluatePopulation, updateMemoisedEdges)")
        return ids, offsets
    angles = np.ascontiguousarray(ang.astype(np.float32)) if has_angles else None
    return ids, angles, offsets


def _fp(a):
    return None if a is None else a.ctypes.data_as(_lib.c_f32p)


def load_mesh(scene_file_name):
    """The evaluator's mesh ingest on its own (host code): float32 [T, 3, 3], triangles in file order -- what
    osgDB::readNodeFile + the TriangleFunctor enumeration hand to TriangleData (SceneKeeper.cpp:73-86, :29)."""
    L = _lib.load()
    n = C.c_int64()
    _lib.check(L.ipp_load_mesh(os.fsencode(scene_file_name), None, 0, C.byref(n)))
    out = np.zeros((int(n.value), 3, 3), dtype=np.float32)
    _lib.check(L.ipp_load_mesh(os.fsencode(scene_file_name), out.ctypes.data_as(_lib.c_f32p), int(n.value), C.byref(n)))
    return out


class PlanCoverageEstimator(object):
    """PlanCoverageEstimator (cpp_binding.i:25-39; PlanCoverageEstimator.h:51-165)."""

    def __init__(self, sceneFileName, sensorSpecs, postProcessing=False, startLocation=None, planLoopsAround=False,
                 printerFriendly=True, device=None):
        L = _lib.load()
        if L.ipp_device_count() < 1:
            raise RuntimeError("no CUDA device visible: the B200 evaluator has no CPU fallback")
        specs = np.ascontiguousarray(list(sensorSpecs), dtype=np.float64)
        if specs.size != 8:
            raise ValueError("sensorSpecs needs 8 values: [sampleInterval, fov_x, fov_y, imageHeight, zNear, zFar, front, down]")
        start = None
        if startLocation is not None:
            start = np.ascontiguousarray(list(startLocation), dtype=np.float64)
            if start.size != 3:
                raise ValueError("startLocation needs 3 values")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % max(1, L.ipp_device_count())
        h = _lib.H()
        _lib.check(L.ipp_create(os.fsencode(sceneFileName), _dp(specs), int(bool(postProcessing)), None if start is None else _dp(start),
                                int(bool(planLoopsAround)), int(bool(printerFriendly)), int(device), C.byref(h)))
        self._L = L
        self._h = h
        self.device = device
        self.T = int(L.ipp_num_triangles(h))
        self.W32 = int(L.ipp_bitmap_words(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.ipp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's methods ----
    def evaluatePlan(self, plan, memoization, how_to_plot, disableEnergyLimit=False):
        """-> (coverageScore, energy); cpp_binding.i:31.  Genes are [id] or [id, angleOffsetRad]."""
        width = 1
        for g in plan:
            if isinstance(g, (list, tuple, np.ndarray)) and len(g) == 2 and float(g[1]) != 0.0:
                width = 2
        genes = np.zeros((len(plan), width), dtype=np.float64)
        for i, g in enumerate(plan):
            if isinstance(g, (list, tuple, np.ndarray)):
                if len(g) not in (1, 2):
                    raise ValueError("a gene is [id] or [id, angleOffset]")
                genes[i, :min(width, len(g))] = g[:width]
            else:
                genes[i, 0] = g
        out = np.zeros(4, dtype=np.float64)
        _lib.check(self._L.ipp_evaluate_plan(self._h, _dp(genes), genes.shape[0], width, int(bool(memoization)), int(how_to_plot),
                                             int(bool(disableEnergyLimit)), _dp(out)))
        return (float(out[0]), float(out[1]))

    def getNumberOfBoxes(self):
        return int(self._L.ipp_num_boxes(self._h))

    def updateMemoisedEdges(self, allSolutions):
        if isinstance(allSolutions, tuple):
            ids, angles, offsets = (allSolutions[0], None, allSolutions[1]) if len(allSolutions) == 2 else allSolutions
        else:
            ids, angles, offsets = pack_population(allSolutions, with_angles=True)
        rem = C.c_int64()
        _lib.check(self._L.ipp_update_memoised_edges(self._h, ids.ctypes.data_as(_lib.c_i32p), _fp(angles), offsets.ctypes.data_as(_lib.c_i64p),
                                                     offsets.size - 1, C.byref(rem)))
        return int(rem.value)

    def getSimplePlans(self):
        n, g = C.c_int(), C.c_int()
        _lib.check(self._L.ipp_simple_plans(self._h, C.byref(n), C.byref(g), None, None))
        ids = np.zeros(max(g.value, 1), dtype=np.float64)
        off = np.zeros(n.value + 1, dtype=np.int32)
        _lib.check(self._L.ipp_simple_plans(self._h, C.byref(n), C.byref(g), _dp(ids), off.ctypes.data_as(_lib.c_i32p)))
        return tuple(tuple((float(ids[k]),) for k in range(off[i], off[i + 1])) for i in range(n.value))

    def storePlanImage(self, plan, viewMatrix, storagePath):
        raise NotImplementedError("storePlanImage renders through an OpenGL viewer; not part of the B200 evaluator")

    def getMaxAllowedEnergy(self):
        return float(self._L.ipp_max_allowed_energy(self._h))

    def interpretAndExportPlan(self, plan, arg3=None, arg4=None):
        """-> [positions, headings]; cpp_binding.i:38 (SWIG INOUT typemap)."""
        width = 1
        for g in plan:
            if isinstance(g, (list, tuple, np.ndarray)):
                width = max(width, len(g))
        genes = np.zeros((len(plan), width), dtype=np.float64)
        for i, g in enumerate(plan):
            if isinstance(g, (list, tuple, np.ndarray)):
                genes[i, :len(g)] = g
            else:
                genes[i, 0] = g
        n = genes.shape[0]
        pos = np.zeros((n + 1, 3))
        dirs = np.zeros((n + 1, 3))
        npos, ndir = C.c_int(), C.c_int()
        _lib.check(self._L.ipp_interpret_plan(self._h, _dp(genes), n, width, _dp(pos), _dp(dirs), C.byref(npos), C.byref(ndir)))
        return [pos[:npos.value].tolist(), dirs[:ndir.value].tolist()]

    # ---- the batched entry point ----
    def evaluatePopulation(self, plans, memoization=True, disableEnergyLimit=False, union_bitmap=False, out=None):
        """Evaluates a whole population in one call.  `plans` is a list of plans ([id] or [id, angleOffset] genes) or a packed
        (ids int32, offsets int64) pair / (ids, angles float32, offsets) triple.  Returns float64 [pop, 4]: columns 0-1 are evaluatePlan's (coverageScore, energy), 2 = path length,
        3 = number of colliding edges.  With union_bitmap=True also returns the OR of the seen sets (uint32 words)."""
        if isinstance(plans, tuple):
            ids, angles, offsets = (plans[0], None, plans[1]) if len(plans) == 2 else plans
        else:
            ids, angles, offsets = pack_population(plans, with_angles=True)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        pop = offsets.size - 1
        fit = out if out is not None else np.empty((pop, 4), dtype=np.float64)
        ub = np.zeros(self.W32, dtype=np.uint32) if union_bitmap else None
        _lib.check(self._L.ipp_evaluate_population(self._h, ids.ctypes.data_as(_lib.c_i32p), _fp(angles), offsets.ctypes.data_as(_lib.c_i64p), pop,
                                                   int(bool(memoization)), int(bool(disableEnergyLimit)), _dp(fit),
                                                   None if ub is None else ub.ctypes.data_as(_lib.c_u32p)))
        return (fit, ub) if union_bitmap else fit

    def evaluatePopulationSharded(self, plans, memoization=True, disableEnergyLimit=False, out=None):
        """Multi-GPU form (after comm_init): every rank passes the same whole population.  The edges nobody has rendered yet are
        rendered once across the ranks and their bitmap rows exchanged (NCCL all-gather), then this rank evaluates its contiguous
        shard and one more all-gather returns every plan's fitness row to every rank."""
        if isinstance(plans, tuple):
            ids, angles, offsets = (plans[0], None, plans[1]) if len(plans) == 2 else plans
        else:
            ids, angles, offsets = pack_population(plans, with_angles=True)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        pop = offsets.size - 1
        fit = out if out is not None else np.empty((pop, 4), dtype=np.float64)
        _lib.check(self._L.ipp_evaluate_population_sharded(self._h, ids.ctypes.data_as(_lib.c_i32p), _fp(angles), offsets.ctypes.data_as(_lib.c_i64p),
                                                           pop, int(bool(memoization)), int(bool(disableEnergyLimit)), _dp(fit)))
        return fit

    def evaluatePopulationResident(self, plans, memoization=True, disableEnergyLimit=False):
        """The device-resident entry point (ipp_evaluate_population_device) behind a convenience wrapper: the population is put
        into device memory first, the call itself only sees device pointers, the fitness rows are read back afterwards."""
        if isinstance(plans, tuple):
            ids, angles, offsets = (plans[0], None, plans[1]) if len(plans) == 2 else plans
        else:
            ids, angles, offsets = pack_population(plans, with_angles=True)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        pop, n_genes = offsets.size - 1, int(ids.size)
        fit = np.zeros((pop, 4), dtype=np.float64)
        if pop == 0:
            return fit
        bufs = []

        def dev(arr):
            p = C.c_void_p()
            _lib.check(self._L.ipp_device_alloc(self._h, max(int(arr.nbytes), 16), C.byref(p)))
            bufs.append(p)
            if arr.nbytes:
                _lib.check(self._L.ipp_memcpy_h2d(self._h, p, arr.ctypes.data_as(C.c_void_p), int(arr.nbytes)))
            return p
        try:
            d_ids, d_off, d_fit = dev(ids), dev(offsets), dev(fit)
            d_ang = dev(angles) if angles is not None else None
            _lib.check(self._L.ipp_evaluate_population_device(self._h, d_ids, d_ang, d_off, pop, n_genes, int(bool(memoization)),
                                                              int(bool(disableEnergyLimit)), d_fit))
            self.sync()
            _lib.check(self._L.ipp_memcpy_d2h(self._h, fit.ctypes.data_as(C.c_void_p), d_fit, int(fit.nbytes)))
        finally:
            for p in bufs:
                self._L.ipp_device_free(self._h, p)
        return fit

    # ---- scene facts and test hooks ----
    def memo_size(self):
        return int(self._L.ipp_memo_size(self._h))

    def evictions(self):
        """How often the finite edge-bitmap row pool ran full and dropped the rows of edges the running call did not use."""
        return int(self._L.ipp_evictions(self._h))

    def clear_memo(self):
        _lib.check(self._L.ipp_clear_memo(self._h))

    def total_area(self):
        return float(self._L.ipp_total_area(self._h))

    def collision_penalty(self):
        return float(self._L.ipp_collision_penalty(self._h))

    def scene_info(self):
        c, b = np.zeros(3), np.zeros(6)
        self._L.ipp_scene_info(self._h, _dp(c), _dp(b))
        return c, b

    def image_size(self):
        w, h = C.c_int(), C.c_int()
        self._L.ipp_image_size(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def areas(self):
        a = np.zeros(self.T)
        _lib.check(self._L.ipp_areas(self._h, _dp(a)))
        return a

    def box_centres(self):
        out = np.zeros((self.getNumberOfBoxes(), 3))
        _lib.check(self._L.ipp_box_centres(self._h, _dp(out)))
        return out

    def grid(self):
        nz, nx, ny = C.c_int(), C.c_int(), C.c_int()
        self._L.ipp_grid_dims(self._h, C.byref(nz), C.byref(nx), C.byref(ny))
        g = np.zeros((nz.value, nx.value, ny.value), dtype=np.int32)
        _lib.check(self._L.ipp_grid(self._h, g.ctypes.data_as(_lib.c_i32p)))
        return g

    def edge_bitmap(self, prev, curr, angle=0.0):
        out = np.zeros(self.W32, dtype=np.uint32)
        _lib.check(self._L.ipp_edge_bitmap(self._h, int(prev), int(curr), float(angle), out.ctypes.data_as(_lib.c_u32p)))
        return out

    def edges_collide(self, edges):
        e = np.ascontiguousarray(edges, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(e.shape[0], dtype=np.uint8)
        _lib.check(self._L.ipp_edges_collide(self._h, _dp(e), e.shape[0], out.ctypes.data_as(_lib.c_u8p)))
        return out.astype(bool)

    def boxes_hit_mesh(self, centres, half_side):
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(c.shape[0], dtype=np.uint8)
        _lib.check(self._L.ipp_boxes_hit_mesh(self._h, _dp(c), c.shape[0], float(half_side), out.ctypes.data_as(_lib.c_u8p)))
        return out.astype(bool)

    def render_ids(self, eye, center, up):
        w, h = self.image_size()
        out = np.zeros((h, w), dtype=np.int32)
        e, c, u = (np.ascontiguousarray(x, dtype=np.float64) for x in (eye, center, up))
        _lib.check(self._L.ipp_render_ids(self._h, _dp(e), _dp(c), _dp(u), out.ctypes.data_as(_lib.c_i32p)))
        return out

    def bvh_selftest(self, rays):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        ids = np.zeros(r.shape[0], dtype=np.int32)
        mism = C.c_int64()
        _lib.check(self._L.ipp_bvh_selftest(self._h, r.ctypes.data_as(_lib.c_f32p), r.shape[0], ids.ctypes.data_as(_lib.c_i32p), C.byref(mism)))
        return ids, int(mism.value)

    def leaf_table(self, eye, center, up=(0.0, 0.0, 1.0), cap=16384):
        """Depth-sorted leaf table of one camera pose (binned visibility kernel): dict of numpy arrays + the 32-row tile mask."""
        e, c, u = (np.ascontiguousarray(v, dtype=np.float64) for v in (eye, center, up))
        out = np.zeros(cap, dtype=np.uint64)
        mask = np.zeros(32, dtype=np.uint32)
        n = C.c_int64()
        first = np.zeros(cap, dtype=np.int32)
        cnt = np.zeros(cap, dtype=np.int32)
        _lib.check(self._L.ipp_leaf_table(self._h, e.ctypes.data_as(_lib.c_f64p), c.ctypes.data_as(_lib.c_f64p), u.ctypes.data_as(_lib.c_f64p),
                                          out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n), mask.ctypes.data_as(_lib.c_u32p),
                                          first.ctypes.data_as(_lib.c_i32p), cnt.ctypes.data_as(_lib.c_i32p)))
        m = min(int(n.value), cap)
        k = out[:m]
        return {"count": int(n.value), "zq": (k >> np.uint64(48)).astype(np.int64), "tx0": ((k >> np.uint64(43)) & np.uint64(31)).astype(np.int64),
                "tx1": ((k >> np.uint64(38)) & np.uint64(31)).astype(np.int64), "ty0": ((k >> np.uint64(33)) & np.uint64(31)).astype(np.int64),
                "ty1": ((k >> np.uint64(28)) & np.uint64(31)).astype(np.int64), "first": first[:m].astype(np.int64),
                "cnt": cnt[:m].astype(np.int64), "tile_mask": mask}

    def bvh_order(self):
        """Triangle id (file order) at every position of the BVH-ordered triangle arrays."""
        out = np.zeros(self.T, dtype=np.int32)
        _lib.check(self._L.ipp_bvh_order(self._h, out.ctypes.data_as(_lib.c_i32p)))
        return out

    def bvh_selftest_ex(self, rays):
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        ids = np.zeros(r.shape[0], dtype=np.int32)
        brute = np.zeros(r.shape[0], dtype=np.int32)
        t = np.zeros((r.shape[0], 2), dtype=np.float32)
        mism = C.c_int64()
        _lib.check(self._L.ipp_bvh_selftest_ex(self._h, r.ctypes.data_as(_lib.c_f32p), r.shape[0], ids.ctypes.data_as(_lib.c_i32p),
                                               brute.ctypes.data_as(_lib.c_i32p), t.ctypes.data_as(_lib.c_f32p), C.byref(mism)))
        return ids, brute, t, int(mism.value)

    # ---- plumbing used by bench.py ----
    def sync(self):
        _lib.check(self._L.ipp_sync(self._h))

    def counters(self, reset=False):
        out = np.zeros(8, dtype=np.int64)
        _lib.check(self._L.ipp_counters(self._h, out.ctypes.data_as(_lib.c_i64p), int(reset)))
        keys = ["launches", "rays", "snapshots", "edges_rendered", "collision_queries", "reduce_launches", "plans_reduced", "snapshots_rendered"]
        return dict(zip(keys, (int(x) for x in out)))

    def visibility_stats(self, reset=False):
        out = np.zeros(8, dtype=np.int64)
        _lib.check(self._L.ipp_visibility_stats(self._h, out.ctypes.data_as(_lib.c_i64p), int(reset)))
        return dict(zip(["nodes", "triangles", "tiles", "rays", "live_triangles", "open_tiles", "supertiles", "supertile_splits"],
                        (int(x) for x in out)))

    def visibility_variant(self, use_bins=None):
        """Selects (True / False) or queries (None) the edge-visibility kernel: leaf bins or BVH walk per tile."""
        out = C.c_int()
        _lib.check(self._L.ipp_visibility_variant(self._h, -1 if use_bins is None else int(bool(use_bins)), C.byref(out)))
        return "bins" if out.value else "bvh"

    def kernel_time(self, which, reset=False):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self._L.ipp_kernel_time(self._h, int(which), C.byref(ms), C.byref(n), int(reset)))
        return ms.value, int(n.value)
