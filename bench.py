#!/usr/bin/env python
"""bench.py -- plan-evaluations/s of the B200 plan-fitness evaluator on BASELINE.json's config 2.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): mockup_only (stand-in for the missing 3d_models/mockup.stl), 1 000 plans x 20 viewpoints per GPU,
sensor [2.0, 46, 46, 1024, 0.1, 10, 1, 1], genes uniform on [0, N) without consecutive duplicates, energy limit disabled
(with the limit on, every 20-viewpoint random plan exceeds the circling-plan energy and coverage would never be computed).

A step = one generation's fitness evaluation of a population the evaluator has never seen: the edge memo is cleared and L2 is
flushed at the start of every step, so every step does the full work -- collision prisms, energy, edge visibility for every
distinct plan edge (two 1024x1024 item-buffer snapshots per 2 m of path), bitmap packing, plan reduce.  Nothing is cached
across steps.  `value` times the steps with the population already resident in HBM; `e2e` times the same steps through the
host-pointer C-ABI call ipp_evaluate_population (H2D of the genes, D2H of the fitness rows inside the timed region).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MESH = os.path.join(ROOT, "tests", "golden", "meshes", "mockup_only.ippm")
SPECS = [2.0, 46, 46, 1024, 0.1, 10, 1, 1]  # plan_optimizer/settings/Parameters.py:55
POP, VIEWPOINTS, SEED = 1000, 20, 20261019
METRIC, UNIT = "plan_evals_per_s", "plans/s"


def workload_name(n_gpus):
    return ("mockup_only.stl (stand-in for mockup.stl), %d plans x %d viewpoints per GPU, sensor 1024x1024 46deg, "
            "front+down cameras, 2 m sampling, energy limit disabled, edge memo cleared every step" % (POP, VIEWPOINTS))


def make_population(n_boxes, pop, viewpoints, seed):
    """genes i.i.d. uniform on [0, N), consecutive duplicates re-drawn (cleanUpIndividual,
    plan_optimizer/evolutionary_operators/mutationAndRecombination.py:64-87) -> CSR (ids int32, offsets int64)"""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, n_boxes, size=(pop, viewpoints))
    for k in range(1, viewpoints):
        dup = ids[:, k] == ids[:, k - 1]
        while dup.any():
            ids[dup, k] = rng.integers(0, n_boxes, size=int(dup.sum()))
            dup = ids[:, k] == ids[:, k - 1]
    offsets = np.arange(pop + 1, dtype=np.int64) * viewpoints
    return np.ascontiguousarray(ids.reshape(-1).astype(np.int32)), offsets


# ----------------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe), sampled during the timed region
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "250"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the CPU restatement of the reference algorithm (oracle/), all host cores
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    plan, = args
    from oracle.oracle import Oracle
    o = Oracle(MESH, SPECS)
    t = time.perf_counter()
    fit = o.evaluatePopulation([plan], True, True)
    dt = time.perf_counter() - t
    c = o.counters()
    return dt, c["pixels"], c["snapshots"], fit[0].tolist()


def cpu_reference_rate(n_boxes, cores, viewpoints, seed, pool):
    """one bounded sample: `cores` plan prefixes of `viewpoints` viewpoints, one per process.  Returns the rate in units of
    full VIEWPOINTS-viewpoint plans per second (a plan's cost is linear in its edges) and the pixel-ray rate."""
    ids, _ = make_population(n_boxes, cores, viewpoints, seed)
    plans = [[[int(x)] for x in ids[i * viewpoints:(i + 1) * viewpoints]] for i in range(cores)]
    t = time.perf_counter()
    res = pool.map(_oracle_worker, [(p,) for p in plans])
    wall = time.perf_counter() - t
    edges = cores * (viewpoints - 1)
    plans_equiv = edges / float(VIEWPOINTS - 1)
    pixels = sum(r[1] for r in res)
    return plans_equiv / wall, pixels / wall, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import oracle as om
    om.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    o = om.Oracle(MESH, [2.0, 46, 46, 8, 0.1, 10, 1, 1])
    n_boxes = o.getNumberOfBoxes()
    o.close()
    # bounded sample: the whole run should end within a few minutes -> budget per step; one plan edge costs ~1.1 s of one core
    total_steps = args.steps + args.warmup
    per_step_s = max(4.0, min(25.0, 200.0 / max(1, total_steps)))
    vp = int(max(2, min(VIEWPOINTS, 1 + per_step_s / 1.1)))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_reference_rate(n_boxes, cores, vp, SEED + 1000 + w, pool)
        t0 = time.perf_counter()
        pl = px = 0.0
        for k in range(args.steps):
            r, p, wall = cpu_reference_rate(n_boxes, cores, vp, SEED + 2000 + k, pool)
            pl += r * wall
            px += p * wall
        total = time.perf_counter() - t0
    value = pl / total
    sample = "%d plan prefixes of %d viewpoints (%d edges) per step, one per process, scaled to %d-viewpoint plans" % (
        cores, vp, cores * (vp - 1), VIEWPOINTS)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus), "note": "CPU restatement of the reference algorithm (software item "
                       "buffer); the reference's OpenSceneGraph/OpenGL build cannot be built in this image"},
            "rays_per_s": px / total,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class DeviceBuf:
    def __init__(self, ev, nbytes):
        self.ev, self.ptr = ev, C.c_void_p()
        from inspection_path_planning_b200 import _lib
        _lib.check(ev._L.ipp_device_alloc(ev._h, int(nbytes), C.byref(self.ptr)))

    def upload(self, arr):
        from inspection_path_planning_b200 import _lib
        _lib.check(self.ev._L.ipp_memcpy_h2d(self.ev._h, self.ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def download(self, arr):
        from inspection_path_planning_b200 import _lib
        _lib.check(self.ev._L.ipp_memcpy_d2h(self.ev._h, arr.ctypes.data_as(C.c_void_p), self.ptr, arr.nbytes))


def pinned_array(L, shape, dtype):
    from inspection_path_planning_b200 import _lib
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(L.ipp_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def run_gpu(args):
    from inspection_path_planning_b200 import _lib, cpp_binding
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
    L = _lib.load()
    ev = cpp_binding.PlanCoverageEstimator(MESH, SPECS, device=local)
    h = ev._h
    n_boxes = ev.getNumberOfBoxes()
    T = ev.T

    dist = None
    if world > 1:
        # torch.distributed (gloo, CPU) only carries the 128-byte NCCL id; every data-path collective is NCCL inside libipp.so
        import torch.distributed as dist
        dist.init_process_group("gloo")
        idbuf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(L.ipp_comm_unique_id(idbuf.ctypes.data_as(_lib.c_u8p)))
        obj = [idbuf.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        idbuf = np.frombuffer(obj[0], dtype=np.uint8).copy()
        # NCCL prints its version banner to stdout when NCCL_DEBUG is set in the environment; stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            _lib.check(L.ipp_comm_init(h, rank, world, idbuf.ctypes.data_as(_lib.c_u8p)))
            _lib.check(L.ipp_comm_barrier(h))  # first collective: lazy NCCL set-up (and its log lines) happens here
            ev.sync()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            _lib.check(L.ipp_comm_barrier(h))
        ev.sync()

    def allmax(x):
        a = np.asarray([x], dtype=np.float64)
        if world > 1:
            _lib.check(L.ipp_comm_allreduce_max(h, a.ctypes.data_as(_lib.c_f64p), 1))
        return float(a[0])

    def allsum(x):
        a = np.asarray([x], dtype=np.float64)
        if world > 1:
            _lib.check(L.ipp_comm_allreduce_sum(h, a.ctypes.data_as(_lib.c_f64p), 1))
        return float(a[0])

    # ---- populations: one never-seen population per step and rank, resident in HBM before the timed region ----
    n_steps_total = args.warmup + args.steps
    pops = [make_population(n_boxes, POP, VIEWPOINTS, SEED + 7919 * s + 104729 * rank) for s in range(n_steps_total)]
    n_genes = POP * VIEWPOINTS
    d_ids = [DeviceBuf(ev, n_genes * 4) for _ in pops]
    d_off = DeviceBuf(ev, (POP + 1) * 8)
    d_off.upload(pops[0][1])
    for b, (ids, _) in zip(d_ids, pops):
        b.upload(ids)
    d_fit = DeviceBuf(ev, world * POP * 4 * 8)
    ev.sync()

    def step_resident(s):
        _lib.check(L.ipp_flush_l2(h))
        _lib.check(L.ipp_clear_memo(h))
        if world > 1:
            _lib.check(L.ipp_evaluate_population_sharded_device(h, d_ids[s].ptr, None, d_off.ptr, POP, n_genes, POP, 1, 1, d_fit.ptr))
        else:
            _lib.check(L.ipp_evaluate_population_device(h, d_ids[s].ptr, None, d_off.ptr, POP, n_genes, 1, 1, d_fit.ptr))

    for s in range(args.warmup):
        step_resident(s)
    barrier()
    ev.counters(reset=True)
    for k in range(6):
        ev.kernel_time(k, reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    _lib.check(L.ipp_timer_start(h))
    for s in range(args.warmup, n_steps_total):
        step_resident(s)
    ms = C.c_double()
    _lib.check(L.ipp_timer_stop(h, C.byref(ms)))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = allmax(ms.value)
    cnt = ev.counters()
    vstat = ev.visibility_stats()
    vis_ms, vis_n = ev.kernel_time(1)
    red_ms, red_n = ev.kernel_time(0)
    col_ms, col_n = ev.kernel_time(2)
    tab_ms, tab_n = ev.kernel_time(5)
    variant = ev.visibility_variant()
    rays_total = allsum(cnt["rays"])
    allsum_snap = allsum(cnt["snapshots"])
    fit_resident = np.zeros((world * POP, 4))
    d_fit.download(fit_resident)

    # ---- e2e: the same steps through the host-pointer C-ABI call, genes in pinned host memory ----
    h_ids = [pinned_array(L, (n_genes,), np.int32) for _ in pops]
    h_off = pinned_array(L, (POP + 1,), np.int64)
    h_off[:] = pops[0][1]
    for a, (ids, _) in zip(h_ids, pops):
        a[:] = ids
    h_fit = pinned_array(L, (world * POP, 4), np.float64)
    g_off = g_ids = None
    if world > 1:
        # the sharded host entry point takes the whole population on every rank; rank r evaluates rows [r*POP, (r+1)*POP)
        g_off = pinned_array(L, (world * POP + 1,), np.int64)
        g_off[:] = np.arange(world * POP + 1, dtype=np.int64) * VIEWPOINTS
        g_ids = []
        for s in range(n_steps_total):
            a = pinned_array(L, (world * n_genes,), np.int32)
            for r in range(world):
                a[r * n_genes:(r + 1) * n_genes] = make_population(n_boxes, POP, VIEWPOINTS, SEED + 7919 * s + 104729 * r)[0]
            g_ids.append(a)

    def step_e2e(s):
        _lib.check(L.ipp_flush_l2(h))
        _lib.check(L.ipp_clear_memo(h))
        if world > 1:
            _lib.check(L.ipp_evaluate_population_sharded(h, g_ids[s].ctypes.data_as(_lib.c_i32p), None, g_off.ctypes.data_as(_lib.c_i64p),
                                                         world * POP, 1, 1, h_fit.ctypes.data_as(_lib.c_f64p)))
        else:
            _lib.check(L.ipp_evaluate_population(h, h_ids[s].ctypes.data_as(_lib.c_i32p), None, h_off.ctypes.data_as(_lib.c_i64p), POP, 1, 1,
                                                 h_fit.ctypes.data_as(_lib.c_f64p), None))

    step_e2e(0)  # one untimed pass of the host path (buffers of the host entry point get allocated here)
    barrier()
    _lib.check(L.ipp_timer_start(h))
    for s in range(args.warmup, n_steps_total):
        step_e2e(s)
    ms2 = C.c_double()
    _lib.check(L.ipp_timer_stop(h, C.byref(ms2)))
    barrier()
    ms_e2e = allmax(ms2.value)
    # the host path and the resident path must agree on the last step's fitness rows
    same = np.array_equal(np.asarray(h_fit), fit_resident)

    # ---- warm edges: the same population evaluated again (memoised edges; plan_reduce is the whole step) ----
    warm = None
    if world == 1:
        s = n_steps_total - 1
        _lib.check(L.ipp_evaluate_population_device(h, d_ids[s].ptr, None, d_off.ptr, POP, n_genes, 1, 1, d_fit.ptr))
        ev.sync()
        ev.kernel_time(0, reset=True)
        reps = 20
        for _ in range(reps):
            _lib.check(L.ipp_flush_l2(h))
            _lib.check(L.ipp_evaluate_population_device(h, d_ids[s].ptr, None, d_off.ptr, POP, n_genes, 1, 1, d_fit.ptr))
        ev.sync()
        wms, wn = ev.kernel_time(0)
        words = (T + 31) // 32
        b_plan = (VIEWPOINTS - 1) * 4 * words + 4 * (VIEWPOINTS + 1) + 32
        per_launch_s = wms / max(1, wn) * 1e-3
        warm = {"kernel": "k_plan_reduce", "plans_per_launch": POP, "us_per_launch": per_launch_s * 1e6,
                "plan_evals_per_s_kernel": POP / per_launch_s, "bytes_per_plan": b_plan,
                "achieved_gbs": POP * b_plan / per_launch_s / 1e9, "note": "edge memo warm: not the headline; L2 flushed before each launch"}

    if rank != 0:
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # K4 reads 8 B per leaf-bin entry and 4 B per record header it scans (BVH variant: 64 B per node a tile packet fetches) and at
    # least 48 B per triangle it sets up (48 B of vertices in, 64 B of record out) or stages (64 B record in) -- 48 B is charged
    # for each; the counts are data dependent and come from the kernel's own device counters for exactly
    # the timed launches ("nodes" is in 64-byte units for either variant).
    alg_bytes_launch = (64.0 * vstat["nodes"] + 48.0 * vstat["triangles"]) / max(1, vis_n)
    rays_launch = cnt["rays"] / max(1, vis_n)
    vis_launch_s = vis_ms / max(1, vis_n) * 1e-3
    achieved = alg_bytes_launch / vis_launch_s / 1e9 if vis_launch_s > 0 else 0.0
    # SURVEY 8(d)'s independent-ray model, for reference: one root-to-leaf descent (2 x 32 B boxes per level) + one 48 B triangle per ray
    b_ray = 2 * math.ceil(math.log2(T)) * 32 + 48
    pop_total = world * POP
    value = pop_total * args.steps / (ms_total * 1e-3)
    e2e_value = pop_total * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(world), "triangles": T, "candidate_viewpoints": n_boxes, "plans_per_gpu": POP,
                   "viewpoints_per_plan": VIEWPOINTS, "l2": "flushed every step (256 MiB write)", "seed": SEED,
                   "dtype_note": "visibility in f32 with f64 camera bases and f64-derived normals; energy, collision prisms and area sums in f64"},
        "rays_per_s": rays_total / (ms_total * 1e-3),
        "rays_note": "pixel-centre rays of the 32x32 tiles actually rasterised; tiles the leaf tables prove empty issue none",
        "pixels_resolved_per_s": allsum_snap * SPECS[3] * SPECS[3] / (ms_total * 1e-3),
        "snapshots_per_step": cnt["snapshots"] / args.steps, "edges_rendered_per_step": cnt["edges_rendered"] / args.steps,
        "kernel_ms_per_step": {"edge_visibility": vis_ms / args.steps, "leaf_tables": tab_ms / args.steps, "plan_reduce": red_ms / args.steps,
                               "edge_collision": col_ms / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pop_total * VIEWPOINTS * 4 + (pop_total + 1) * 8),
                "d2h_bytes_per_step": int(pop_total * 4 * 8), "ms_per_step": ms_e2e / args.steps, "matches_resident_path": bool(same),
                **({"note": "host entry point: every rank passes all ranks' plans as one population; edges two ranks' plans share are "
                            "rendered once across the ranks and their bitmap rows all-gathered (DESIGN.md section 9)"} if world > 1 else {})},
        "gpu_launches": int(cnt["launches"]),
        "roofline": {"kernel": "k_visibility_bins" if variant == "bins" else "k_visibility", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": None, "algorithmic_bytes_per_launch": alg_bytes_launch, "us_per_launch": vis_launch_s * 1e6,
                     "tiles_per_launch": vstat["tiles"] / max(1, vis_n), "accel_64B_units_per_tile": vstat["nodes"] / max(1, vstat["tiles"]),
                     "triangles_per_tile": vstat["triangles"] / max(1, vstat["tiles"]), "rays_per_launch": rays_launch,
                     "peak_source": peak_src,
                     "survey_independent_ray_model": {"bytes_per_ray": b_ray, "gbs": rays_launch * b_ray / vis_launch_s / 1e9 if vis_launch_s > 0 else 0.0,
                                                      "note": "what one independent BVH descent per issued ray would have requested; a tile shares one leaf list"},
                     "note": "bytes = 8 B per leaf-bin entry / 4 B per record header scanned (64 B per BVH node in the bvh variant) + 48 B per triangle set up or staged, counted "
                             "on the device; triangles and leaf tables are served from L1/L2 (ncu DRAM traffic in profiles/), the kernel is "
                             "instruction-issue bound, not HBM bound"},
        "roofline_plan_reduce": warm,
    }
    # CPU baseline beside it: the CPU restatement on this box's host cores, bounded sample
    if not args.no_cpu_baseline:
        import multiprocessing as mp
        from oracle import oracle as om
        om.build()
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        vp = 12
        with mp.get_context("spawn").Pool(cores) as pool:
            r, px, wall = cpu_reference_rate(n_boxes, cores, vp, SEED + 3000, pool)
        line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port", "rays_per_s": px,
                                "sample": "%d plan prefixes of %d viewpoints, one per process, %.1f s wall, scaled to %d-viewpoint plans"
                                          % (cores, vp, wall, VIEWPOINTS)}
    if os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            line["roofline"]["traffic"] = tr.get(line["roofline"]["kernel"] + "_dram_bytes_per_launch")
            line["roofline"]["traffic_source"] = tr.get("source")
        except (OSError, ValueError):
            pass
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ipp", choices=["ipp", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
