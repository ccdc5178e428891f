#!/usr/bin/env python
"""bench.py -- plan-evaluations/s of the B200 plan-fitness evaluator.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline (value, e2e, roofline): BASELINE.json config 2 -- mockup_only (stand-in for the missing 3d_models/mockup.stl), 1 000 plans x
20 viewpoints per GPU, sensor [2.0, 46, 46, 1024, 0.1, 10, 1, 1], genes uniform on [0, N) without consecutive duplicates, energy limit
disabled (with the limit on, every 20-viewpoint random plan exceeds the circling-plan energy and coverage would never be computed).
A step = one generation's fitness evaluation of a population the evaluator has never seen: the edge memo is cleared and L2 is
flushed at the start of every step, so every step does the full work -- collision prisms, energy, edge visibility for every
distinct plan edge (two 1024x1024 item-buffer snapshots per 2 m of path), bitmap packing, plan reduce.  Nothing is cached across
steps.  `value` times the steps with the population already resident in HBM; `e2e` times the same steps through the host-pointer
C-ABI call ipp_evaluate_population (H2D of the genes, D2H of the fitness rows inside the timed region).

The same JSON line carries a `c3` object: BASELINE.json config 3, the north star's target configuration -- luis4 (stand-in for the
missing oil_manifold.stl), ONE population of 100 000 plans x 30 viewpoints, sharded over the N ranks (strong scaling; the edges
of the whole population are rendered once across the ranks and their bitmap rows all-gathered): one cold step, warm steps, and an
NSGA-II generation mix with the energy gate on (optimize_coverage.run_nsga2, Python time included).
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
# config 3 (north star): one population, sharded over the ranks
C3_MESH = os.path.join(ROOT, "tests", "golden", "meshes", "luis4.ippm")
C3_POP, C3_VIEWPOINTS, C3_WARM_STEPS = 100000, 30, 5
C3_NSGA_POP, C3_NSGA_GENERATIONS = 1000, 20


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


def _oracle_memo_leg(args):
    """one plan prefix evaluated twice by one oracle: with memoisation the second evaluation is a lookup (evaluateMemoisedPlan,
    PlanCoverageEstimator.cpp:43-118), without it the edges are rendered again (evaluatePlanInternal :128-151)"""
    plan, memo = args
    from oracle.oracle import Oracle
    o = Oracle(MESH, SPECS)
    t = time.perf_counter()
    o.evaluatePopulation([plan], memo, True)
    o.evaluatePopulation([plan], memo, True)
    return time.perf_counter() - t


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline_legs(n_boxes, cores, budget_s=24.0):
    """The CPU legs BASELINE.md section 3 lists, on bounded samples of the config 2 workload (plan prefixes; a plan's cost is linear
    in its edges, rates are scaled to full 20-viewpoint plans and the sample says so): all host cores (one oracle per process
    over disjoint plans), one thread (the faithful analogue of the reference's serial toolbox.map, optimize_coverage.py:141), and
    one thread evaluating the same plans twice with the edge memo on and off."""
    import multiprocessing as mp
    from oracle import oracle as om
    om.build()
    ctx = mp.get_context("spawn")
    edge_s = 0.35  # one plan edge costs the oracle ~0.25-0.35 s of one core at 1024 x 1024 on this mesh
    vp_all = int(max(3, min(VIEWPOINTS, 1 + budget_s * 0.45 / edge_s)))
    vp_one = int(max(3, min(VIEWPOINTS, 1 + budget_s * 0.2 / edge_s)))
    vp_memo = int(max(3, min(VIEWPOINTS, 1 + budget_s * 0.1 / edge_s)))
    with ctx.Pool(cores) as pool:
        r_all, px_all, wall_all = cpu_reference_rate(n_boxes, cores, vp_all, SEED + 3000, pool)
    with ctx.Pool(1) as pool:
        r_one, px_one, wall_one = cpu_reference_rate(n_boxes, 1, vp_one, SEED + 3001, pool)
        ids, _ = make_population(n_boxes, 1, vp_memo, SEED + 3002)
        plan = [[int(x)] for x in ids]
        t_on = pool.map(_oracle_memo_leg, [(plan, True)])[0]
        t_off = pool.map(_oracle_memo_leg, [(plan, False)])[0]
    evals = 2 * (vp_memo - 1) / float(VIEWPOINTS - 1)  # two evaluations of the prefix, in units of full plans
    return {"value": r_all, "unit": UNIT, "cores": cores, "nproc": os.cpu_count(), "kind": "port", "rays_per_s": px_all,
            "sample": "SCALED PREFIX SAMPLE, memo on: %d plan prefixes of %d viewpoints, one per process, %.1f s wall, scaled to %d-viewpoint "
                      "plans" % (cores, vp_all, wall_all, VIEWPOINTS),
            "legs": {"all_cores_memo_on": {"value": r_all, "cores": cores, "rays_per_s": px_all},
                     "one_thread_memo_on": {"value": r_one, "cores": 1, "rays_per_s": px_one,
                                            "sample": "1 plan prefix of %d viewpoints, %.1f s, scaled" % (vp_one, wall_one)},
                     "one_thread_same_plan_twice_memo_on": {"value": evals / t_on, "cores": 1},
                     "one_thread_same_plan_twice_memo_off": {"value": evals / t_off, "cores": 1,
                                                             "sample": "1 plan prefix of %d viewpoints evaluated twice, scaled" % vp_memo}},
            "note": "CPU restatement of the reference algorithm (software item buffer, oracle/); the reference's OpenSceneGraph/OpenGL "
                    "build cannot be built in this image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import oracle as om
    om.build()
    cores = host_cores()
    o = om.Oracle(MESH, [2.0, 46, 46, 8, 0.1, 10, 1, 1])
    n_boxes = o.getNumberOfBoxes()
    o.close()
    # bounded sample: the whole run should end within a few minutes -> budget per step; one plan edge costs ~0.35 s of one core
    total_steps = args.steps + args.warmup
    per_step_s = max(2.0, min(25.0, 200.0 / max(1, total_steps)))
    vp = int(max(2, min(VIEWPOINTS, 1 + per_step_s / 0.35)))
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
    sample = "SCALED PREFIX SAMPLE: %d plan prefixes of %d viewpoints (%d edges) per step, one per process, scaled to %d-viewpoint plans" % (
        cores, vp, cores * (vp - 1), VIEWPOINTS)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus), "note": "CPU restatement of the reference algorithm (software item "
                       "buffer); the reference's OpenSceneGraph/OpenGL build cannot be built in this image"},
            "rays_per_s": px / total,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "nproc": os.cpu_count(), "kind": "port", "sample": sample},
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

    def free(self):
        if self.ptr:
            self.ev._L.ipp_device_free(self.ev._h, self.ptr)
            self.ptr = None


def pinned_array(L, shape, dtype):
    from inspection_path_planning_b200 import _lib
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _lib.check(L.ipp_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def load_profile_facts():
    """What only a profiler can count -- warp instructions, L2 and DRAM bytes of the dominant kernel per 32x32 tile it rasterises --
    from the committed ncu capture of this kernel on this workload (profiles/traffic.json names the capture and the commit)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return {}


class Comm:
    """NCCL plumbing of the N > 1 runs: torch.distributed (gloo, CPU) only carries the 128-byte NCCL id; every data-path
    collective is NCCL inside libipp.so."""

    def __init__(self, L, rank, world):
        self.L, self.rank, self.world, self.id = L, rank, world, None
        if world > 1:
            from inspection_path_planning_b200 import _lib
            import torch.distributed as dist
            dist.init_process_group("gloo")
            idbuf = np.zeros(128, dtype=np.uint8)
            if rank == 0:
                _lib.check(L.ipp_comm_unique_id(idbuf.ctypes.data_as(_lib.c_u8p)))
            obj = [idbuf.tobytes()]
            dist.broadcast_object_list(obj, src=0)
            self.id = np.frombuffer(obj[0], dtype=np.uint8).copy()
            self.dist = dist

    def new_id(self):
        """a fresh NCCL id for a second evaluator (one communicator per evaluator)"""
        from inspection_path_planning_b200 import _lib
        idbuf = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            _lib.check(self.L.ipp_comm_unique_id(idbuf.ctypes.data_as(_lib.c_u8p)))
        obj = [idbuf.tobytes()]
        self.dist.broadcast_object_list(obj, src=0)
        return np.frombuffer(obj[0], dtype=np.uint8).copy()

    def attach(self, ev, idbuf):
        from inspection_path_planning_b200 import _lib
        if self.world <= 1:
            return
        # NCCL prints its version banner to stdout when NCCL_DEBUG is set in the environment; stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            _lib.check(self.L.ipp_comm_init(ev._h, self.rank, self.world, idbuf.ctypes.data_as(_lib.c_u8p)))
            _lib.check(self.L.ipp_comm_barrier(ev._h))  # first collective: lazy NCCL set-up (and its log lines) happens here
            ev.sync()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)


def run_c3(args, L, comm, local, peaks):
    """BASELINE.json config 3 -- the north star's target configuration: ONE population of C3_POP plans x 30 viewpoints on luis4,
    evaluated by all ranks together (strong scaling): the host-pointer sharded call renders every edge of the population once
    across the ranks, all-gathers the fresh bitmap rows, evaluates a contiguous shard per rank and all-gathers the fitness rows.
    Cold step (empty memo), warm steps (same population again: K6 is the whole step; its HBM roofline is taken here, at a size
    where it is not launch-bound), then an NSGA-II generation mix with the energy gate on."""
    from inspection_path_planning_b200 import _lib, cpp_binding
    from inspection_path_planning_b200 import optimize_coverage as oc
    rank, world = comm.rank, comm.world
    ev = cpp_binding.PlanCoverageEstimator(C3_MESH, SPECS, device=local)
    h = ev._h
    if world > 1:
        comm.attach(ev, comm.new_id())

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

    n_boxes, T = ev.getNumberOfBoxes(), ev.T
    pop = args.c3_pop
    ids, offsets = make_population(n_boxes, pop, C3_VIEWPOINTS, SEED + 3)
    h_ids = pinned_array(L, ids.shape, np.int32)
    h_off = pinned_array(L, offsets.shape, np.int64)
    h_fit = pinned_array(L, (pop, 4), np.float64)
    h_ids[:] = ids
    h_off[:] = offsets

    def evaluate(no_limit=True):
        if world > 1:
            _lib.check(L.ipp_evaluate_population_sharded(h, h_ids.ctypes.data_as(_lib.c_i32p), None, h_off.ctypes.data_as(_lib.c_i64p), pop, 1,
                                                         int(no_limit), h_fit.ctypes.data_as(_lib.c_f64p)))
        else:
            _lib.check(L.ipp_evaluate_population(h, h_ids.ctypes.data_as(_lib.c_i32p), None, h_off.ctypes.data_as(_lib.c_i64p), pop, 1,
                                                 int(no_limit), h_fit.ctypes.data_as(_lib.c_f64p), None))

    # ---- cold: every edge of the population is rendered (once across the ranks) ----
    ev.counters(reset=True)
    for k in range(6):
        ev.kernel_time(k, reset=True)
    barrier()
    t = time.perf_counter()
    evaluate()
    barrier()
    cold_s = allmax(time.perf_counter() - t)
    cnt = ev.counters()
    cold = {"s": cold_s, "plans_per_s": pop / cold_s, "edges_rendered_all_ranks": int(allsum(cnt["edges_rendered"])),
            "edges_rendered_this_rank": cnt["edges_rendered"], "camera_poses_all_ranks": int(allsum(cnt["snapshots"])),
            "camera_poses_rendered_all_ranks": int(allsum(cnt["snapshots_rendered"])),
            "rays_per_s": allsum(cnt["rays"]) / cold_s, "memo_rows_per_rank": ev.memo_size(),
            "kernel_ms_this_rank": {"edge_visibility": ev.kernel_time(1)[0], "leaf_tables": ev.kernel_time(5)[0],
                                    "edge_collision": ev.kernel_time(2)[0], "plan_reduce": ev.kernel_time(0)[0]}}
    cold_fit = np.array(h_fit)

    # ---- warm: the same population again, L2 flushed before every step ----
    ev.kernel_time(0, reset=True)
    warm_s = 0.0
    for _ in range(C3_WARM_STEPS):
        _lib.check(L.ipp_flush_l2(h))
        barrier()
        t = time.perf_counter()
        evaluate()
        barrier()
        warm_s += allmax(time.perf_counter() - t)
    same = bool(np.array_equal(np.asarray(h_fit), cold_fit))
    red_ms, red_n = ev.kernel_time(0)
    words = (T + 31) // 32
    b_plan = (C3_VIEWPOINTS - 1) * 4 * words + 4 * (C3_VIEWPOINTS + 1) + 32
    shard = (pop + world - 1) // world
    k6_s = red_ms / max(1, red_n) * 1e-3
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    warm = {"ms_per_step": warm_s / C3_WARM_STEPS * 1e3, "plans_per_s": pop * C3_WARM_STEPS / warm_s, "same_rows_as_cold": same,
            "note": "wall clock around the host-pointer call: H2D of %d B of genes, probe kernel, K6 on this rank's shard, all-gather of "
                    "the fitness rows, D2H of %d B" % (ids.nbytes + offsets.nbytes, pop * 32)}
    k6 = {"kernel": "k_plan_reduce_rows_g<8, 4, 5> (row-major streaming form, areas from global memory)", "bound": "hbm", "plans_per_launch": shard, "us_per_launch": k6_s * 1e6,
          "bytes_per_plan": b_plan, "achieved": shard * b_plan / k6_s / 1e9 if k6_s > 0 else 0.0, "peak": hbm, "unit": "GB/s",
          "frac": (shard * b_plan / k6_s / 1e9 / hbm) if k6_s > 0 else 0.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback",
          "row_pool_gb": ev.memo_size() * ((words + 3) // 4 * 16) / 1e9,
          "note": "config 3 warm step on this rank's shard of the population, %d memoised rows (the row pool exceeds L2 many times), L2 "
                  "flushed before each launch; algorithmic bytes = SURVEY 8(d) B_plan" % ev.memo_size()}

    # ---- NSGA-II generation mix (optimize_coverage.py:109-213): energy gate on, memo GC every generation, Python time included ----
    ev.clear_memo()

    class Through:
        """run_nsga2 sees the reference's evaluator API; evaluatePopulation goes through the sharded call when N > 1 (every rank
        runs the same seeded loop and holds the same populations)"""

        def __init__(self):
            self.gen_stats = []

        def getNumberOfBoxes(self):
            return ev.getNumberOfBoxes()

        def getMaxAllowedEnergy(self):
            return ev.getMaxAllowedEnergy()

        def getSimplePlans(self):
            return ev.getSimplePlans()

        def evaluatePopulation(self, plans, memo, no_limit):
            c0 = ev.counters()
            t0 = time.perf_counter()
            fit = ev.evaluatePopulationSharded(plans, memo, no_limit) if world > 1 else ev.evaluatePopulation(plans, memo, no_limit)
            dt = time.perf_counter() - t0
            c1 = ev.counters()
            uses = sum(max(0, len(p) - 1) for p in plans)
            self.gen_stats.append({"plans": len(plans), "edge_uses": uses, "edges_rendered_this_rank": c1["edges_rendered"] - c0["edges_rendered"],
                                   "call_ms": dt * 1e3})
            return fit

        def updateMemoisedEdges(self, pop_):
            return ev.updateMemoisedEdges(pop_)

    # the reference's shipped run first (settings/Parameters.py:7-8: 40 individuals, 400 generations -- "a full optimize_coverage.py run")
    thr0 = Through()
    params0 = oc.Parameters()
    barrier()
    t = time.perf_counter()
    pop0, pf0, _, stats0 = oc.run_nsga2(thr0, params0, seed=SEED)
    barrier()
    run0_s = allmax(time.perf_counter() - t)
    call_ms = sorted(g["call_ms"] for g in thr0.gen_stats)
    full_run = {"pop": params0.NUM_INDIVS, "generations": params0.NUM_GENERATIONS, "wall_s": run0_s, "evaluations": stats0["evaluations"],
                "batched_calls": stats0["calls"], "evaluator_s": stats0["eval_s"], "median_call_ms": call_ms[len(call_ms) // 2],
                "p90_call_ms": call_ms[int(0.9 * (len(call_ms) - 1))], "pareto_front_size": len(pf0),
                "best_coverage_score": min(i.fitness[0] for i in pop0), "memoised_edges_at_end": stats0["memoised_edges"],
                "note": "optimize_coverage.run_nsga2 with the reference's shipped parameters on this mesh, seeded, energy gate on, memo GC every "
                        "generation; wall clock of the whole loop, Python included (evaluator set-up not included)"}
    ev.clear_memo()
    thr = Through()
    params = oc.Parameters(NUM_INDIVS=args.c3_nsga_pop, NUM_GENERATIONS=C3_NSGA_GENERATIONS)
    barrier()
    t = time.perf_counter()
    popn, pf, logbook, stats = oc.run_nsga2(thr, params, seed=SEED)
    barrier()
    nsga_s = allmax(time.perf_counter() - t)
    per_gen = []
    for g in thr.gen_stats:
        rendered = allsum(g["edges_rendered_this_rank"])
        per_gen.append({"plans": g["plans"], "new_edge_fraction": (rendered / g["edge_uses"]) if g["edge_uses"] else 0.0,
                        "call_ms": g["call_ms"]})
    nsga = {"pop": args.c3_nsga_pop, "generations": C3_NSGA_GENERATIONS, "evaluations": stats["evaluations"], "batched_calls": stats["calls"],
            "wall_s": nsga_s, "evals_per_s_incl_python": stats["evaluations"] / nsga_s, "evaluator_s": stats["eval_s"],
            "evals_per_s_evaluator_only": stats["evaluations"] / max(stats["eval_s"], 1e-9),
            "energy_gate": "on (maxAllowedEnergy %.3f)" % ev.getMaxAllowedEnergy(), "seeded": True,
            "pareto_front_size": len(pf), "best_coverage_score": min(i.fitness[0] for i in popn),
            "memoised_edges_at_end": stats["memoised_edges"],
            "calls": per_gen[:3] + per_gen[-(C3_NSGA_GENERATIONS):] if len(per_gen) > C3_NSGA_GENERATIONS + 3 else per_gen,
            "note": "calls = the seed plans, the rejection-sampled random individuals of the initial population (several calls), then one "
                    "call per generation (the invalid offspring: ~10 % crossover + 10 % mutation); new_edge_fraction = edges rendered / "
                    "edge uses of the call"}
    ev.close()
    return {"workload": "luis4.obj (stand-in for oil_manifold.stl, T = %d, %d candidate viewpoints), ONE population of %d plans x %d viewpoints "
                        "sharded over %d rank(s), sensor 1024x1024 46deg, front+down cameras, energy limit disabled for the cold/warm steps"
                        % (T, n_boxes, pop, C3_VIEWPOINTS, world),
            "scaling": "strong", "n_gpus": world, "cold": cold, "warm": warm, "roofline_plan_reduce": k6, "nsga2": nsga, "nsga2_reference_defaults": full_run}


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
    comm = Comm(L, rank, world)
    if world > 1:
        comm.attach(ev, comm.id)

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

    # ---- warm edges on config 2: launch-bound at 1 000 plans (the K6 roofline is taken on config 3) ----
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
                "achieved_gbs": POP * b_plan / per_launch_s / 1e9, "note": "config 2 with a warm edge memo: 1 000 plans are launch/ramp-bound; see "
                "c3.roofline_plan_reduce for the kernel at a size where it is HBM-bound"}

    # ---- measured L2 / HBM read peaks of this box (plain 16-byte loads, all SMs) ----
    l2_gbs, hbm_read_gbs = C.c_double(), C.c_double()
    _lib.check(L.ipp_read_bandwidth(h, 32 << 20, 40, C.byref(l2_gbs)))
    _lib.check(L.ipp_read_bandwidth(h, 256 << 20, 8, C.byref(hbm_read_gbs)))
    sm_count = 148

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass

    for b in d_ids:
        b.free()
    d_off.free()
    d_fit.free()
    ev.close()  # the row pool of this evaluator (70 % of the free memory) goes back before config 3 makes its own

    c3 = None
    if not args.no_c3:
        c3 = run_c3(args, L, comm, local, peaks)

    if rank != 0:
        return 0

    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    facts = load_profile_facts()
    kname = "k_visibility_bins" if variant == "bins" else "k_visibility"
    # Device-counted work of exactly the timed launches: tiles rasterised, table / list bytes scanned, triangles set up or staged.
    tiles_launch = vstat["tiles"] / max(1, vis_n)
    alg_bytes_launch = (64.0 * vstat["nodes"] + 48.0 * vstat["triangles"]) / max(1, vis_n)
    rays_launch = cnt["rays"] / max(1, vis_n)
    vis_launch_s = vis_ms / max(1, vis_n) * 1e-3
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = sm_count * 4 * sm_hz * 1e6  # warp instructions / s: 4 schedulers per SM, one instruction per cycle each
    inst_per_tile = facts.get(kname + "_warp_inst_per_tile")
    lts_per_tile = facts.get(kname + "_lts_bytes_per_tile")
    dram_per_tile = facts.get(kname + "_dram_bytes_per_tile")
    inst_launch = inst_per_tile * tiles_launch if inst_per_tile else None
    achieved_inst = inst_launch / vis_launch_s if inst_launch and vis_launch_s > 0 else None
    pop_total = world * POP
    value = pop_total * args.steps / (ms_total * 1e-3)
    e2e_value = pop_total * args.steps / (ms_e2e * 1e-3)
    roofline = {
        "kernel": kname, "bound": "issue",
        "achieved": achieved_inst / 1e12 if achieved_inst else None, "peak": issue_peak / 1e12, "unit": "T warp-instructions/s",
        "frac": (achieved_inst / issue_peak) if achieved_inst else None,
        "traffic": dram_per_tile * tiles_launch if dram_per_tile else None,
        "us_per_launch": vis_launch_s * 1e6, "tiles_per_launch": tiles_launch, "rays_per_launch": rays_launch,
        "warp_instructions_per_launch": inst_launch, "warp_instructions_per_tile": inst_per_tile,
        "peak_source": "148 SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % sm_hz,
        "l2": {"achieved": (lts_per_tile * tiles_launch / vis_launch_s / 1e9) if lts_per_tile and vis_launch_s > 0 else None,
               "peak": l2_gbs.value, "unit": "GB/s",
               "frac": (lts_per_tile * tiles_launch / vis_launch_s / 1e9 / l2_gbs.value) if lts_per_tile and vis_launch_s > 0 and l2_gbs.value > 0 else None,
               "peak_source": "measured in this run: 32 MiB read 40 times with plain 16-byte loads by 8 CTAs per SM (ipp_read_bandwidth)"},
        "hbm": {"achieved": (dram_per_tile * tiles_launch / vis_launch_s / 1e9) if dram_per_tile and vis_launch_s > 0 else None,
                "algorithmic_gbs": alg_bytes_launch / vis_launch_s / 1e9 if vis_launch_s > 0 else 0.0, "peak": hbm, "unit": "GB/s",
                "frac": (dram_per_tile * tiles_launch / vis_launch_s / 1e9 / hbm) if dram_per_tile and vis_launch_s > 0 else None,
                "peak_source": peak_src, "hbm_read_gbs_this_run": hbm_read_gbs.value,
                "algorithmic_bytes_per_launch": alg_bytes_launch,
                "note": "algorithmic bytes = 8 B per cluster-table entry and 4 B per tile-list entry scanned + 48 B per triangle set up or "
                        "staged (device counters of the timed launches); they are served by L1/L2"},
        "profile_source": facts.get("source"),
        "note": "K4 is bound by instruction issue (ncu capture named in profile_source: %.1f %% of the issue slots busy at %.0f %% of the warp "
                "slots, L2 at ~11 %% and DRAM at ~4 %% of their peaks), not by HBM: frac = warp instructions per second / (148 x 4 x SM "
                "clock).  Warp instructions, L2 bytes and DRAM bytes per rasterised tile come from that committed capture of this kernel on "
                "this workload; tiles per launch and the launch time are measured live (device counter, CUDA events on the evaluator's "
                "stream)." % (facts.get(kname + "_issue_active_pct", float("nan")), facts.get(kname + "_warps_active_pct", float("nan")))}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(world), "triangles": T, "candidate_viewpoints": n_boxes, "plans_per_gpu": POP,
                   "viewpoints_per_plan": VIEWPOINTS, "l2": "flushed every step (256 MiB write)", "seed": SEED,
                   "dtype_note": "visibility in f32 with f64 camera bases and f64-derived normals; energy, collision prisms and area sums in f64"},
        "rays_per_s": rays_total / (ms_total * 1e-3),
        "rays_note": "pixel-centre rays of the 32x32 tiles actually rasterised; tiles the cluster tables prove empty issue none",
        "pixels_resolved_per_s": allsum_snap * SPECS[3] * SPECS[3] / (ms_total * 1e-3),
        "snapshots_per_step": cnt["snapshots"] / args.steps, "snapshots_rendered_per_step": cnt["snapshots_rendered"] / args.steps,
        "snapshots_note": "camera poses of the edges rendered (what the reference renders) / poses K4 actually rendered: bit-identical poses that "
                          "several edges of a batch share are rendered once (pose_dedupe.cu)",
        "edges_rendered_per_step": cnt["edges_rendered"] / args.steps,
        "kernel_ms_per_step": {"edge_visibility": vis_ms / args.steps, "leaf_tables": tab_ms / args.steps, "plan_reduce": red_ms / args.steps,
                               "edge_collision": col_ms / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pop_total * VIEWPOINTS * 4 + (pop_total + 1) * 8),
                "d2h_bytes_per_step": int(pop_total * 4 * 8), "ms_per_step": ms_e2e / args.steps, "matches_resident_path": bool(same),
                **({"note": "host entry point: every rank passes all ranks' plans as one population; edges two ranks' plans share are "
                            "rendered once across the ranks and their bitmap rows all-gathered (DESIGN.md section 9)"} if world > 1 else {})},
        "gpu_launches": int(cnt["launches"]),
        "roofline": roofline,
        "roofline_plan_reduce": (c3 or {}).get("roofline_plan_reduce") or warm,
        "plan_reduce_config2_warm": warm,
        "c3": c3,
    }
    # CPU baseline beside it: the CPU restatement on this box's host cores, bounded samples (rank 0 at N = 1 only)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_legs(n_boxes, host_cores())
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ipp", choices=["ipp", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the config 3 object (quick A/B runs of the headline)")
    ap.add_argument("--c3-pop", type=int, default=C3_POP)
    ap.add_argument("--c3-nsga-pop", type=int, default=C3_NSGA_POP)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
