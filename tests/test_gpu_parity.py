"""GPU parity tests: libipp.so (through its C ABI, via the ctypes mirror of cpp_binding) against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): collision flags, box classification, ids, seed plans bit-exact;
energy / path length within 1e-12 relative (same fp64 formulas, no FMA); per-edge seen sets may differ from the oracle in at
most 0.05 % of the triangles and only on triangles the oracle classifies as edge-epsilon (lenient minus strict); coverage
within 1e-5 relative."""
import numpy as np
import pytest

from conftest import bit_ids, mesh_path, popcount, random_plans, specs

pytestmark = pytest.mark.gpu

FLIP_BUDGET = 5e-4  # 0.05 % of the triangles
COVERAGE_RTOL = 1e-5
ENERGY_RTOL = 1e-12


def _pair(oracle_mod, mesh, height, **kw):
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(height, **{k: v for k, v in kw.items() if k in ("interval", "front", "down", "near", "far", "fov")})
    okw = {}
    gkw = {}
    if "start" in kw:
        okw["start_location"] = kw["start"]
        gkw["startLocation"] = kw["start"]
    if kw.get("loops"):
        okw["plan_loops_around"] = True
        gkw["planLoopsAround"] = True
    return cpp_binding.PlanCoverageEstimator(mesh_path(mesh), sp, **gkw), oracle_mod.Oracle(mesh_path(mesh), sp, **okw)


@pytest.fixture(scope="module")
def sphere(oracle_mod):
    g, o = _pair(oracle_mod, "sphere", 64)
    yield g, o
    g.close()


@pytest.fixture(scope="module")
def mockup256(oracle_mod):
    g, o = _pair(oracle_mod, "mockup_only", 256)
    yield g, o
    g.close()


@pytest.fixture(scope="module")
def luis1(oracle_mod):
    g, o = _pair(oracle_mod, "luis1", 128)
    yield g, o
    g.close()


# ------------------------------------------------------------------ setup-time facts
@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_scene_constants_bit_exact(fx, request):
    g, o = request.getfixturevalue(fx)
    assert g.T == o.T
    assert np.array_equal(g.areas(), o.areas())
    assert g.total_area() == o.total_area()
    assert g.collision_penalty() == o.collision_penalty()
    gc, gb = g.scene_info()
    oc, ob = o.scene_info()
    assert np.array_equal(gc, oc) and np.array_equal(gb, ob)
    assert g.image_size() == o.image_size()


@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_candidate_grid_and_seed_plans_bit_exact(fx, request):
    g, o = request.getfixturevalue(fx)
    assert np.array_equal(g.grid(), o.grid())
    assert g.getNumberOfBoxes() == o.getNumberOfBoxes()
    assert np.array_equal(g.box_centres(), o.box_centres())
    assert g.getSimplePlans() == o.getSimplePlans()
    assert g.getMaxAllowedEnergy() == pytest.approx(o.getMaxAllowedEnergy(), rel=ENERGY_RTOL)


@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_bvh_matches_brute_force(fx, request):
    g, _ = request.getfixturevalue(fx)
    rng = np.random.default_rng(5)
    c, bb = g.scene_info()
    n = 20000
    org = c + rng.normal(size=(n, 3)) * (bb[3:] - bb[:3]).max()
    tgt = bb[:3] + rng.random((n, 3)) * (bb[3:] - bb[:3])
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ids, mism = g.bvh_selftest(np.concatenate([org, d], axis=1))
    assert mism == 0
    assert (ids >= 0).mean() > 0.3


# ------------------------------------------------------------------ K5 / K2: exact predicates
@pytest.mark.parametrize("fx,n", [("sphere", 3000), ("mockup256", 10000), ("luis1", 4000)])
def test_edge_collision_flags_bit_exact(fx, n, request):
    g, o = request.getfixturevalue(fx)
    rng = np.random.default_rng(11)
    c = o.box_centres()
    a = c[rng.integers(0, len(c), n)]
    b = c[rng.integers(0, len(c), n)]
    keep = np.any(a != b, axis=1)
    a, b = a[keep], b[keep]
    # add near-miss edges: random points close to the structure
    _, bb = o.scene_info()
    m = n // 4
    a2 = bb[:3] - 2 + rng.random((m, 3)) * (bb[3:] - bb[:3] + 4)
    b2 = bb[:3] - 2 + rng.random((m, 3)) * (bb[3:] - bb[:3] + 4)
    a, b = np.concatenate([a, a2]), np.concatenate([b, b2])
    got = g.edges_collide(np.concatenate([a, b], axis=1))
    want = np.array([o.edge_collides(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want)
    assert 0.02 < want.mean() < 0.98


def test_vertical_and_sloped_edges_collision(mockup256):
    g, o = mockup256
    rng = np.random.default_rng(12)
    _, bb = o.scene_info()
    n = 2000
    a = bb[:3] - 1 + rng.random((n, 3)) * (bb[3:] - bb[:3] + 2)
    b = a.copy()
    b[:, 2] += rng.uniform(-4, 4, n)  # exactly vertical: up = (1, 0, 0) branch
    got = g.edges_collide(np.concatenate([a, b], axis=1))
    want = np.array([o.edge_collides(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want)


def test_box_classification_bit_exact(mockup256):
    g, o = mockup256
    rng = np.random.default_rng(13)
    _, bb = o.scene_info()
    pts = bb[:3] - 6 + rng.random((1500, 3)) * (bb[3:] - bb[:3] + 12)
    for half in (2.0, 10.0, 0.3):
        got = g.boxes_hit_mesh(pts, half)
        want = np.array([o.box_hits_mesh(p, half) for p in pts])
        assert np.array_equal(got, want), half


# ------------------------------------------------------------------ K4: visibility
def _poses(o, rng, n):
    c, bb = o.scene_info()
    ext = (bb[3:] - bb[:3])
    out = []
    for _ in range(n):
        eye = c + rng.normal(size=3) * ext * 0.8
        tgt = bb[:3] + rng.random(3) * ext
        up = np.array([0.0, 0.0, 1.0])
        out.append((eye.astype(np.float32).astype(np.float64), tgt, up))
    return out


@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_item_buffer_images(fx, request):
    g, o = request.getfixturevalue(fx)
    rng = np.random.default_rng(21)
    tot = bad = 0
    for eye, tgt, up in _poses(o, rng, 6):
        a = g.render_ids(eye, tgt, up)
        b = o.render_ids(eye, tgt, up)
        tot += a.size
        bad += int((a != b).sum())
    # pixel-level disagreements are confined to silhouettes / shared edges hit exactly at a pixel centre
    assert bad / tot < 2e-3, (bad, tot)


def _check_edge(g, o, prev, curr, stats):
    got = g.edge_bitmap(prev, curr)
    seen, strict, lenient = o.edge_bitmap(prev, curr, classify=True)
    flips = got ^ seen
    nflip = popcount(flips)
    stats["flips"] += nflip
    stats["edges"] += 1
    stats["seen"] += popcount(seen)
    # every disagreement must be an edge-epsilon triangle of the oracle: strict <= got <= lenient
    outside = popcount(strict & ~got) + popcount(got & ~lenient)
    stats["outside"] += outside
    return nflip, outside


@pytest.mark.parametrize("fx,n", [("sphere", 12), ("mockup256", 10), ("luis1", 10)])
def test_edge_bitmaps_within_flip_budget(fx, n, request):
    g, o = request.getfixturevalue(fx)
    rng = np.random.default_rng(31)
    N = o.getNumberOfBoxes()
    stats = dict(flips=0, edges=0, seen=0, outside=0)
    worst = 0
    for _ in range(n):
        prev, curr = (int(x) for x in rng.choice(N, 2, replace=False))
        nflip, _ = _check_edge(g, o, prev, curr, stats)
        worst = max(worst, nflip)
    assert stats["seen"] > 0
    assert worst <= max(1, int(FLIP_BUDGET * o.T)), stats
    assert stats["outside"] == 0, stats


def test_edge_bitmap_full_resolution(oracle_mod):
    g, o = _pair(oracle_mod, "mockup_only", 1024)
    try:
        rng = np.random.default_rng(32)
        N = o.getNumberOfBoxes()
        c = o.box_centres()
        stats = dict(flips=0, edges=0, seen=0, outside=0)
        k = 0
        while k < 3:
            prev, curr = (int(x) for x in rng.choice(N, 2, replace=False))
            if np.linalg.norm(c[prev] - c[curr]) > 4.5:
                continue
            nflip, outside = _check_edge(g, o, prev, curr, stats)
            assert nflip <= int(FLIP_BUDGET * o.T) and outside == 0, (prev, curr, stats)
            k += 1
        assert stats["seen"] > 100
    finally:
        g.close()


def _check_edges_pooled(g, mesh_file, sp, edges, T):
    """Per-edge seen sets of libipp.so against the oracle's (computed in a process pool) with the strict / lenient classification:
    at most 0.05 % of the triangles flip on any edge, and every flip is an edge-epsilon triangle of the oracle."""
    from _oracle_pool import classify_edges
    res = classify_edges(mesh_file, sp, edges)
    stats = dict(flips=0, edges=0, seen=0, outside=0, worst=0)
    seen_by_edge = {}
    for (a, b), (seen, strict, lenient) in zip(edges, res):
        got = g.edge_bitmap(a, b)
        nflip = popcount(got ^ seen)
        outside = popcount(strict & ~got) + popcount(got & ~lenient)
        stats["flips"] += nflip; stats["edges"] += 1; stats["seen"] += popcount(seen); stats["outside"] += outside
        stats["worst"] = max(stats["worst"], nflip)
        assert nflip <= max(1, int(FLIP_BUDGET * T)) and outside == 0, ((a, b), nflip, outside, stats)
        seen_by_edge[(a, b)] = seen
    return stats, seen_by_edge


def test_full_resolution_config2_edges_of_the_bench_population(oracle_mod):
    """BASELINE.json config 2 as bench.py runs it (mockup_only, 1024 x 1024, the first step's population of rank 0): 40 of its own
    edges -- all 19 of plan 0, the 10 longest of the population and random ones -- against the oracle, through the record path of
    K4 (asserted), and the fitness row of plan 0 against the coverage of the oracle's union."""
    import bench
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(1024)
    assert sp == bench.SPECS
    g = cpp_binding.PlanCoverageEstimator(mesh_path("mockup_only"), sp)
    o = oracle_mod.Oracle(mesh_path("mockup_only"), specs(16))  # geometry, areas and the exact-energy side only
    try:
        N = g.getNumberOfBoxes()
        ids, offsets = bench.make_population(N, bench.POP, bench.VIEWPOINTS, bench.SEED)
        plans = ids.reshape(bench.POP, bench.VIEWPOINTS)
        c = g.box_centres()
        pairs = np.stack([plans[:, :-1].ravel(), plans[:, 1:].ravel()], axis=1)
        length = np.linalg.norm(c[pairs[:, 0]] - c[pairs[:, 1]], axis=1)
        chosen = [tuple(int(x) for x in e) for e in zip(plans[0, :-1], plans[0, 1:])]
        for i in np.argsort(-length):
            e = (int(pairs[i, 0]), int(pairs[i, 1]))
            if e not in chosen:
                chosen.append(e)
            if len(chosen) >= 29:
                break
        rng = np.random.default_rng(2)
        while len(chosen) < 40:
            i = int(rng.integers(0, len(pairs)))
            e = (int(pairs[i, 0]), int(pairs[i, 1]))
            if e not in chosen:
                chosen.append(e)
        assert length.max() > 12.0  # long edges are in
        g.visibility_stats(reset=True)
        stats, seen = _check_edges_pooled(g, mesh_path("mockup_only"), sp, chosen, g.T)
        vs = g.visibility_stats()
        assert vs["supertiles"] > 0 and vs["supertile_splits"] <= 0.02 * vs["supertiles"], vs
        assert stats["seen"] > 20000, stats
        # plan 0: fitness row = area of the union of its edges' seen sets (energy gate off, as in the bench)
        plan0 = [[int(x)] for x in plans[0]]
        got = g.evaluatePopulation([plan0], True, True)
        union = np.zeros(g.W32, dtype=np.uint32)
        for e in zip(plans[0, :-1], plans[0, 1:]):
            union |= seen[(int(e[0]), int(e[1]))]
        want_cov = 1.0 - o.coverage_of_bitmap(union)
        assert abs((1.0 - got[0, 0]) - want_cov) <= max(COVERAGE_RTOL * want_cov, FLIP_BUDGET * o.areas().max() * g.T / o.total_area())
        e_want, pl_want, nc_want = o.energy_of_positions(c[plans[0]])
        assert got[0, 1] == pytest.approx(e_want, rel=ENERGY_RTOL) and got[0, 2] == pytest.approx(pl_want, rel=ENERGY_RTOL)
        assert got[0, 3] == nc_want
    finally:
        g.close()


def test_full_resolution_luis4_edges_take_the_record_path(oracle_mod):
    """The stand-in of config 3 (luis4, 70 226 triangles) at 1024 x 1024: 12 edges against the oracle, K4 through its cached set-up
    records (the path the benchmark numbers of that mesh are made on)."""
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(1024)
    g = cpp_binding.PlanCoverageEstimator(mesh_path("luis4"), sp)
    try:
        N = g.getNumberOfBoxes()
        rng = np.random.default_rng(33)
        c = g.box_centres()
        edges = []
        while len(edges) < 12:
            a, b = (int(x) for x in rng.choice(N, 2, replace=False))
            # (the oracle needs a second per camera pose on this mesh: keep the sampled edges below 16 m = 18 poses)
            if np.linalg.norm(c[a] - c[b]) <= 16.0 and (a, b) not in edges:
                edges.append((a, b))
        g.visibility_stats(reset=True)
        stats, _ = _check_edges_pooled(g, mesh_path("luis4"), sp, edges, g.T)
        vs = g.visibility_stats()
        assert vs["supertiles"] > 0 and vs["supertile_splits"] <= 0.02 * vs["supertiles"], vs
        assert stats["seen"] > 5000, stats
    finally:
        g.close()


def test_full_resolution_large_mesh_through_the_bvh_walk(oracle_mod, tmp_path):
    """A 327 680-triangle displaced icosphere (the generator of config 5, one subdivision level below it) at 1024 x 1024: 6 edges
    against the oracle through the BVH-walking variant of K4, and through whatever variant the evaluator picks by default."""
    from inspection_path_planning_b200 import cpp_binding, synth_meshes
    sp = specs(1024)
    path = str(tmp_path / "ico7.ippm")
    synth_meshes.write_ippm(path, synth_meshes.icosphere(level=7))
    g = cpp_binding.PlanCoverageEstimator(path, sp)
    try:
        assert g.T == 327680
        N = g.getNumberOfBoxes()
        rng = np.random.default_rng(34)
        c = g.box_centres()
        edges = []
        while len(edges) < 6:
            a, b = (int(x) for x in rng.choice(N, 2, replace=False))
            if np.linalg.norm(c[a] - c[b]) <= 12.0 and (a, b) not in edges:
                edges.append((a, b))
        default_variant = g.visibility_variant()
        from _oracle_pool import classify_edges
        res = classify_edges(path, sp, edges)
        for variant in ("bvh", default_variant):
            assert g.visibility_variant(variant == "bins") == variant
            seen_total = 0
            for (a, b), (seen, strict, lenient) in zip(edges, res):
                got = g.edge_bitmap(a, b)
                assert popcount(got ^ seen) <= int(FLIP_BUDGET * g.T), (variant, a, b, popcount(got ^ seen))
                assert popcount(strict & ~got) == 0 and popcount(got & ~lenient) == 0, (variant, a, b)
                seen_total += popcount(seen)
            assert seen_total > 10000
            if variant == default_variant == "bvh":
                break
    finally:
        g.close()


def test_single_camera_modes(oracle_mod):
    for front, down in ((1, 0), (0, 1)):
        g, o = _pair(oracle_mod, "sphere", 64, front=front, down=down)
        try:
            for prev, curr in ((100, 101), (250, 260), (30, 300)):
                got, want = g.edge_bitmap(prev, curr), o.edge_bitmap(prev, curr)
                assert popcount(got ^ want) <= 1
        finally:
            g.close()


# ------------------------------------------------------------------ plans
def _plan_edges(plan, has_start, loops):
    """[(prev, curr, angle offset of the edge)] of one plan: node sequence [start] g_0 .. g_{n-1} [g_0 if the plan loops around]
    (PlanCoverageEstimator.cpp:216-220), prev = -1 is the start location; the offset of an edge is its end gene's."""
    if not plan:
        return []
    genes = [(int(x[0]), float(x[1]) if len(x) > 1 else 0.0) for x in plan]
    nodes = ([(-1, 0.0)] if has_start else []) + genes + ([genes[0]] if loops else [])
    return [(a[0], b[0], b[1]) for a, b in zip(nodes[:-1], nodes[1:])]


def _compare_population(g, o, plans, memo=True, no_limit=False, got=None):
    """Fitness rows against the oracle.  Energy, path length, collision counts: exact formulas.  Coverage: within 1e-5 relative,
    except for plans whose seen set differs from the oracle's by edge-epsilon triangles -- there the north star's flip budget
    applies instead: at most 0.05 % of the triangles, every one of them inside the oracle's edge-epsilon band on the edge that
    shows it, and the coverage the evaluator reports must be exactly the area of the set it saw.  Works for [id] and
    [id, angleOffset] genes, start locations, loop closure and post-processing pairs (the oracle's per-edge classification then
    uses the heading the oracle's own search chose)."""
    if got is None:
        got = g.evaluatePopulation(plans, memo, no_limit)
    want = o.evaluatePopulation(plans, memo, no_limit)
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=ENERGY_RTOL, atol=0)
    np.testing.assert_allclose(got[:, 2], want[:, 2], rtol=ENERGY_RTOL, atol=0)
    assert np.array_equal(got[:, 3], want[:, 3])  # colliding-edge counts: bit-exact flags
    gated = want[:, 0] == 1.0
    assert np.array_equal(got[:, 0] == 1.0, gated) or np.allclose(got[:, 0], want[:, 0], rtol=COVERAGE_RTOL)
    close = np.isclose(1.0 - got[:, 0], 1.0 - want[:, 0], rtol=COVERAGE_RTOL, atol=1e-9)
    for p in np.nonzero(~close)[0]:
        fit1, ub = g.evaluatePopulation([plans[p]], memo, no_limit, union_bitmap=True)
        assert fit1[0, 0] == got[p, 0]
        # what the evaluator reports is exactly the area-weighted coverage of the set it saw (K6 against the oracle's sum)
        assert abs((1.0 - fit1[0, 0]) - (1.0 - o.coverage_of_bitmap(ub))) <= 1e-12
        seen = np.zeros_like(ub); strict = np.zeros_like(ub); lenient = np.zeros_like(ub)
        for a, b, ang in _plan_edges(plans[p], getattr(o, "has_start", False), getattr(o, "loops", False)):
            s1, s2, s3 = o.edge_bitmap(a, b, ang, classify=True)
            seen |= s1; strict |= s2; lenient |= s3
        assert popcount(ub ^ seen) <= max(1, int(FLIP_BUDGET * o.T)), (p, popcount(ub ^ seen))
        assert popcount(strict & ~ub) == 0 and popcount(ub & ~lenient) == 0, p
    return got, want


def test_population_fitness_sphere(sphere):
    g, o = sphere
    rng = np.random.default_rng(41)
    plans = random_plans(rng, o.getNumberOfBoxes(), 60, (2, 8)) + [[]]
    got, want = _compare_population(g, o, plans)
    assert got[-1].tolist() == [1.0, 0.0, 0.0, 0.0]
    assert (want[:, 0] < 1.0).sum() > 5  # some plans pass the energy gate and see the sphere


def test_population_fitness_gate_disabled(mockup256):
    g, o = mockup256
    rng = np.random.default_rng(42)
    plans = random_plans(rng, o.getNumberOfBoxes(), 12, (2, 5))
    got, want = _compare_population(g, o, plans, memo=True, no_limit=True)
    assert (want[:, 0] < 1.0).all()


def test_single_plan_api_matches_batched(sphere):
    g, o = sphere
    from inspection_path_planning_b200 import cpp_binding as cb
    plan = [[100], [101], [113], [250]]
    cov, en = g.evaluatePlan(plan, True, cb.nothing, True)
    row = g.evaluatePopulation([plan], True, True)[0]
    assert (cov, en) == (row[0], row[1])
    ocov, oen = o.evaluatePlan(plan, True, 3, True)
    assert en == pytest.approx(oen, rel=ENERGY_RTOL) and 1 - cov == pytest.approx(1 - ocov, rel=COVERAGE_RTOL)
    assert g.evaluatePlan([], True, cb.nothing) == (1.0, 0.0)
    with pytest.raises(RuntimeError):
        g.evaluatePlan(plan, True, cb.image)
    with pytest.raises(IndexError):
        g.evaluatePlan([[g.getNumberOfBoxes()]], True, cb.nothing)


def test_memoisation_semantics(oracle_mod):
    g, o = _pair(oracle_mod, "sphere", 32)
    try:
        p1, p2 = [[1], [2], [3]], [[7], [8]]
        a = g.evaluatePopulation([p1, p2], True, True)
        assert g.memo_size() == 3
        b = g.evaluatePopulation([p1, p2], True, True)  # warm: identical
        assert np.array_equal(a, b)
        c = g.evaluatePopulation([p1, p2], False, True)  # memoisation off: same values, nothing remembered
        assert np.array_equal(a, c)
        assert g.updateMemoisedEdges([p1]) == 2
        assert g.updateMemoisedEdges([[[2], [3]]]) == 1
        assert g.updateMemoisedEdges([]) == 0
        g.evaluatePopulation([p2], False, True)
        assert g.memo_size() == 0
    finally:
        g.close()


def test_loop_closure_and_start_location(oracle_mod):
    start = [6.0, 6.0, 8.0]
    g, o = _pair(oracle_mod, "sphere", 48, start=start, loops=True)
    try:
        rng = np.random.default_rng(43)
        plans = random_plans(rng, o.getNumberOfBoxes(), 20, (1, 6))
        _compare_population(g, o, plans, no_limit=True)
        _compare_population(g, o, plans, no_limit=False)
        gp, gd = g.interpretAndExportPlan(plans[0])
        op, od = o.interpretAndExportPlan(plans[0])
        assert gp == op
        np.testing.assert_allclose(gd, od, atol=1e-15)
    finally:
        g.close()


def test_union_bitmap_and_idempotence(mockup256):
    g, o = mockup256
    rng = np.random.default_rng(44)
    plans = random_plans(rng, o.getNumberOfBoxes(), 200, 6)
    fit, ub = g.evaluatePopulation(plans, True, True, union_bitmap=True)
    fit2 = g.evaluatePopulation(plans, True, True)
    assert np.array_equal(fit, fit2)
    # the union covers at least as much as the best single plan, and concatenating two plans never covers less
    cov_union = 1.0 - o.coverage_of_bitmap(ub)
    assert cov_union >= (1.0 - fit[:, 0]).max() - 1e-12
    both = g.evaluatePopulation([plans[0] + plans[1]], True, True)[0]
    assert 1 - both[0] >= max(1 - fit[0, 0], 1 - fit[1, 0]) - 1e-12
    # permuting the population permutes the rows
    perm = rng.permutation(len(plans))
    fit3 = g.evaluatePopulation([plans[i] for i in perm], True, True)
    assert np.array_equal(fit3, fit[perm])


def test_ragged_and_degenerate_plans(sphere):
    g, o = sphere
    plans = [[], [[5]], [[5], [5]], [[5], [6]], [[0], [491]], [[100], [101], [100], [101]]]
    got = g.evaluatePopulation(plans, True, True)
    want = o.evaluatePopulation(plans, True, True)
    np.testing.assert_allclose(got[:, 1:], want[:, 1:], rtol=ENERGY_RTOL)
    np.testing.assert_allclose(1 - got[:, 0], 1 - want[:, 0], rtol=COVERAGE_RTOL, atol=1e-9)
    assert got[0].tolist() == [1.0, 0.0, 0.0, 0.0]
    with pytest.raises(IndexError):
        g.evaluatePopulation([[[0], [492]]], True, True)


# ------------------------------------------------------------------ the two K4 variants (leaf bins / BVH walk per tile)
@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_default_variant_is_the_binned_kernel(fx, request):
    g, _ = request.getfixturevalue(fx)
    assert g.visibility_variant() == "bins"


@pytest.mark.parametrize("fx,n", [("sphere", 8), ("mockup256", 8), ("luis1", 8)])
def test_bvh_walk_variant_parity_and_agreement(fx, n, request):
    """The BVH-walking kernel (the path of meshes too large for the leaf tables) meets the same bar as the binned one, and the
    two differ from each other by no more than the flip budget (the order of depth ties is the only difference)."""
    g, o = request.getfixturevalue(fx)
    rng = np.random.default_rng(33)
    N = o.getNumberOfBoxes()
    stats = dict(flips=0, edges=0, seen=0, outside=0)
    try:
        for _ in range(n):
            prev, curr = (int(x) for x in rng.choice(N, 2, replace=False))
            assert g.visibility_variant(True) == "bins"
            bins = g.edge_bitmap(prev, curr)
            assert g.visibility_variant(False) == "bvh"
            nflip, outside = _check_edge(g, o, prev, curr, stats)
            bvh = g.edge_bitmap(prev, curr)
            assert nflip <= max(1, int(FLIP_BUDGET * o.T)) and outside == 0, (prev, curr, stats)
            assert popcount(bins ^ bvh) <= max(1, int(FLIP_BUDGET * o.T))
        # images through the BVH walk
        tot = bad = 0
        for eye, tgt, up in _poses(o, rng, 3):
            a, b = g.render_ids(eye, tgt, up), o.render_ids(eye, tgt, up)
            tot += a.size
            bad += int((a != b).sum())
        assert bad / tot < 2e-3, (bad, tot)
    finally:
        g.visibility_variant(True)


@pytest.mark.parametrize("fx", ["sphere", "mockup256", "luis1"])
def test_leaf_table_is_sorted_and_covers_every_visible_pixel(fx, request):
    g, o = request.getfixturevalue(fx)
    rng = np.random.default_rng(34)
    w, h = g.image_size()
    for eye, tgt, up in _poses(o, rng, 4):
        t = g.leaf_table(eye, tgt, up)
        n = t["count"]
        assert n == len(t["zq"])
        assert np.all(np.diff(t["zq"]) >= 0)  # front to back
        assert np.all(t["tx0"] <= t["tx1"]) and np.all(t["ty0"] <= t["ty1"])
        assert np.all((t["cnt"] >= 1) & (t["cnt"] <= 8)) and np.all(t["first"] + t["cnt"] <= g.T)
        img = g.render_ids(eye, tgt, up)
        ys, xs = np.nonzero(img >= 0)
        # every pixel that shows a triangle lies in a tile of the occupancy mask
        mask = t["tile_mask"]
        assert np.all((mask[ys // 32] >> (xs // 32).astype(np.uint32)) & 1)
        # ... and inside the rectangle of a leaf that holds that triangle (checked on a sample of pixels)
        if len(ys):
            pick = rng.choice(len(ys), size=min(200, len(ys)), replace=False)
            pos = np.empty(g.T, dtype=np.int64)
            pos[g.bvh_order()] = np.arange(g.T)
            for k in pick:
                p = pos[img[ys[k], xs[k]]]
                leaf = np.nonzero((t["first"] <= p) & (p < t["first"] + t["cnt"]))[0]
                assert len(leaf) == 1
                j = leaf[0]
                assert t["tx0"][j] <= xs[k] // 32 <= t["tx1"][j] and t["ty0"][j] <= ys[k] // 32 <= t["ty1"][j]


def test_population_fitness_identical_across_variants_up_to_ties(mockup256):
    g, o = mockup256
    rng = np.random.default_rng(35)
    plans = random_plans(rng, o.getNumberOfBoxes(), 10, (2, 5))
    try:
        g.visibility_variant(False)
        g.clear_memo()
        a, _ = _compare_population(g, o, plans, memo=True, no_limit=True)
        g.visibility_variant(True)
        g.clear_memo()
        b, _ = _compare_population(g, o, plans, memo=True, no_limit=True)
        assert np.array_equal(a[:, 1:], b[:, 1:])
    finally:
        g.visibility_variant(True)
        g.clear_memo()


def test_non_square_sensor_and_small_images(oracle_mod):
    """W != H (aspect = specs[2] / specs[1], the reference's argument swap), images smaller than one supertile, one camera only."""
    from inspection_path_planning_b200 import cpp_binding
    for sp in ([2.0, 60, 46, 96, 0.1, 10, 1, 1], [1.5, 40, 55, 130, 0.2, 8, 1, 0], [2.0, 46, 46, 40, 0.1, 10, 0, 1]):
        g = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
        o = oracle_mod.Oracle(mesh_path("sphere"), sp)
        try:
            assert g.image_size() == o.image_size()
            rng = np.random.default_rng(51)
            N = o.getNumberOfBoxes()
            stats = dict(flips=0, edges=0, seen=0, outside=0)
            for _ in range(6):
                prev, curr = (int(x) for x in rng.choice(N, 2, replace=False))
                for variant in (True, False):
                    g.visibility_variant(variant)
                    nflip, outside = _check_edge(g, o, prev, curr, stats)
                    assert nflip <= 1 and outside == 0, (sp, prev, curr, variant, stats)
            assert stats["seen"] > 0
        finally:
            g.close()


def test_render_batches_split_by_pose_budget_give_identical_results(monkeypatch):
    """IPP_MAX_POSES_PER_BATCH forces the new edges of one call through many small render batches (the path a 100 k-plan
    population takes when its leaf tables exceed the table budget): same fitness rows, bit for bit."""
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(64)
    rng = np.random.default_rng(61)
    a = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
    plans = random_plans(rng, a.getNumberOfBoxes(), 40, (2, 9))
    want = a.evaluatePopulation(plans, True, True)
    n_whole = a.counters()["launches"]
    a.close()
    monkeypatch.setenv("IPP_MAX_POSES_PER_BATCH", "37")
    b = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
    try:
        got = b.evaluatePopulation(plans, True, True)
        assert np.array_equal(got, want)
        assert b.counters()["launches"] > n_whole  # really went through several batches
    finally:
        b.close()


@pytest.mark.parametrize("mesh", ["mockup_only", "luis4"])
def test_supertile_task_granularity_does_not_change_results(mesh, monkeypatch):
    """A launch with few occupied supertiles hands them out as quadrants or single tiles (the small calls of an NSGA-II generation,
    csrc/visibility_bins.cu launch_visibility_bins); IPP_BINS_SPLIT forces each granularity on the same edges at 1024 x 1024: same
    edge bitmaps and fitness rows, bit for bit."""
    from inspection_path_planning_b200 import cpp_binding
    g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), specs(1024))
    try:
        rng = np.random.default_rng(77)
        plans = random_plans(rng, g.getNumberOfBoxes(), 6, (2, 5))
        edges = [(p[k][0], p[k + 1][0]) for p in plans for k in range(len(p) - 1)][:8]
        ref_fit = ref_maps = None
        for split in ("0", "2", "4"):
            monkeypatch.setenv("IPP_BINS_SPLIT", split)
            g.updateMemoisedEdges([])
            fit = g.evaluatePopulation(plans, True, True)
            maps = [g.edge_bitmap(a, b) for a, b in edges]
            if ref_fit is None:
                ref_fit, ref_maps = fit, maps
                assert any(popcount(m) > 0 for m in maps)
            else:
                assert np.array_equal(fit, ref_fit), split
                assert all(np.array_equal(m, r) for m, r in zip(maps, ref_maps)), split
    finally:
        g.close()


def test_images_above_1024_pixels_take_the_bvh_walk(oracle_mod):
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(1100)
    g = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
    o = oracle_mod.Oracle(mesh_path("sphere"), sp)
    try:
        assert g.visibility_variant() == "bvh" and g.visibility_variant(True) == "bvh"  # the tables do not apply, the request is refused
        stats = dict(flips=0, edges=0, seen=0, outside=0)
        nflip, outside = _check_edge(g, o, 100, 101, stats)
        assert nflip <= 1 and outside == 0 and stats["seen"] > 0
    finally:
        g.close()


def test_config2_full_size_properties():
    """BASELINE.json config 2 at full size (mockup_only, 1024x1024, 1 000 plans x 20 viewpoints, 18 k distinct edges, 2e5 camera
    poses) is out of the oracle's reach; size-independent properties instead: cold == warm == cold again after clearing the memo
    (determinism of the rendering), rows follow a permutation of the plans, a prefix plan never covers more than the whole plan,
    path length is additive, and the memo holds exactly the distinct directed edges of the population."""
    from inspection_path_planning_b200 import cpp_binding
    g = cpp_binding.PlanCoverageEstimator(mesh_path("mockup_only"), specs(1024))
    try:
        rng = np.random.default_rng(20261019)
        N = g.getNumberOfBoxes()
        plans = random_plans(rng, N, 1000, 20)
        cold = g.evaluatePopulation(plans, True, True)
        edges = {(p[k][0], p[k + 1][0]) for p in plans for k in range(len(p) - 1)}
        assert g.memo_size() == len(edges)
        warm = g.evaluatePopulation(plans, True, True)
        assert np.array_equal(cold, warm)
        g.clear_memo()
        again = g.evaluatePopulation(plans, True, True)
        assert np.array_equal(cold, again)
        perm = rng.permutation(len(plans))
        assert np.array_equal(g.evaluatePopulation([plans[i] for i in perm], True, True), cold[perm])
        # prefixes: coverage is monotone, path length additive
        pre = g.evaluatePopulation([p[:10] for p in plans], True, True)
        suf = g.evaluatePopulation([p[9:] for p in plans], True, True)
        assert np.all(1 - pre[:, 0] <= 1 - cold[:, 0] + 1e-12)
        np.testing.assert_allclose(pre[:, 2] + suf[:, 2], cold[:, 2], rtol=1e-12)
        assert np.all(cold[:, 0] > 0) and np.all(cold[:, 0] < 1) and (1 - cold[:, 0]).mean() > 0.3
        # the other kernel variant renders the same population to within the flip budget's area
        g.visibility_variant(False)
        g.clear_memo()
        sub = g.evaluatePopulation(plans[:60], True, True)
        assert np.array_equal(sub[:, 1:], cold[:60, 1:])
        assert np.abs(sub[:, 0] - cold[:60, 0]).max() < 5e-4 * 20
    finally:
        g.close()


def test_mesh_ingest_stl_and_obj_match_the_oracle_loader(oracle_mod, tmp_path):
    """osgDB::readNodeFile replacement (csrc/mesh_io.cpp): binary STL, ASCII STL and quad OBJ (Y-up -> Z-up, fan split) give the
    same file-order triangles as the oracle's loader -- compared through the per-triangle Heron areas and the bounding box."""
    from inspection_path_planning_b200 import cpp_binding, synth_meshes
    tris = synth_meshes.cube(side=2.0, n=3, centre=(0.5, -0.25, 1.0))
    stl_b, stl_a, obj = str(tmp_path / "cube_b.stl"), str(tmp_path / "cube_a.stl"), str(tmp_path / "grid.obj")
    synth_meshes.write_binary_stl(stl_b, tris)
    with open(stl_a, "w") as f:
        f.write("solid cube\n")
        for t in tris:
            f.write(" facet normal 0 0 0\n  outer loop\n")
            for v in t:
                f.write("   vertex %.9g %.9g %.9g\n" % tuple(v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid cube\n")
    # a 3 x 3 grid of quads bent out of plane plus one pentagon, Y-up, with vt/vn style face tokens
    gx, gy = np.meshgrid(np.arange(4.0), np.arange(4.0), indexing="ij")
    verts = np.stack([gx.ravel(), 0.3 * np.sin(gx.ravel() + gy.ravel()) + 2.0, gy.ravel()], axis=1)
    with open(obj, "w") as f:
        f.write("# test\no grid\n")
        for v in verts:
            f.write("v %.6f %.6f %.6f\n" % tuple(v))
        for i in range(3):
            for j in range(3):
                a, b, c, d = i * 4 + j, (i + 1) * 4 + j, (i + 1) * 4 + j + 1, i * 4 + j + 1
                f.write("f %d/1/1 %d/1/1 %d/1/1 %d/1/1\n" % (a + 1, b + 1, c + 1, d + 1))
        f.write("f 1 5 9 13 14\nf -1 -2 -6\n")
    sp = specs(32)
    for path, want_t in ((stl_b, len(tris)), (stl_a, len(tris)), (obj, 9 * 2 + 3 + 1)):
        # postProcessing=True: no seed plans are generated (a thin sheet has no contour to circle), only the ingest is exercised
        g = cpp_binding.PlanCoverageEstimator(path, sp, True)
        o = oracle_mod.Oracle(path, sp, post_processing=True)
        try:
            assert g.T == o.T == want_t
            assert np.array_equal(g.areas(), o.areas())
            assert np.array_equal(g.scene_info()[1], o.scene_info()[1])
            assert g.getNumberOfBoxes() == o.getNumberOfBoxes()
        finally:
            g.close()
    with pytest.raises(ValueError):
        cpp_binding.PlanCoverageEstimator(str(tmp_path / "missing.stl"), sp)
    with pytest.raises(ValueError):
        cpp_binding.PlanCoverageEstimator(__file__, sp)


@pytest.mark.parametrize("mesh,height,n_plans", [("sphere", 64, 16), ("mockup_only", 128, 6)])
def test_post_processing_heading_search_matches_the_oracle(oracle_mod, mesh, height, n_plans):
    """postProcessing = true: calculateViewingDirectionTowardsPrimitives -- both candidate headings of an edge are rendered, the one
    that shows more surface wins, a tie keeps the quick rule (OsgHelpers.cpp:451-489)."""
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(height)
    g = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), sp, True)
    o = oracle_mod.Oracle(mesh_path(mesh), sp, post_processing=True)
    q = cpp_binding.PlanCoverageEstimator(mesh_path(mesh), sp, False)
    try:
        assert g.getMaxAllowedEnergy() == o.getMaxAllowedEnergy() == -1
        assert g.evaluatePlan([], True, cpp_binding.nothing) == (1.0, 0.0)
        rng = np.random.default_rng(95)
        plans = random_plans(rng, o.getNumberOfBoxes(), n_plans, (2, 5))
        got, want = _compare_population(g, o, plans, True, False)  # strict <= seen <= lenient on every plan that is not within 1e-5
        # every edge shows at least what the quick heading shows (the quick heading is one of the two candidates)
        quick = q.evaluatePopulation(plans, True, True)
        assert np.all(got[:, 0] <= quick[:, 0] + 1e-12) and np.array_equal(got[:, 1:], quick[:, 1:])
        for p in plans[:6]:
            gp, gd = g.interpretAndExportPlan(p)
            op, od = o.interpretAndExportPlan(p)
            assert gp == op
            np.testing.assert_allclose(gd, od, atol=1e-12)
        # a plan never evaluated before: its headings are decided on demand
        fresh = [[3], [40], [77]]
        np.testing.assert_allclose(g.interpretAndExportPlan(fresh)[1], o.interpretAndExportPlan(fresh)[1], atol=1e-12)
        # what the mode does not do
        with pytest.raises(RuntimeError, match="postProcessing"):
            g.evaluatePlan([[1, 0.2], [2, 0.0]], True, cpp_binding.nothing)
        with pytest.raises(RuntimeError, match="postProcessing"):
            g.edge_bitmap(1, 2)
    finally:
        g.close()
        q.close()


def test_zero_angle_genes_evaluate_like_plain_ids(sphere):
    g, o = sphere
    from inspection_path_planning_b200 import cpp_binding as cb
    plan = [[100], [101], [113]]
    with_zero = [[100, 0], [101, 0.0], [113, 0]]
    assert g.evaluatePlan(with_zero, True, cb.nothing) == g.evaluatePlan(plan, True, cb.nothing)
    assert np.array_equal(g.evaluatePopulation([with_zero], True, False), g.evaluatePopulation([plan], True, False))


def _angle_plans(rng, n_boxes, pop, length):
    plans = random_plans(rng, n_boxes, pop, length)
    return [[[gene[0], float(rng.uniform(-1.0, 1.0)) if rng.random() < 0.7 else 0.0] for gene in p] for p in plans]


@pytest.mark.parametrize("with_start_and_loop", [False, True])
def test_angle_offset_genes_match_the_oracle(oracle_mod, with_start_and_loop):
    """[id, angleOffset] genes (SIMPLIFIED_VIEW_ANGLE = False): heading of the edge into a viewpoint turned by its offset; memo keyed
    by (previous id, id, offset) like the reference's string key."""
    from inspection_path_planning_b200 import cpp_binding as cb
    kw = dict(start=[6.0, 6.0, 8.0], loops=True) if with_start_and_loop else {}
    g, o = _pair(oracle_mod, "sphere", 64, **kw)
    try:
        rng = np.random.default_rng(91)
        plans = _angle_plans(rng, o.getNumberOfBoxes(), 24, (2, 6)) + [[]]
        g.counters(reset=True)
        got, want = _compare_population(g, o, plans, True, True)
        assert got[-1].tolist() == [1.0, 0.0, 0.0, 0.0]
        assert g.memo_size() == o.memo_size() > 0
        rendered = g.counters()["edges_rendered"]
        # the turned heading matters: the same plans with the offsets dropped score differently somewhere
        plain = g.evaluatePopulation([[[gene[0]] for gene in p] for p in plans], True, True)
        assert np.abs(plain[:, 0] - got[:, 0]).max() > 1e-3 and np.array_equal(plain[:, 1:], got[:, 1:])
        # warm: same rows, nothing rendered; single-plan call = its row
        g.counters(reset=True)
        assert np.array_equal(g.evaluatePopulation(plans, True, True), got)
        assert g.counters()["edges_rendered"] == 0
        assert g.evaluatePlan(plans[3], True, cb.nothing, True) == (got[3, 0], got[3, 1])
        # with the energy limit on, gated plans score exactly 1 and render nothing
        lim_g, lim_o = g.evaluatePopulation(plans, True, False), o.evaluatePopulation(plans, True, False)
        assert np.array_equal(lim_g[:, 0] == 1.0, lim_o[:, 0] == 1.0)
        # memo garbage collection by a sub-population, reference count
        keep = plans[:7]
        assert g.updateMemoisedEdges(keep) == o.updateMemoisedEdges(keep)
        assert g.memo_size() == o.memo_size()
        # memoization == false leaves the memo as it was
        before = g.memo_size()
        nomemo = g.evaluatePopulation(plans[7:12], False, True)
        assert np.array_equal(nomemo, got[7:12]) and g.memo_size() == before
        assert rendered > 0
    finally:
        g.close()


def test_tiny_meshes_one_triangle_and_a_twelve_triangle_box(oracle_mod, tmp_path):
    """BVH and leaf-table corner cases: a mesh of one triangle (root with a single leaf) and a box of 12 (fewer triangles than a
    leaf holds twice over), both variants, postProcessing constructor (no seed plans needed)."""
    from inspection_path_planning_b200 import cpp_binding, synth_meshes
    one = np.array([[[-3.0, -2.0, 1.0], [3.0, -2.0, 1.5], [0.0, 3.0, 2.0]]], dtype=np.float32)
    box = synth_meshes.cube(side=3.0, n=1, centre=(0.0, 0.0, 2.0))
    for name, tris in (("one", one), ("box", box)):
        path = str(tmp_path / (name + ".ippm"))
        synth_meshes.write_ippm(path, tris)
        sp = specs(96)
        g = cpp_binding.PlanCoverageEstimator(path, sp, False if name == "box" else True)
        o = oracle_mod.Oracle(path, sp, post_processing=(name != "box"))
        try:
            assert g.T == o.T == len(tris) and g.getNumberOfBoxes() == o.getNumberOfBoxes()
            rng = np.random.default_rng(81)
            c, bb = o.scene_info()
            for _ in range(6):
                eye = c + rng.normal(size=3) * 4.0
                eye = eye.astype(np.float32).astype(np.float64)
                for variant in (True, False):
                    g.visibility_variant(variant)
                    a, b = g.render_ids(eye, c, (0, 0, 1)), o.render_ids(eye, c, (0, 0, 1))
                    assert (a != b).mean() < 4e-3, (name, variant)
            if name == "box":
                stats = dict(flips=0, edges=0, seen=0, outside=0)
                N = o.getNumberOfBoxes()
                for _ in range(8):
                    prev, curr = (int(x) for x in rng.choice(N, 2, replace=False))
                    for variant in (True, False):
                        g.visibility_variant(variant)
                        nflip, outside = _check_edge(g, o, prev, curr, stats)
                        assert nflip == 0 or outside == 0
                plans = random_plans(rng, N, 12, (2, 6))
                g.visibility_variant(True)
                _compare_population(g, o, plans, memo=True, no_limit=True)
        finally:
            g.close()


def test_small_row_pool_evicts_and_splits_instead_of_refusing(oracle_mod, monkeypatch):
    """The reference's memo never refuses an entry (std::map, PlanCoverageEstimator.cpp:70,102-110).  A row pool far too small for
    a population (IPP_CACHE_GB) drops the rows of edges the call does not use and evaluates the population in pieces: same rows as an
    evaluator with room for everything, bit for bit, and as the oracle.  A bad box id raises IndexError and leaves nothing behind:
    the same evaluator keeps giving the right energies (collision flags included) afterwards."""
    from inspection_path_planning_b200 import cpp_binding
    sp = specs(32)
    big = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
    monkeypatch.setenv("IPP_CACHE_GB", "0.000004")  # ~4 kB: 31 rows of the 960-triangle sphere
    g = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), sp)
    monkeypatch.delenv("IPP_CACHE_GB")
    o = oracle_mod.Oracle(mesh_path("sphere"), sp)
    try:
        rng = np.random.default_rng(7)
        N = g.getNumberOfBoxes()
        plans = random_plans(rng, N, 30, 6)  # ~150 distinct edges
        want = big.evaluatePopulation(plans, True, True)
        got, ub = g.evaluatePopulation(plans, True, True, union_bitmap=True)
        assert g.evictions() > 0 and big.evictions() == 0
        assert np.array_equal(got, want)
        assert np.array_equal(ub, big.evaluatePopulation(plans, True, True, union_bitmap=True)[1])
        _compare_population(g, o, plans, True, True, got=got)
        assert g.memo_size() <= 31
        # again (some edges cached, most not), without memoisation, and through the device-resident entry point
        assert np.array_equal(g.evaluatePopulation(plans, True, True), want)
        assert np.array_equal(g.evaluatePopulation(plans, False, True), want)
        assert np.array_equal(g.evaluatePopulationResident(plans, True, True), want)
        # a bad id: IndexError, nothing left pending -- the next evaluations still match the oracle (energy column included)
        bad = plans[:10] + [[[0], [N]]] + plans[10:20]
        for ev in (g, big):
            with pytest.raises(IndexError):
                ev.evaluatePopulation(bad, True, False)
            ev.clear_memo() if ev is big else None
            again = ev.evaluatePopulation(plans, True, False)
            _compare_population(ev, o, plans, True, False, got=again)
        fresh = random_plans(np.random.default_rng(8), N, 20, 5)
        with pytest.raises(IndexError):
            g.evaluatePopulation([[[3], [N + 5], [4]]] + fresh, True, False)
        _compare_population(g, o, fresh, True, False)
        # updateMemoisedEdges still empties the memo
        assert g.updateMemoisedEdges([]) == 0 and g.memo_size() == 0
        # a pool that cannot hold the edges of one plan is the only refusal left
        long_plan = [random_plans(rng, N, 1, 40)[0]]
        with pytest.raises(RuntimeError, match="single plan"):
            g.evaluatePopulation(long_plan, True, True)
        assert np.array_equal(g.evaluatePopulation(plans[:3], True, True), want[:3])  # and it leaves the evaluator usable
        # [id, angleOffset] genes take the same route (their memo lives on the host)
        ang = [[[gene[0], 0.3] for gene in p] for p in plans]
        assert np.array_equal(g.evaluatePopulation(ang, True, True), big.evaluatePopulation(ang, True, True))
    finally:
        g.close()
        big.close()
