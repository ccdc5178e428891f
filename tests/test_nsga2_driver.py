"""The caller side of the hot path: NSGA-II pieces restated from DEAP, the generational loop of
plan_optimizer/optimize_coverage.py with one batched evaluator call per generation, and the result-store dictionaries."""
import os
import random

import numpy as np
import pytest

from conftest import mesh_path
from inspection_path_planning_b200 import optimize_coverage as oc


def _ind(fit, genes=((1,), (2,))):
    i = oc.Individual([list(g) for g in genes])
    i.fitness = tuple(fit)
    return i


def test_dominance_minimises_both_objectives():
    assert oc.dominates((0.2, 5.0), (0.3, 5.0))
    assert oc.dominates((0.2, 4.0), (0.3, 5.0))
    assert not oc.dominates((0.2, 6.0), (0.3, 5.0))
    assert not oc.dominates((0.2, 5.0), (0.2, 5.0))


def test_fronts_and_crowding_distance_known_answer():
    pts = [(0.1, 9.0), (0.5, 5.0), (0.9, 1.0), (0.6, 6.0), (0.95, 9.5), (0.5, 5.0)]
    pop = [_ind(p) for p in pts]
    fronts = oc.sort_nondominated(pop)
    assert [sorted(i.fitness for i in f) for f in fronts] == [[(0.1, 9.0), (0.5, 5.0), (0.5, 5.0), (0.9, 1.0)], [(0.6, 6.0)], [(0.95, 9.5)]]
    front = [_ind(p) for p in [(0.0, 4.0), (1.0, 3.0), (2.0, 1.0), (4.0, 0.0)]]
    oc.assign_crowding_dist(front)
    d = [i.crowding_dist for i in front]
    assert d[0] == float("inf") and d[3] == float("inf")
    assert d[1] == pytest.approx((2.0 - 0.0) / 4.0 + (4.0 - 1.0) / 4.0)
    assert d[2] == pytest.approx((4.0 - 1.0) / 4.0 + (3.0 - 0.0) / 4.0)


def test_sel_nsga2_takes_whole_fronts_then_the_least_crowded():
    pts = [(0.0, 4.0), (1.0, 3.0), (1.1, 2.9), (2.0, 1.0), (4.0, 0.0), (5.0, 5.0)]
    pop = [_ind(p) for p in pts]
    chosen = oc.sel_nsga2(pop, 4)
    got = sorted(i.fitness for i in chosen)
    assert (5.0, 5.0) not in got and (0.0, 4.0) in got and (4.0, 0.0) in got and len(got) == 4
    # of the two crowded neighbours at most one survives
    assert sum(f in got for f in [(1.0, 3.0), (1.1, 2.9)]) == 1


def test_tournament_and_crossover():
    rng = random.Random(3)
    pop = [_ind((i / 10.0, 10.0 - i)) for i in range(8)]
    oc.sel_nsga2(pop, len(pop))
    out = oc.sel_tournament_dcd(pop, len(pop), rng)
    assert len(out) == 8 and all(o in pop for o in out)
    with pytest.raises(ValueError):
        oc.sel_tournament_dcd(pop[:6], 6, rng)
    a, b = oc.Individual([[1], [2], [3]]), oc.Individual([[7], [8], [9], [10], [11]])
    before = sorted(g[0] for g in a + b)
    oc.cx_messy_one_point(a, b, rng)
    assert sorted(g[0] for g in a + b) == before and len(a) + len(b) == 8


def test_vectorised_front_sort_keeps_the_order_of_the_pairwise_algorithm():
    """sort_nondominated builds the dominance relation with numpy; fronts and the order inside them (which decides who survives
    a crowding-distance tie and how selTournamentDCD pairs the survivors) must be those of the pairwise loop of tools.sortNondominated."""
    import random

    def pairwise(individuals):
        by_fit = {}
        for ind in individuals:
            by_fit.setdefault(tuple(ind.fitness), []).append(ind)
        fits = list(by_fit)
        dominated_by = {f: [] for f in fits}
        n_dom = {f: 0 for f in fits}
        current = []
        for i, fi in enumerate(fits):
            for fj in fits[i + 1:]:
                if oc.dominates(fi, fj):
                    n_dom[fj] += 1
                    dominated_by[fi].append(fj)
                elif oc.dominates(fj, fi):
                    n_dom[fi] += 1
                    dominated_by[fj].append(fi)
            if n_dom[fi] == 0:
                current.append(fi)
        fronts = []
        while current:
            fronts.append([ind for f in current for ind in by_fit[f]])
            nxt = []
            for f in current:
                for g in dominated_by[f]:
                    n_dom[g] -= 1
                    if n_dom[g] == 0:
                        nxt.append(g)
            current = nxt
        return fronts

    rng = random.Random(3)
    for _ in range(60):
        n = rng.randint(1, 90)
        inds = []
        for i in range(n):
            ind = oc.Individual([[i]])
            ind.fitness = (rng.randint(0, 10) / 10.0, float(rng.randint(0, 12)))
            inds.append(ind)
        got = [[x[0][0] for x in f] for f in oc.sort_nondominated(inds)]
        want = [[x[0][0] for x in f] for f in pairwise(inds)]
        assert got == want
    assert oc.sort_nondominated([]) == []


def test_pareto_front_keeps_only_non_dominated_and_no_duplicates():
    pf = oc.ParetoFront()
    pf.update([_ind((0.5, 5.0)), _ind((0.6, 6.0)), _ind((0.4, 7.0))])
    assert sorted(i.fitness for i in pf) == [(0.4, 7.0), (0.5, 5.0)]
    pf.update([_ind((0.5, 5.0)), _ind((0.3, 4.0))])
    assert sorted(i.fitness for i in pf) == [(0.3, 4.0)]
    assert oc.hypervolume_2d([(0.3, 4.0)], [1.0, 10.0]) == pytest.approx(0.7 * 6.0)
    assert oc.hypervolume_2d([(0.2, 8.0), (0.6, 2.0)], [1.0, 10.0]) == pytest.approx(0.8 * 2.0 + 0.4 * 6.0)


def test_cleanup_and_mutation_respect_the_reference_rules():
    ind = oc.Individual([[3], [3], [4], [4], [4], [3]])
    oc.clean_up_individual(ind)
    assert [g[0] for g in ind] == [3, 4, 3]
    p = oc.Parameters()
    rng = random.Random(0)
    short = oc.Individual([[1], [2]])
    for _ in range(50):
        oc.mutate_waypoints(short, 10, p, rng)
        assert len(short) >= p.MIN_PLAN_SIZE
    with pytest.raises(AttributeError):
        oc.Parameters(NOT_A_PARAMETER=1)


@pytest.fixture(scope="module")
def sphere_run(oracle_mod, tmp_path_factory):
    sp = list(oc.Parameters.SENSOR_PARAMETERS)
    sp[3] = 64
    params = oc.Parameters(NUM_INDIVS=16, NUM_GENERATIONS=4, SENSOR_PARAMETERS=sp)
    ev = oracle_mod.Oracle(mesh_path("sphere"), sp)
    pop, pf, logbook, stats = oc.run_nsga2(ev, params, seed=11)
    return ev, params, pop, pf, logbook, stats, tmp_path_factory.mktemp("run")


def test_generational_loop_with_the_cpu_evaluator(sphere_run, oracle_mod):
    ev, params, pop, pf, logbook, stats, _ = sphere_run
    assert len(pop) == params.NUM_INDIVS and len(logbook) == params.NUM_GENERATIONS + 1
    assert all(len(i) >= params.MIN_PLAN_SIZE for i in pop)
    # fitness values are the evaluator's own answers for the stored genotypes
    for ind in pop[:6]:
        assert tuple(ev.evaluatePlan(list(ind), True, 3, False)) == pytest.approx(ind.fitness)
    # the front is mutually non-dominated and the memo was trimmed to the population once per generation
    fits = [i.fitness for i in pf]
    assert not any(oc.dominates(a, b) for a in fits for b in fits)
    assert stats["calls"] <= 3 + params.NUM_GENERATIONS + 10 and stats["memoised_edges"] == ev.memo_size()
    # same seed, same run
    ev2 = oracle_mod.Oracle(mesh_path("sphere"), params.SENSOR_PARAMETERS)
    pop2, pf2, _, _ = oc.run_nsga2(ev2, params, seed=11)
    assert [list(i) for i in pop2] == [list(i) for i in pop] and [i.fitness for i in pop2] == [i.fitness for i in pop]


def test_result_stores_use_the_reference_keys(sphere_run):
    ev, params, pop, pf, _, stats, out = sphere_run
    for name in ("winner_indivs.yml", "winner_indivs.pkl"):
        path = oc.store_population_and_parameters(pop, name, str(out), "sphere.obj", params, ev, pf, 1.5, stats["memoised_edges"])
        d = oc.load_dumped_population(path)
        assert set(d) == {"population", "scene", "plan_origin", "sensor_params", "pareto_front", "max_energy", "num_memoized_solutions",
                          "elapsed_time", "plan_loops_around"}
        assert len(d["population"]) == len(pop) and d["population"][0]["genotype1"] == [list(g) for g in pop[0]]
        assert d["population"][0]["fitness_weights"] == [-1, -1] and d["max_energy"] == ev.getMaxAllowedEnergy()
        assert len(d["pareto_front"]) == len(pf)


@pytest.mark.gpu
def test_generational_loop_on_the_gpu_matches_the_cpu_evaluator(sphere_run, oracle_mod):
    """The loop of optimize_coverage.py:109-213 with libipp.so as the evaluator.  Generation 0 is evaluated on the same genotypes as
    the CPU run; after that an edge-epsilon flip may send selection another way, so the later generations are checked through the
    invariant that matters: every fitness the loop stored is the oracle's answer for that genotype (strict / lenient classification
    of test_gpu_parity._compare_population), for the whole final population and the whole Pareto front."""
    from inspection_path_planning_b200 import cpp_binding
    from test_gpu_parity import _compare_population
    ev, params, pop, pf, logbook, stats, _ = sphere_run
    g = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), params.SENSOR_PARAMETERS)
    o = oracle_mod.Oracle(mesh_path("sphere"), params.SENSOR_PARAMETERS)
    try:
        gpop, gpf, glog, gstats = oc.run_nsga2(g, params, seed=11)
        assert glog[0]["fit_min"][1] == pytest.approx(logbook[0]["fit_min"][1], rel=1e-12)
        assert glog[0]["fit_avg"][1] == pytest.approx(logbook[0]["fit_avg"][1], rel=1e-12)
        assert glog[0]["fit_avg"][0] == pytest.approx(logbook[0]["fit_avg"][0], abs=2e-4)
        assert len(gpop) == params.NUM_INDIVS and len(glog) == params.NUM_GENERATIONS + 1
        assert gstats["memoised_edges"] == g.memo_size()
        for group in (gpop, list(gpf)):
            plans = [[list(gene) for gene in ind] for ind in group]
            stored = np.array([[ind.fitness[0], ind.fitness[1]] for ind in group])
            got = g.evaluatePopulation(plans, True, False)
            assert np.array_equal(got[:, :2], stored)  # what the loop stored is what the evaluator answers (memo or not)
            _compare_population(g, o, plans, True, False, got=got)
        fits = [i.fitness for i in gpf]
        assert not any(oc.dominates(a, b) for a in fits for b in fits)
    finally:
        g.close()


@pytest.mark.gpu
def test_result_stores_and_plan_export_through_libipp(oracle_mod, tmp_path):
    """SURVEY 8(f) N2 / N3 with libipp.so on both sides of the files: a short run is stored (winner_indivs.yml / .pkl, the
    reference's keys), read back, and every stored plan exported to Rock waypoint YAML through a postProcessing = True evaluator
    (planEvaluatorSetup.py:13-24) -- positions and yaws equal to what the oracle's post-processing evaluator exports."""
    import math
    import yaml
    from inspection_path_planning_b200 import cpp_binding
    from inspection_path_planning_b200 import plan_exporter as pe
    params = oc.Parameters(NUM_INDIVS=12, NUM_GENERATIONS=3, SENSOR_PARAMETERS=[2.0, 46.0, 46.0, 64, 0.1, 10, 1, 1])
    g = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), params.SENSOR_PARAMETERS)
    try:
        pop, pf, logbook, stats = oc.run_nsga2(g, params, seed=5)
        for name in ("winner_indivs.yml", "winner_indivs.pkl"):
            path = oc.store_population_and_parameters(pop, name, str(tmp_path), mesh_path("sphere"), params, g, pf, 0.5, stats["memoised_edges"])
            d = oc.load_dumped_population(path)
            assert d["max_energy"] == g.getMaxAllowedEnergy() and d["num_memoized_solutions"] == g.memo_size()
            assert [p["genotype1"] for p in d["population"]] == [[list(x) for x in ind] for ind in pop]
            assert [tuple(p["fitness"]) for p in d["population"]] == [ind.fitness for ind in pop]
    finally:
        g.close()
    oc.store_run_info(str(tmp_path), params)
    gp = cpp_binding.PlanCoverageEstimator(mesh_path("sphere"), params.SENSOR_PARAMETERS, True)
    op = oracle_mod.Oracle(mesh_path("sphere"), params.SENSOR_PARAMETERS, post_processing=True)
    try:
        paths = pe.export_all_winner_individuals(str(tmp_path), position_offsets=(1.0, 2.0, 3.0), evaluator=gp)
        assert len(paths) == len(pop)
        for path, ind in zip(paths, pop):
            wps = yaml.safe_load(open(path))["waypoints"]
            want = pe.plan_to_waypoints(op, [list(x) for x in ind], position_offsets=(1.0, 2.0, 3.0))
            assert len(wps) == len(want) == len(ind)
            for a, b in zip(wps, want):
                assert a["position"] == pytest.approx(b["position"])
                dy = (a["orientation"]["y"] - b["orientation"]["y"] + math.pi) % (2 * math.pi) - math.pi
                assert abs(dy) < 1e-9
    finally:
        gp.close()


def test_plan_export_yaml_follows_the_rock_format(sphere_run):
    import math
    import yaml
    from inspection_path_planning_b200 import plan_exporter as pe
    ev, params, pop, pf, _, stats, out = sphere_run
    plan = [list(g) for g in max(pop, key=len)]
    pos, heads = ev.interpretAndExportPlan(plan)
    path = pe.export_plan_to_yaml(ev, plan, str(out / "plan_0.yml"), position_offsets=(35.0, -10.0, -15.0))
    d = yaml.safe_load(open(path))
    wps = d["waypoints"]
    assert len(wps) == len(pos) == len(heads) + 1
    assert wps[0]["orientation"] == {"r": 0.0, "p": 0.0, "y": 0.0}
    for k, w in enumerate(wps):
        assert w["position"]["x"] == pytest.approx(pos[k][0] + 35.0) and w["position"]["z"] == pytest.approx(pos[k][2] - 15.0)
        if k:
            assert w["orientation"]["y"] == pytest.approx(math.atan2(heads[k - 1][1], heads[k - 1][0]))
    # a quarter turn about +z turns every yaw by 90 degrees and maps (x, y) to (-y, x)
    rot = pe.plan_to_waypoints(ev, plan, rotation_degrees_around_z=90.0)
    base = pe.plan_to_waypoints(ev, plan)
    assert rot[1]["position"]["x"] == pytest.approx(-base[1]["position"]["y"]) and rot[1]["position"]["y"] == pytest.approx(base[1]["position"]["x"])
    dy = (rot[1]["orientation"]["y"] - base[1]["orientation"]["y"] - math.pi / 2 + math.pi) % (2 * math.pi) - math.pi
    assert abs(dy) < 1e-9
    assert pe.plan_to_waypoints(ev, []) == []


def test_circling_plan_generator_stores_the_seed_plans_with_their_fitness(oracle_mod, tmp_path):
    from inspection_path_planning_b200 import circling_plan_generator as cg
    sp = list(oc.Parameters.SENSOR_PARAMETERS)
    sp[3] = 64
    params = oc.Parameters(SENSOR_PARAMETERS=sp)
    ev = oracle_mod.Oracle(mesh_path("sphere"), sp)
    inds = cg.generate(ev, params, "sphere.obj", str(tmp_path))
    plans = ev.getSimplePlans()
    assert len(inds) == len(plans) > 0
    # the first complete plan defines the energy limit (PlanEnergyEvaluator.cpp:40-46)
    assert inds[0].fitness[1] == pytest.approx(ev.getMaxAllowedEnergy(), rel=1e-12)
    for ind, p in zip(inds, plans):
        assert [g[0] for g in ind] == [int(g[0]) for g in p]
        assert tuple(ev.evaluatePlan(list(ind), False, 3, False)) == pytest.approx(ind.fitness)
    d = oc.load_dumped_population(str(tmp_path / "populations" / "circling_plans.yml"))
    assert len(d["population"]) == len(inds) and d["pareto_front"] is None and d["scene"] == "sphere.obj"


def test_polynomial_mutation_stays_in_bounds_and_moves_values():
    rng = random.Random(5)
    vals = [0.0, 0.9, -0.95, 0.3] * 50
    out = oc.mut_polynomial_bounded(list(vals), 20.0, -1.0, 1.0, 1.0, rng)
    assert all(-1.0 <= v <= 1.0 for v in out) and sum(a != b for a, b in zip(out, vals)) > 150
    assert max(abs(a - b) for a, b in zip(out, vals)) < 1.0  # eta = 20: small moves
    assert oc.mut_polynomial_bounded(list(vals), 20.0, -1.0, 1.0, 0.0, rng) == vals


def test_generational_loop_with_view_angle_genes(oracle_mod):
    sp = list(oc.Parameters.SENSOR_PARAMETERS)
    sp[3] = 48
    params = oc.Parameters(NUM_INDIVS=12, NUM_GENERATIONS=3, SENSOR_PARAMETERS=sp, SIMPLIFIED_VIEW_ANGLE=False, ANGLE_MUTATION_PROBABILITY=0.3)
    ev = oracle_mod.Oracle(mesh_path("sphere"), sp)
    pop, pf, logbook, stats = oc.run_nsga2(ev, params, seed=3)
    assert all(len(g) == 2 and -1.0 <= g[1] <= 1.0 for ind in pop for g in ind)
    assert any(g[1] != 0.0 for ind in pop for g in ind)  # offsets were mutated and kept
    for ind in pop[:4]:
        assert tuple(ev.evaluatePlan(list(ind), True, 3, False)) == pytest.approx(ind.fitness)


def test_parameter_file_round_trip(tmp_path):
    p = oc.Parameters(NUM_INDIVS=12, PLAN_ORIGIN=[1.0, 2.0, 3.0], SIMPLIFIED_VIEW_ANGLE=False)
    path = p.write_python(str(tmp_path / "Parameters.py"))
    q = oc.Parameters.from_python(path)
    assert q.as_dict() == p.as_dict() and q.NUM_INDIVS == 12 and q.PLAN_ORIGIN == [1.0, 2.0, 3.0]
    (tmp_path / "other.py").write_text("NUM_GENERATIONS = 3\nNOT_A_PARAMETER = 1\nlowercase = 2\n")
    r = oc.Parameters.from_python(str(tmp_path / "other.py"))
    assert r.NUM_GENERATIONS == 3 and not hasattr(r, "NOT_A_PARAMETER") and r.NUM_INDIVS == oc.Parameters.NUM_INDIVS
    # the shape of the reference's own settings/Parameters.py (and of the copy its runs leave in experiment_parameters/): an import
    # at the top, comments, a parameter that names an earlier one.  Nothing is executed.
    (tmp_path / "ref_style.py").write_text(
        "import settings.Constants_and_Datastructures as const\n\n__author__ = 'someone'\n\n"
        "NUM_INDIVS = 48  # divisible by 4\nNUM_GENERATIONS = 2 * 50\nPLAN_ORIGIN = None\nPLAN_LOOPS_AROUND = not False\n"
        "DOWN_CAM_ACTIVE = 0\nSENSOR_PARAMETERS = [2.0, 46.0, 46.0, 512, 0.1, 10, 1, DOWN_CAM_ACTIVE]\n"
        "HYPERVOLUME_REF_POINT = [1.01, const.SOMETHING]\nopen('/nonexistent/side_effect', 'w')\n")
    q = oc.Parameters.from_python(str(tmp_path / "ref_style.py"))
    assert q.NUM_INDIVS == 48 and q.NUM_GENERATIONS == 100 and q.PLAN_LOOPS_AROUND is True
    assert q.SENSOR_PARAMETERS == [2.0, 46.0, 46.0, 512, 0.1, 10, 1, 0]
    assert q.HYPERVOLUME_REF_POINT == oc.Parameters.HYPERVOLUME_REF_POINT  # needs the skipped import: the default stays


def test_command_line_run_and_multirun_store_what_the_reference_stores(oracle_mod, tmp_path):
    """optimize_coverage.py:216-270 and multiRun.py:14-27 with the CPU evaluator standing in: base_plans, genN, winner_indivs as
    .pkl and .yml, plan_ymls/ for a normal run; parameters read from experiment_parameters/ and no plan_ymls/ in a multirun."""
    from inspection_path_planning_b200 import multi_run

    def factory(scene, params, post_processing):
        return oracle_mod.Oracle(scene, params.SENSOR_PARAMETERS, post_processing, params.PLAN_ORIGIN, params.PLAN_LOOPS_AROUND)

    out = tmp_path / "single"
    assert oc.main(["-i", mesh_path("sphere"), "-o", str(out), "--generations", "2", "--pop", "8", "--image_height", "32", "--seed", "5"],
                   factory) == 0
    stored = sorted(os.listdir(out / "populations"))
    assert stored == ["base_plans.pkl", "base_plans.yml", "gen0.pkl", "gen0.yml", "winner_indivs.pkl", "winner_indivs.yml"]
    win = oc.load_dumped_population(str(out / "populations" / "winner_indivs.yml"))
    assert len(win["population"]) == 8 and win["sensor_params"][3] == 32 and win["elapsed_time"] > 0
    base = oc.load_dumped_population(str(out / "populations" / "base_plans.pkl"))
    assert len(base["population"]) > 0 and all(p["fitness"] is not None for p in base["population"]) and base["pareto_front"] is None
    assert sorted(os.listdir(out / "plan_ymls")) == sorted("plan_%d.yml" % i for i in range(8))

    exp = tmp_path / "experiment"
    assert multi_run.main(["-i", mesh_path("sphere"), "-o", str(exp), "--runs", "2", "--generations", "1", "--pop", "8",
                           "--image_height", "32", "--seed", "3"], factory) == 0
    stored_params = oc.Parameters.from_python(str(exp / "experiment_parameters" / "Parameters.py"))
    assert stored_params.NUM_INDIVS == 8 and stored_params.NUM_GENERATIONS == 1 and stored_params.SENSOR_PARAMETERS[3] == 32
    for i in range(2):
        run = exp / ("run%d" % i)
        assert "winner_indivs.pkl" in os.listdir(run / "populations") and not (run / "plan_ymls").exists()
        d = oc.load_dumped_population(str(run / "populations" / "winner_indivs.pkl"))
        assert len(d["population"]) == 8 and d["sensor_params"][3] == 32
    a = oc.load_dumped_population(str(exp / "run0" / "populations" / "winner_indivs.pkl"))["population"]
    b = oc.load_dumped_population(str(exp / "run1" / "populations" / "winner_indivs.pkl"))["population"]
    assert [p["genotype1"] for p in a] != [p["genotype1"] for p in b]  # run i uses seed + i


def test_sharp_turns_are_split_into_two_waypoints():
    """plan_exporter.split_sharply_turning_waypoints (:34-81): a U-turn corner becomes two waypoints split_distance before and
    after it, the edge between them heading along the average of the two headings; gentle corners and the end points stay."""
    import math
    from inspection_path_planning_b200 import plan_exporter as pe

    class Fake:
        def interpretAndExportPlan(self, plan, a, b):
            pos = [(0.0, 0.0, 5.0), (20.0, 0.0, 5.0), (20.0, 20.0, 5.0), (21.0, 40.0, 5.0), (21.0, 43.0, 5.0), (24.0, 43.0, 5.0)]
            head = [(0.0, 1.0, 0.0), (-1.0, 0.0, 0.0), (-1.0, 0.0, 0.0), (-1.0, 0.0, 0.0), (0.0, -1.0, 0.0)]
            return pos, head

    assert pe.angle_between(np.array([1.0, 0, 0]), np.array([0, 1.0, 0])) == pytest.approx(math.pi / 2)
    assert pe.angle_between(np.zeros(3), np.array([0, 1.0, 0])) == 0.0
    plain = pe.plan_to_waypoints(Fake(), [[0]] * 6)
    split = pe.plan_to_waypoints(Fake(), [[0]] * 6, min_turn_angle=math.pi / 4)
    assert len(plain) == 6 and len(split) == 8  # the right angles at (20, 0) and (21, 43) are split, the 3 degree bend is not
    xy = [(w["position"]["x"], w["position"]["y"]) for w in split]
    assert xy[0] == (0.0, 0.0) and xy[-1] == (24.0, 43.0)
    assert xy[1] == pytest.approx((16.0, 0.0)) and xy[2] == pytest.approx((20.0, 4.0))  # 4 m before and after the corner
    assert xy[3] == pytest.approx((20.0, 20.0))
    assert xy[5] == pytest.approx((21.0, 41.5)) and xy[6] == pytest.approx((22.5, 43.0))  # short edges: half the edge at most
    yaw = [w["orientation"]["y"] for w in split]
    assert yaw[1] == pytest.approx(math.pi / 2)                       # on the way to the first new waypoint: the old heading
    assert yaw[2] == pytest.approx(math.atan2(1.0, -1.0))             # between the two: the average of (0, 1) and (-1, 0)
    assert yaw[3] == pytest.approx(math.pi) and yaw[7] == pytest.approx(-math.pi / 2)
    assert all(w["position"]["z"] == 5.0 for w in split)
