#!/usr/bin/env python
"""NSGA-II coverage path planner on top of the batched B200 evaluator -- the caller side of the hot path.

Follows plan_optimizer/optimize_coverage.py:62-213 step by step (seeding at USE_PLAN_SEED_CHANCE, rejection-sampled random
individuals, selTournamentDCD -> cxMessyOnePoint -> insert / delete mutation -> re-evaluation of the changed individuals ->
selNSGA2 over parents + valid offspring, ParetoFront, updateMemoisedEdges once per generation) with one change, the one
BASELINE.json's north star allows: the `toolbox.map(evaluate, invalid_ind)` of the reference (:141, :191) is a single
`evaluator.evaluatePopulation(invalid_ind, memoization, False)` call per generation.

DEAP is not a dependency: the four DEAP pieces the reference uses (tools.selNSGA2, tools.selTournamentDCD, tools.cxMessyOnePoint,
tools.ParetoFront, with base.Fitness(weights=(-1, -1)) dominance) are restated here from their published definitions
(Deb et al. 2002; DEAP 1.x documentation).

run_nsga2 takes any object with the cpp_binding.PlanCoverageEstimator surface + evaluatePopulation; the command line always
builds the CUDA evaluator (inspection_path_planning_b200.cpp_binding) -- there is no CPU path in the product.

    python -m inspection_path_planning_b200.optimize_coverage -i tests/golden/meshes/sphere.ippm -o /tmp/run --generations 20 --pop 100
"""
import argparse
import copy
import os
import random
import time

import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# settings/Parameters.py:1-58 (the shipped defaults)
# ---------------------------------------------------------------------------------------------------------------------
class Parameters:
    NUM_INDIVS = 40  # has to be divisible by 4 (selTournamentDCD)
    NUM_GENERATIONS = 400
    STORAGE_INTERVAL = 10
    WAYPOINT_MUTATION_PROBABILITY = 0.1
    CROSSOVER_PROB = 0.1
    USE_PLAN_SEED_CHANCE = 0.35
    MIN_PLAN_SIZE = 2
    MAX_PLAN_SIZE = 15
    PLAN_ORIGIN = None
    PLAN_LOOPS_AROUND = False
    HYPERVOLUME_REF_POINT = [1.01, 215]
    SIMPLIFIED_VIEW_ANGLE = True  # False: genes are [id, angleOffset] and the offsets are mutated too (mutateViewingAngles)
    ANGLE_MUTATION_PROBABILITY = 0.01  # per waypoint
    POLY_MUTATION_ETA = 20.0  # the reference uses an undefined POLY_MUTATION_ETA_PARAMETER (mutationAndRecombination.py:52); DEAP's usual value
    SENSOR_PARAMETERS = [2.0, 46.0, 46.0, 1024, 0.1, 10, 1, 1]
    USING_EDGE_MEMOISATION = True

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(Parameters, k):
                raise AttributeError("unknown parameter " + k)
            setattr(self, k, v)

    @classmethod
    def names(cls):
        return [k for k in vars(cls) if k.isupper()]

    def as_dict(self):
        return {k: copy.deepcopy(getattr(self, k)) for k in self.names()}

    def write_python(self, path):
        """A settings/Parameters.py for this run: one `NAME = value` line per parameter (multiRun keeps it beside the runs)."""
        with open(path, "w") as f:
            f.write("# experiment parameters (inspection_path_planning_b200.optimize_coverage.Parameters)\n")
            for k, v in self.as_dict().items():
                f.write("%s = %r\n" % (k, v))
        return path

    @classmethod
    def from_python(cls, path):
        """Reads a parameter file as multirun does (optimize_coverage.py:223-227): every top-level `NAME = value` whose value is a
        literal, a name assigned earlier in the file (DOWN_CAM_ACTIVE = True ... [.., DOWN_CAM_ACTIVE]) or simple arithmetic on
        those.  Nothing is executed: the reference's own settings/Parameters.py -- and the copy store_run_info leaves in
        experiment_parameters/ -- starts with `import settings.Constants_and_Datastructures as const`; imports and any other
        statement are skipped.  Names this class does not know are ignored."""
        import ast
        import operator
        scope = {"True": True, "False": False, "None": None}
        ops = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.FloorDiv: operator.floordiv,
               ast.Mod: operator.mod, ast.Pow: operator.pow}

        def ev(node):
            if isinstance(node, ast.Constant):
                return node.value
            if isinstance(node, ast.Name):
                return scope[node.id]
            if isinstance(node, (ast.List, ast.Tuple)):
                vals = [ev(e) for e in node.elts]
                return vals if isinstance(node, ast.List) else tuple(vals)
            if isinstance(node, ast.Dict):
                return {ev(k): ev(v) for k, v in zip(node.keys, node.values)}
            if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd, ast.Not)):
                v = ev(node.operand)
                return -v if isinstance(node.op, ast.USub) else (+v if isinstance(node.op, ast.UAdd) else (not v))
            if isinstance(node, ast.BinOp) and type(node.op) in ops:
                return ops[type(node.op)](ev(node.left), ev(node.right))
            raise ValueError("not a literal")

        with open(path) as f:
            tree = ast.parse(f.read(), path)
        for stmt in tree.body:
            if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name):
                try:
                    scope[stmt.targets[0].id] = ev(stmt.value)
                except (ValueError, KeyError, TypeError, ZeroDivisionError):
                    pass  # e.g. a value that needs the skipped import
        return cls(**{k: v for k, v in scope.items() if k.isupper() and hasattr(cls, k)})


FITNESS_WEIGHTS = (-1, -1)  # settings/Constants_and_Datastructures.py:59-60: both objectives minimised


class Individual(list):
    """evolutionary_operators/individuals.py:8-31: a list of genes [id] with a fitness (coverageScore, energy)."""

    def __init__(self, genes):
        list.__init__(self, genes)
        self.fitness = None  # (coverageScore, energy) or None = invalid
        self.crowding_dist = 0.0

    @classmethod
    def rand_init(cls, length, num_boxes, rng, with_angles=False):
        """individuals.py:17-26: random viewpoints; with view angles every gene starts looking straight ahead ([id, 0])."""
        return cls([[rng.randrange(0, num_boxes)] + ([0.0] if with_angles else []) for _ in range(length)])

    def clone(self):
        c = Individual([list(g) for g in self])
        c.fitness = self.fitness
        c.crowding_dist = self.crowding_dist
        return c


def dominates(a, b):
    """base.Fitness.dominates with weights (-1, -1): a is no worse in every objective and better in one."""
    not_equal = False
    for x, y in zip(a, b):
        if x > y:
            return False
        if x < y:
            not_equal = True
    return not_equal


# ---------------------------------------------------------------------------------------------------------------------
# DEAP pieces
# ---------------------------------------------------------------------------------------------------------------------
def sort_nondominated(individuals):
    """Fast non-dominated sort (Deb 2002); individuals with equal fitness share a front, as in tools.sortNondominated.  The
    dominance relation of the distinct fitness values is one numpy comparison; fronts and the order inside them are the ones the
    pairwise loop of the textbook algorithm produces (first front in order of appearance, then every value when its last
    dominator of the previous front has been taken, dominated values in order of appearance)."""
    by_fit = {}
    for ind in individuals:
        by_fit.setdefault(tuple(ind.fitness), []).append(ind)
    fits = list(by_fit)
    if not fits:
        return []
    f = np.asarray(fits, dtype=np.float64)
    le = (f[:, None, :] <= f[None, :, :]).all(axis=2)
    lt = (f[:, None, :] < f[None, :, :]).any(axis=2)
    dom = le & lt  # dom[i, j]: value i dominates value j
    n_dom = dom.sum(axis=0)
    current = np.nonzero(n_dom == 0)[0]
    done = np.zeros(len(fits), dtype=bool)
    fronts = []
    while current.size:
        fronts.append([ind for i in current for ind in by_fit[fits[i]]])
        done[current] = True
        sub = dom[current]
        n_dom = n_dom - sub.sum(axis=0)
        nxt = np.nonzero((n_dom == 0) & ~done)[0]
        if nxt.size:
            last = sub.shape[0] - 1 - np.argmax(sub[::-1, nxt], axis=0)  # position in `current` of the last dominator
            nxt = nxt[np.lexsort((nxt, last))]
        current = nxt
    return fronts


def assign_crowding_dist(front):
    """tools.emo.assignCrowdingDist"""
    if not front:
        return
    dist = [0.0] * len(front)
    crowd = [(ind.fitness, i) for i, ind in enumerate(front)]
    for k in range(len(front[0].fitness)):
        crowd.sort(key=lambda e: e[0][k])
        dist[crowd[0][1]] = float("inf")
        dist[crowd[-1][1]] = float("inf")
        if crowd[-1][0][k] == crowd[0][0][k]:
            continue
        norm = float(crowd[-1][0][k] - crowd[0][0][k])
        for prev, cur, nxt in zip(crowd[:-2], crowd[1:-1], crowd[2:]):
            dist[cur[1]] += (nxt[0][k] - prev[0][k]) / norm
    for i, d in enumerate(dist):
        front[i].crowding_dist = d


def sel_nsga2(individuals, k):
    """tools.selNSGA2: whole fronts while they fit, the last one by descending crowding distance."""
    fronts = sort_nondominated(individuals)
    for f in fronts:
        assign_crowding_dist(f)
    chosen = []
    for f in fronts:
        if len(chosen) + len(f) <= k:
            chosen.extend(f)
        else:
            rest = sorted(f, key=lambda ind: ind.crowding_dist, reverse=True)
            chosen.extend(rest[:k - len(chosen)])
            break
    return chosen


def sel_tournament_dcd(individuals, k, rng):
    """tools.selTournamentDCD: binary tournaments on dominance, then crowding distance; len(individuals) divisible by 4."""
    if len(individuals) % 4 != 0:
        raise ValueError("selTournamentDCD: individuals length must be a multiple of 4")

    def tourn(a, b):
        if dominates(a.fitness, b.fitness):
            return a
        if dominates(b.fitness, a.fitness):
            return b
        if a.crowding_dist < b.crowding_dist:
            return b
        if a.crowding_dist > b.crowding_dist:
            return a
        return a if rng.random() <= 0.5 else b

    s1 = rng.sample(individuals, len(individuals))
    s2 = rng.sample(individuals, len(individuals))
    chosen = []
    for i in range(0, k, 4):
        chosen.append(tourn(s1[i], s1[i + 1]))
        chosen.append(tourn(s1[i + 2], s1[i + 3]))
        chosen.append(tourn(s2[i], s2[i + 1]))
        chosen.append(tourn(s2[i + 2], s2[i + 3]))
    return chosen


def cx_messy_one_point(a, b, rng):
    """tools.cxMessyOnePoint: one cut per parent, tails swapped; lengths change."""
    p1, p2 = rng.randint(0, len(a)), rng.randint(0, len(b))
    a[p1:], b[p2:] = b[p2:], a[p1:]
    return a, b


class ParetoFront:
    """tools.ParetoFront: the non-dominated individuals ever seen (equal genotypes kept once)."""

    def __init__(self):
        self.items = []

    def update(self, population):
        for ind in population:
            # an individual this front has already been offered is either in it or dominated by it, and stays so: skip
            # (the flag is not cloned: offspring are offered again)
            if getattr(ind, "_offered_to", None) is self:
                continue
            ind._offered_to = self
            dominated = False
            same = False
            keep = []
            for h in self.items:
                if not dominated and dominates(h.fitness, ind.fitness):
                    dominated = True
                if dominates(ind.fitness, h.fitness):
                    continue  # h leaves the front
                if tuple(h.fitness) == tuple(ind.fitness) and list(h) == list(ind):
                    same = True
                keep.append(h)
            if dominated:
                continue
            self.items = keep
            if not same:
                self.items.append(ind.clone())

    def __len__(self):
        return len(self.items)

    def __iter__(self):
        return iter(self.items)


def hypervolume_2d(points, ref):
    """Hypervolume of a 2-objective minimisation front w.r.t. ref (plots of the reference use HYPERVOLUME_REF_POINT)."""
    pts = sorted((p for p in points if p[0] <= ref[0] and p[1] <= ref[1]), key=lambda p: (p[0], p[1]))
    hv, best_y = 0.0, ref[1]
    for x, y in pts:
        if y < best_y:
            hv += (ref[0] - x) * (best_y - y)
            best_y = y
    return hv


# ---------------------------------------------------------------------------------------------------------------------
# operators (evolutionary_operators/mutationAndRecombination.py)
# ---------------------------------------------------------------------------------------------------------------------
def clean_up_individual(ind):
    """:64-87 -- consecutive identical viewpoints collapse to one."""
    out, prev = [], None
    for g in ind:
        if prev is None or g[0] != prev:
            out.append(g)
        prev = g[0]
    if len(out) != len(ind):
        ind[:] = out


def mut_polynomial_bounded(values, eta, low, up, indpb, rng):
    """tools.mutPolynomialBounded (Deb's polynomial mutation as implemented in NSGA-II / DEAP), in place."""
    for i, x in enumerate(values):
        if rng.random() > indpb:
            continue
        delta_1 = (x - low) / (up - low)
        delta_2 = (up - x) / (up - low)
        rand = rng.random()
        mut_pow = 1.0 / (eta + 1.0)
        if rand < 0.5:
            xy = 1.0 - delta_1
            val = 2.0 * rand + (1.0 - 2.0 * rand) * xy ** (eta + 1.0)
            delta_q = val ** mut_pow - 1.0
        else:
            xy = 1.0 - delta_2
            val = 2.0 * (1.0 - rand) + 2.0 * (rand - 0.5) * xy ** (eta + 1.0)
            delta_q = 1.0 - val ** mut_pow
        values[i] = min(max(x + delta_q * (up - low), low), up)
    return values


def mutate_viewing_angles(ind, params, rng):
    """:43-61 -- polynomial mutation of the angle offsets in [-1, 1] rad; True when something changed."""
    angles = [g[1] for g in ind]
    before = list(angles)
    mut_polynomial_bounded(angles, params.POLY_MUTATION_ETA, -1.0, 1.0, params.ANGLE_MUTATION_PROBABILITY, rng)
    for g, a in zip(ind, angles):
        g[1] = a
    return before != angles


def mutate_waypoints(ind, num_boxes, params, rng):
    """:20-41 -- insert a random viewpoint at a random place, or delete one (never below MIN_PLAN_SIZE)."""
    if rng.random() < 0.5:
        new_point = [rng.randrange(0, num_boxes)] + ([] if params.SIMPLIFIED_VIEW_ANGLE else [0.0])  # new points look straight ahead
        ind.insert(rng.randrange(len(ind) + 1), new_point)
        return True
    if len(ind) > params.MIN_PLAN_SIZE:
        ind.pop(rng.randrange(len(ind)))
        return True
    return False


# ---------------------------------------------------------------------------------------------------------------------
# the run
# ---------------------------------------------------------------------------------------------------------------------
def evaluate_batch(evaluator, inds, params, stats):
    """optimize_coverage.evaluate (:30-33) for a whole list: clean up, then ONE batched call."""
    if not inds:
        return
    for ind in inds:
        clean_up_individual(ind)
    t = time.perf_counter()
    fit = np.asarray(evaluator.evaluatePopulation([list(ind) for ind in inds], params.USING_EDGE_MEMOISATION, False))
    stats["eval_s"] += time.perf_counter() - t
    stats["evaluations"] += len(inds)
    stats["calls"] += 1
    for ind, f in zip(inds, fit):
        ind.fitness = (float(f[0]), float(f[1]))


def initial_population(evaluator, params, seeds, num_boxes, max_energy, rng, stats):
    """:75-92.  Random individuals are re-sampled until their energy is below the limit; candidates are evaluated in batches
    (the reference evaluates them one at a time -- same accepted distribution, fewer calls)."""
    pop = []
    slots = []  # index into pop of every slot still waiting for a valid random individual
    for _ in range(params.NUM_INDIVS):
        if seeds and rng.random() < params.USE_PLAN_SEED_CHANCE:
            pop.append(rng.choice(seeds).clone())
        else:
            pop.append(None)
            slots.append(len(pop) - 1)
    while slots:
        cands = [Individual.rand_init(rng.randint(params.MIN_PLAN_SIZE, params.MAX_PLAN_SIZE), num_boxes, rng, not params.SIMPLIFIED_VIEW_ANGLE)
                 for _ in slots]
        evaluate_batch(evaluator, cands, params, stats)
        left = []
        for s, c in zip(slots, cands):
            if c.fitness[1] < max_energy:
                pop[s] = c
            else:
                left.append(s)
        slots = left
    return pop


def run_nsga2(evaluator, params, seed=None, use_seeds=True, log=None, on_generation=None, on_seeds=None):
    """main_nsga2 (:112-213).  Returns (population, pareto_front, logbook, stats)."""
    rng = random.Random(seed)
    stats = {"evaluations": 0, "calls": 0, "eval_s": 0.0, "memoised_edges": 0}
    t0 = time.perf_counter()
    num_boxes = evaluator.getNumberOfBoxes()
    max_energy = evaluator.getMaxAllowedEnergy()
    seeds = []
    if use_seeds:
        for plan in evaluator.getSimplePlans():
            seeds.append(Individual([[int(g[0])] + ([] if params.SIMPLIFIED_VIEW_ANGLE else [0.0]) for g in plan]))  # initializeFromList
        evaluate_batch(evaluator, seeds, params, stats)  # storeSeedFitnesses :35-43
    stats["seed_s"] = time.perf_counter() - t0
    stats["seed_individuals"] = seeds
    if on_seeds and seeds:
        on_seeds(seeds, stats["seed_s"])

    pop = initial_population(evaluator, params, seeds, num_boxes, max_energy, rng, stats)
    evaluate_batch(evaluator, [i for i in pop if i.fitness is None], params, stats)
    pop = sel_nsga2(pop, len(pop))  # only assigns the crowding distances
    pf = ParetoFront()
    logbook = []

    def record(gen, evals):
        f = np.array([i.fitness for i in pop])
        sizes = np.array([len(i) for i in pop])
        row = {"gen": gen, "evals": evals, "fit_min": f.min(axis=0).tolist(), "fit_avg": f.mean(axis=0).tolist(),
               "fit_max": f.max(axis=0).tolist(), "size_min": int(sizes.min()), "size_avg": float(sizes.mean()), "size_max": int(sizes.max()),
               "memoised_edges": stats["memoised_edges"], "elapsed_s": time.perf_counter() - t0}
        logbook.append(row)
        if log:
            log("gen %4d evals %4d  score min %.4f avg %.4f  energy min %.2f avg %.2f  size avg %.1f  memo %d" % (
                gen, evals, row["fit_min"][0], row["fit_avg"][0], row["fit_min"][1], row["fit_avg"][1], row["size_avg"], row["memoised_edges"]))

    record(0, len(pop))
    for gen in range(params.NUM_GENERATIONS):
        offspring = [ind.clone() for ind in sel_tournament_dcd(pop, len(pop), rng)]
        if params.CROSSOVER_PROB > 0:
            for a, b in zip(offspring[::2], offspring[1::2]):
                if rng.random() <= params.CROSSOVER_PROB:
                    cx_messy_one_point(a, b, rng)
                    a.fitness = b.fitness = None
        if params.WAYPOINT_MUTATION_PROBABILITY > 0:
            for ind in offspring:
                if rng.random() <= params.WAYPOINT_MUTATION_PROBABILITY:
                    if mutate_waypoints(ind, num_boxes, params, rng):
                        ind.fitness = None
        if (not params.SIMPLIFIED_VIEW_ANGLE) and params.ANGLE_MUTATION_PROBABILITY > 0:
            for ind in offspring:
                if mutate_viewing_angles(ind, params, rng):
                    ind.fitness = None
        invalid = [ind for ind in offspring if ind.fitness is None]
        evaluate_batch(evaluator, invalid, params, stats)
        valid_offspring = [o for o in offspring if len(o) >= params.MIN_PLAN_SIZE]
        pop = sel_nsga2(pop + valid_offspring, params.NUM_INDIVS)
        pf.update(pop)
        if params.USING_EDGE_MEMOISATION:
            stats["memoised_edges"] = int(evaluator.updateMemoisedEdges([list(ind) for ind in pop]))
        record(gen, len(invalid))
        if on_generation:
            on_generation(gen, pop, pf)
    stats["total_s"] = time.perf_counter() - t0
    return pop, pf, logbook, stats


# ---------------------------------------------------------------------------------------------------------------------
# result stores (settings/Utilities.py:184-204, 288-313; settings/Constants_and_Datastructures.py:12-21, 59-66)
# ---------------------------------------------------------------------------------------------------------------------
def population_object_to_list(pop):
    return [{"fitness": list(ind.fitness) if ind.fitness is not None else None, "fitness_weights": list(FITNESS_WEIGHTS),
             "genotype1": [list(g) for g in ind], "genotype2": None, "type": "Individual"} for ind in pop]


def store_population_and_parameters(indivs, file_name, output_folder, scene, params, evaluator, pareto_front=None, elapsed_time=None,
                                    num_memoized=None):
    """Same dictionary keys as the reference's winner_indivs / genN / base_plans files (.yml or .pkl by extension)."""
    import pickle
    import yaml
    folder = os.path.join(output_folder, "populations")
    os.makedirs(folder, exist_ok=True)
    d = {"population": population_object_to_list(indivs), "scene": scene, "plan_origin": params.PLAN_ORIGIN,
         "sensor_params": list(params.SENSOR_PARAMETERS), "pareto_front": population_object_to_list(pareto_front) if pareto_front else None,
         "max_energy": float(evaluator.getMaxAllowedEnergy()), "num_memoized_solutions": num_memoized, "elapsed_time": elapsed_time,
         "plan_loops_around": params.PLAN_LOOPS_AROUND}
    path = os.path.join(folder, file_name)
    if file_name.endswith(".pkl"):
        with open(path, "wb") as f:
            pickle.dump(d, f)
    else:
        with open(path, "w") as f:
            yaml.safe_dump(d, f, default_flow_style=False, allow_unicode=True)
    return path


def load_dumped_population(path):
    import pickle
    import yaml
    if path.endswith(".pkl"):
        with open(path, "rb") as f:
            return pickle.load(f)
    with open(path) as f:
        return yaml.safe_load(f)


def make_evaluator(scene, params, post_processing=False):
    from . import cpp_binding
    return cpp_binding.PlanCoverageEstimator(scene, params.SENSOR_PARAMETERS, post_processing, params.PLAN_ORIGIN, params.PLAN_LOOPS_AROUND, True)


PARAMETERS_SUBFOLDER = "experiment_parameters"  # settings/Constants_and_Datastructures.py:46-50
SEED_PLAN_FILES = ("base_plans.pkl", "base_plans.yml")
FINAL_POPULATION_FILES = ("winner_indivs.pkl", "winner_indivs.yml")


def store_run_info(storage_folder, params):
    """Utilities.store_run_info (:84-90): the parameters of an experiment next to its runs."""
    folder = os.path.join(storage_folder, PARAMETERS_SUBFOLDER)
    os.makedirs(folder, exist_ok=True)
    return params.write_python(os.path.join(folder, "Parameters.py"))


def main(argv=None, evaluator_factory=None):
    """optimize_coverage.py:216-270.  evaluator_factory(scene, params, post_processing) replaces make_evaluator (tests)."""
    ap = argparse.ArgumentParser(description="NSGA-II inspection path planning on the batched B200 plan evaluator")
    ap.add_argument("-i", "--input_model_path", required=True)
    ap.add_argument("-o", "--output_folder", required=True)
    ap.add_argument("--disable_seeding", action="store_true")
    ap.add_argument("--multirun", action="store_true",
                    help="one run of an experiment: parameters come from <output_folder>/../experiment_parameters/Parameters.py, "
                         "only the population stores are written")
    ap.add_argument("--no_post_processing", action="store_true", help="only optimise and store the populations; no plan_ymls/")
    ap.add_argument("--params", help="a Parameters.py to take the parameters from")
    ap.add_argument("--generations", type=int)
    ap.add_argument("--pop", type=int)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--image_height", type=int)
    args = ap.parse_args(argv)
    make = evaluator_factory or make_evaluator
    if args.multirun:
        # one level above the run's folder (:223-227); the run folder itself may not exist yet
        params = Parameters.from_python(os.path.join(os.path.dirname(os.path.abspath(os.path.normpath(args.output_folder))),
                                                     PARAMETERS_SUBFOLDER, "Parameters.py"))
    elif args.params:
        params = Parameters.from_python(args.params)
    else:
        params = Parameters()
    if args.pop is not None:
        params.NUM_INDIVS = args.pop
    if args.generations is not None:
        params.NUM_GENERATIONS = args.generations
    if args.image_height is not None:
        sp = list(params.SENSOR_PARAMETERS)
        sp[3] = args.image_height
        params.SENSOR_PARAMETERS = sp
    os.makedirs(args.output_folder, exist_ok=True)
    t0 = time.time()
    ev = make(args.input_model_path, params, False)

    def on_seeds(seeds, elapsed):  # storeSeedFitnesses :35-43
        for name in SEED_PLAN_FILES:
            store_population_and_parameters(seeds, name, args.output_folder, args.input_model_path, params, ev, elapsed_time=elapsed)

    def on_gen(gen, pop, pf):
        if params.STORAGE_INTERVAL != -1 and gen % params.STORAGE_INTERVAL == 0:
            for ext in ("pkl", "yml"):
                store_population_and_parameters(pop, "gen%d.%s" % (gen, ext), args.output_folder, args.input_model_path, params, ev)

    pop, pf, logbook, stats = run_nsga2(ev, params, seed=args.seed, use_seeds=not args.disable_seeding, log=print, on_generation=on_gen,
                                        on_seeds=on_seeds)
    for name in FINAL_POPULATION_FILES:
        store_population_and_parameters(pop, name, args.output_folder, args.input_model_path, params, ev, pf, time.time() - t0,
                                        stats["memoised_edges"])
    hv = hypervolume_2d([i.fitness for i in pf], params.HYPERVOLUME_REF_POINT)
    print("pareto front %d plans, hypervolume %.4f; %d evaluations in %d batched calls, %.3f s in the evaluator, %.3f s total" % (
        len(pf), hv, stats["evaluations"], stats["calls"], stats["eval_s"], stats["total_s"]))
    if not args.multirun and not args.no_post_processing:
        # the reference also stores pictures of the plans here (GL viewer, out of scope); the waypoint files are exported
        from . import plan_exporter
        if hasattr(ev, "close"):
            ev.close()
        plan_exporter.export_all_winner_individuals(args.output_folder, evaluator=make(args.input_model_path, params, True))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
