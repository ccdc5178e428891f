#!/usr/bin/env python
"""Circling plans of a structure and their fitness (plan_optimizer/circling_plan_generator.py:16-38): getSimplePlans, one batched
evaluation (the reference calls evaluatePlan(p, False, nothing) per plan: memoisation off), circling_plans.{yml,pkl} with the
reference's dictionary keys.  The image export at the end of the reference script (GL viewer) is out of scope.

    python -m inspection_path_planning_b200.circling_plan_generator -i 3d_models/mockup_only.stl -o results/circling
"""
import argparse
import time

from . import optimize_coverage as oc


def generate(evaluator, params, scene, output_folder=None):
    t0 = time.time()
    inds = [oc.Individual([[int(g[0])] for g in plan]) for plan in evaluator.getSimplePlans()]
    fit = evaluator.evaluatePopulation([list(i) for i in inds], False, False)
    for ind, f in zip(inds, fit):
        ind.fitness = (float(f[0]), float(f[1]))
    elapsed = time.time() - t0
    if output_folder:
        for name in ("circling_plans.pkl", "circling_plans.yml"):
            oc.store_population_and_parameters(inds, name, output_folder, scene, params, evaluator, None, elapsed)
    return inds


def main(argv=None):
    ap = argparse.ArgumentParser(description="Generates the simple plans circling the inspection target and stores them with their fitness")
    ap.add_argument("-i", "--input_model_path", required=True)
    ap.add_argument("-o", "--output_folder", required=True)
    a = ap.parse_args(argv)
    params = oc.Parameters()
    ev = oc.make_evaluator(a.input_model_path, params)
    for ind in generate(ev, params, a.input_model_path, a.output_folder):
        print("%d viewpoints: score %.6f energy %.4f" % (len(ind), ind.fitness[0], ind.fitness[1]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
