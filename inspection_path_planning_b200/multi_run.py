#!/usr/bin/env python
"""One experiment run several times to gather statistics -- plan_optimizer/multiRun.py:14-27.

The parameters are written once to <output_folder>/experiment_parameters/Parameters.py (Utilities.store_run_info) and every run
reads them from there (`optimize_coverage --multirun`), so they cannot change between the runs; run i stores its populations
under <output_folder>/run<i>/populations/.  The reference starts one interpreter per run; here the runs share the process (the
CUDA context and the built library stay loaded), each with its own evaluator.

    python -m inspection_path_planning_b200.multi_run -i model.stl -o /tmp/experiment --runs 20
"""
import argparse
import os

from . import optimize_coverage as oc

NUM_RUNS = 20  # multiRun.py:12


def main(argv=None, evaluator_factory=None):
    ap = argparse.ArgumentParser(description="Runs the NSGA-II planner several times with one set of parameters")
    ap.add_argument("-i", "--input_model_path", required=True)
    ap.add_argument("-o", "--output_folder", required=True)
    ap.add_argument("--disable_seeding", action="store_true")
    ap.add_argument("--runs", type=int, default=NUM_RUNS)
    ap.add_argument("--params", help="a Parameters.py to take the parameters of the experiment from")
    ap.add_argument("--generations", type=int)
    ap.add_argument("--pop", type=int)
    ap.add_argument("--image_height", type=int)
    ap.add_argument("--seed", type=int, help="run i uses seed + i")
    args = ap.parse_args(argv)
    params = oc.Parameters.from_python(args.params) if args.params else oc.Parameters()
    if args.pop is not None:
        params.NUM_INDIVS = args.pop
    if args.generations is not None:
        params.NUM_GENERATIONS = args.generations
    if args.image_height is not None:
        sp = list(params.SENSOR_PARAMETERS)
        sp[3] = args.image_height
        params.SENSOR_PARAMETERS = sp
    print("Outfolder: ", args.output_folder)
    oc.store_run_info(args.output_folder, params)
    for i in range(args.runs):
        call = ["-i", args.input_model_path, "-o", os.path.join(args.output_folder, "run%d" % i) + os.sep, "--multirun"]
        if args.disable_seeding:
            call.append("--disable_seeding")
        if args.seed is not None:
            call += ["--seed", str(args.seed + i)]
        oc.main(call, evaluator_factory)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
