#!/usr/bin/env python
"""Plan export to the Rock waypoint YAML of the reference: plan_optimizer/plan_exporter.py:108-179.

    waypoints:
    - position: {x, y, z}
      orientation: {r: 0, p: 0, y: yaw}

The first waypoint carries a dummy orientation (the way to it is not part of the plan); every other waypoint carries the heading
followed on the way to it, as yaw only (settings/Utilities.py:268-284).  Optional rotation about +z (degrees) is applied to
positions and headings before the position offset, as the reference does.

Like the reference's exporter (evolutionary_operators/planEvaluatorSetup.py:13-24) export_all_winner_individuals builds the
evaluator with postProcessing=True: every exported heading comes from the heading search towards the primitives
(OsgHelpers.cpp:451-489: render both perpendicular directions along the edge, keep the one that shows more surface).

The reference's sharp-turn splitting
(split_sharply_turning_waypoints, :37-81) calls a split_waypoint() that is defined nowhere in its tree; it is not restated.
"""
import argparse
import math
import os

import numpy as np


def rotate_about_z(vectors, degrees):
    """settings/Utilities.py:253-266"""
    a = math.radians(degrees)
    c, s = math.cos(a), math.sin(a)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return [R @ np.asarray(v, dtype=np.float64) for v in vectors]


def plan_to_waypoints(evaluator, plan, position_offsets=None, rotation_degrees_around_z=0.0):
    """-> list of {"position": {x, y, z}, "orientation": {r, p, y}} for one plan (list of [id] genes)."""
    if len(plan) == 0:
        return []
    positions, headings = evaluator.interpretAndExportPlan(plan, [], [])
    positions = [np.asarray(p, dtype=np.float64) for p in positions]
    headings = [np.asarray(h, dtype=np.float64) for h in headings]
    if rotation_degrees_around_z:
        positions = rotate_about_z(positions, rotation_degrees_around_z)
        headings = rotate_about_z(headings, rotation_degrees_around_z)
    rpy = [[0, 0, math.atan2(h[1], h[0])] for h in headings]
    if position_offsets is not None:
        positions = [p + np.asarray(position_offsets, dtype=np.float64) for p in positions]
    rpy.insert(0, [0, 0, 0])  # dummy heading to the initial waypoint
    out = []
    for p, o in zip(positions, rpy):
        out.append({"position": {"x": float(p[0]), "y": float(p[1]), "z": float(p[2])},
                    "orientation": {"r": float(o[0]), "p": float(o[1]), "y": float(o[2])}})
    return out


def export_plan_to_yaml(evaluator, plan, yaml_path, position_offsets=None, rotation_degrees_around_z=0.0):
    import yaml
    wps = plan_to_waypoints(evaluator, plan, position_offsets, rotation_degrees_around_z)
    if not wps:
        return None
    with open(yaml_path, "w") as f:
        yaml.safe_dump({"waypoints": wps}, f, default_flow_style=False, allow_unicode=True)
    return yaml_path


def export_all_winner_individuals(results_root, position_offsets=None, rotation_degrees_around_z=0.0, evaluator=None):
    """plan_exporter.py:180-200: every plan of populations/winner_indivs.(pkl|yml) -> plan_ymls/plan_<i>.yml"""
    from . import optimize_coverage as oc
    pop_dir = os.path.join(results_root, "populations")
    src = os.path.join(pop_dir, "winner_indivs.pkl")
    if not os.path.exists(src):
        src = os.path.join(pop_dir, "winner_indivs.yml")
    data = oc.load_dumped_population(src)
    if evaluator is None:
        params = oc.Parameters(SENSOR_PARAMETERS=data["sensor_params"], PLAN_ORIGIN=data["plan_origin"], PLAN_LOOPS_AROUND=data["plan_loops_around"])
        evaluator = oc.make_evaluator(data["scene"], params, post_processing=True)
    out_dir = os.path.join(results_root, "plan_ymls")
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i, ind in enumerate(data["population"]):
        paths.append(export_plan_to_yaml(evaluator, ind["genotype1"], os.path.join(out_dir, "plan_%d.yml" % i), position_offsets,
                                         rotation_degrees_around_z))
    return paths


def main(argv=None):
    ap = argparse.ArgumentParser(description="Export generated plans as Rock waypoint YAML files")
    ap.add_argument("-i", "--input_folder", required=True)
    ap.add_argument("--position_offset", type=float, nargs=3)
    ap.add_argument("--rotation_offset", type=float, help="radians about +z")
    a = ap.parse_args(argv)
    export_all_winner_individuals(a.input_folder, a.position_offset, math.degrees(a.rotation_offset) if a.rotation_offset else 0.0)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
