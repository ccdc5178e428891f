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

Sharp-turn splitting (split_sharply_turning_waypoints, :34-81, `--min_turn_angle`): the reference calls a split_waypoint() that
is defined nowhere in its tree; split_waypoint() here does what the comment above its call site says -- "two new waypoints
instead of the old one, one split_distance metres before it, the other split_distance after", the edge between them heading
"as the average of the orientation before and after" -- with one guard the comment does not have: a new waypoint never moves
further than half of its edge away from the corner.
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


def angle_between(v1, v2):
    """settings/Utilities.py:213-230: angle in radians between two vectors (0 for a zero vector)."""
    n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
    if n1 == 0.0 or n2 == 0.0:
        return 0.0
    return float(np.arccos(np.clip(np.dot(v1 / n1, v2 / n2), -1.0, 1.0)))


def split_waypoint(incoming, outgoing, corner, heading_before, heading_after, split_distance):
    """The corner waypoint becomes two: one split_distance before it on the incoming edge, one split_distance after it on the
    outgoing edge (at most half of either edge); the edge between them looks along the average of the two headings."""
    li, lo = np.linalg.norm(incoming), np.linalg.norm(outgoing)
    before = corner - incoming / li * min(split_distance, 0.5 * li)
    after = corner + outgoing / lo * min(split_distance, 0.5 * lo)
    between = np.asarray(heading_before, dtype=np.float64) + np.asarray(heading_after, dtype=np.float64)
    n = np.linalg.norm(between)
    between = between / n if n > 0.0 else np.asarray(heading_before, dtype=np.float64)  # opposite headings: keep the first
    return before, after, between


def split_sharply_turning_waypoints(positions, headings, min_turn_angle=0.785398, split_distance=4.0):
    """plan_exporter.py:34-81.  headings[k] is the heading on the way to positions[k + 1]; the first and the last waypoint are
    never split.  Returns (positions, headings) with one more of each per split corner."""
    positions = [np.asarray(p, dtype=np.float64) for p in positions]
    new_pos, new_head = [positions[0]], []
    for k in range(len(headings) - 1):
        start, corner, nxt = positions[k], positions[k + 1], positions[k + 2]
        incoming, outgoing = corner - start, nxt - corner
        if angle_between(incoming, outgoing) > abs(min_turn_angle):
            before, after, between = split_waypoint(incoming, outgoing, corner, headings[k], headings[k + 1], split_distance)
            new_pos += [before, after]
            new_head += [np.asarray(headings[k], dtype=np.float64), between]
        else:
            new_pos.append(corner)
            new_head.append(np.asarray(headings[k], dtype=np.float64))
    new_pos.append(positions[-1])
    new_head.append(np.asarray(headings[-1], dtype=np.float64))
    return new_pos, new_head


def plan_to_waypoints(evaluator, plan, position_offsets=None, rotation_degrees_around_z=0.0, min_turn_angle=None):
    """-> list of {"position": {x, y, z}, "orientation": {r, p, y}} for one plan (list of [id] genes)."""
    if len(plan) == 0:
        return []
    positions, headings = evaluator.interpretAndExportPlan(plan, [], [])
    positions = [np.asarray(p, dtype=np.float64) for p in positions]
    headings = [np.asarray(h, dtype=np.float64) for h in headings]
    if min_turn_angle and len(headings) > 1:
        positions, headings = split_sharply_turning_waypoints(positions, headings, min_turn_angle)
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


def export_plan_to_yaml(evaluator, plan, yaml_path, position_offsets=None, rotation_degrees_around_z=0.0, min_turn_angle=None):
    import yaml
    wps = plan_to_waypoints(evaluator, plan, position_offsets, rotation_degrees_around_z, min_turn_angle)
    if not wps:
        return None
    with open(yaml_path, "w") as f:
        yaml.safe_dump({"waypoints": wps}, f, default_flow_style=False, allow_unicode=True)
    return yaml_path


def export_all_winner_individuals(results_root, position_offsets=None, rotation_degrees_around_z=0.0, evaluator=None,
                                  min_turn_angle=None):
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
                                         rotation_degrees_around_z, min_turn_angle))
    return paths


def main(argv=None):
    ap = argparse.ArgumentParser(description="Export generated plans as Rock waypoint YAML files")
    ap.add_argument("-i", "--input_folder", required=True)
    ap.add_argument("--position_offset", type=float, nargs=3)
    ap.add_argument("--rotation_offset", type=float, help="radians about +z")
    ap.add_argument("--min_turn_angle", type=float, help="radians; a turn sharper than this is split into two")
    a = ap.parse_args(argv)
    export_all_winner_individuals(a.input_folder, a.position_offset, math.degrees(a.rotation_offset) if a.rotation_offset else 0.0,
                                  min_turn_angle=a.min_turn_angle)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
