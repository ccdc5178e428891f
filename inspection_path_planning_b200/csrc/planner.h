// planner.h -- host-side, setup-time pieces of the plan interpreter: candidate-viewpoint grid
// enumeration, circling seed plans (Moore contour tracing) and plan decoding for export.
// Reference: plan_evaluator/Evaluator/src/PlanInterpreterBoxOrder.cpp, ContourTracing.cpp.
#pragma once
#include <utility>
#include <vector>

namespace ipp {

struct ViewpointGrid {
    int nz = 0, nx = 0, ny = 0;
    std::vector<double> points;  // every grid point, xyz, loop order z -> x -> y
    std::vector<int> cell;       // [z][x][y]: -1 too close, -2 nothing visible, else box id
    std::vector<double> centres; // accepted viewpoints, xyz, id order
    int at(int z, int x, int y) const { return cell[((size_t)z * nx + x) * ny + y]; }
};

// divideIntoBoxes loop structure (PlanInterpreterBoxOrder.cpp:193-251): fills g.points and the dims
void enumerate_grid_points(const float bbmin[3], const float bbmax[3], ViewpointGrid& g);
// ids in loop order from the two cube tests per point
void assign_box_ids(ViewpointGrid& g, const unsigned char* tooClose, const unsigned char* visible);

// Moore-neighbour boundary trace of a boolean image indexed [outer][inner] (ContourTracing.cpp:70-252).
// Returns (inner, outer) index pairs in trace order.
std::vector<std::pair<int, int>> trace_boundary(const std::vector<std::vector<bool>>& img);

typedef std::vector<std::vector<double>> Plan;  // genes, each [id] or [id, angle]
// ProduceSimpleCirclingPlans (:40-109)
std::vector<Plan> single_level_circling_plans(const ViewpointGrid& g);
// generateOrGetCompleteCirclingPlans (:111-160)
std::vector<Plan> complete_circling_plans(const std::vector<Plan>& levels);

}  // namespace ipp
