// planner.cpp -- see planner.h
#include "planner.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "common.cuh"

namespace ipp {

void enumerate_grid_points(const float bbmin[3], const float bbmax[3], ViewpointGrid& g) {
    // osg::BoundingBox is float: the extents are float subtractions, widened afterwards
    const double ex = (double)(bbmax[0] - bbmin[0]), ey = (double)(bbmax[1] - bbmin[1]), ez = (double)(bbmax[2] - bbmin[2]);
    const double paddedVolume = (ex + 2 * kBoxPadding) * (ey + 2 * kBoxPadding) * (ez + kBoxPadding);
    const double interval = std::cbrt(paddedVolume / kNumViewpointsInsideVolume);
    const double minZ = bbmin[2] + kMinDistFromBottom;
    double maxZ = bbmax[2];
    if (kDownwardCameraActive) maxZ += kBoxPadding;
    g.points.clear();
    g.nz = g.nx = g.ny = 0;
    for (double z = minZ; z < maxZ; z += interval) {
        int nx = 0;
        for (double x = bbmin[0] - kBoxPadding; x < bbmax[0] + kBoxPadding; x += interval) {
            int ny = 0;
            for (double y = bbmin[1] - kBoxPadding; y < bbmax[1] + kBoxPadding; y += interval) {
                g.points.push_back(x);
                g.points.push_back(y);
                g.points.push_back(z);
                ++ny;
            }
            g.ny = ny;
            ++nx;
        }
        g.nx = nx;
        ++g.nz;
    }
}

void assign_box_ids(ViewpointGrid& g, const unsigned char* tooClose, const unsigned char* visible) {
    size_t n = g.points.size() / 3;
    g.cell.assign(n, -2);
    g.centres.clear();
    int next = 0;
    for (size_t i = 0; i < n; ++i) {
        if (tooClose[i]) {
            g.cell[i] = -1;
        } else if (visible[i]) {
            g.cell[i] = next++;
            g.centres.insert(g.centres.end(), g.points.begin() + 3 * i, g.points.begin() + 3 * i + 3);
        }
    }
}

std::vector<std::pair<int, int>> trace_boundary(const std::vector<std::vector<bool>>& img) {
    const int outer = (int)img.size(), inner = (int)img[0].size();
    // one-cell frame of false around the image; column = outer index, row = inner index
    const int cols = outer + 2, rows = inner + 2;
    std::vector<unsigned char> pix((size_t)cols * rows, 0), seen((size_t)cols * rows, 0);
    for (int r = 0; r < inner; ++r)
        for (int c = 0; c < outer; ++c) pix[(size_t)(r + 1) * cols + (c + 1)] = img[c][r] ? 1 : 0;
    // eight neighbours clockwise from west, and the neighbour to resume from after a hit (1-based)
    static const int dc[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    static const int dr[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    static const int resume[8] = {7, 7, 1, 1, 3, 3, 5, 5};
    std::vector<std::pair<int, int>> out;
    bool latched = false;  // the reference's `inside` flag: once a closed contour is found it never clears
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            if (latched) continue;
            size_t here = (size_t)r * cols + c;
            if (seen[here] || !pix[here]) continue;
            out.emplace_back(r - 1, c - 1);
            seen[here] = 1;
            int cr = r, cc = c, look = 1, closures = 0, misses = 0;
            for (;;) {
                int nr = cr + dr[look - 1], nc = cc + dc[look - 1];
                int nextLook = resume[look - 1];
                if (pix[(size_t)nr * cols + nc]) {
                    out.emplace_back(nr - 1, nc - 1);
                    if (nr == r && nc == c) {
                        ++closures;
                        if (nextLook == 1 || closures >= 3) {  // Jacob's stopping criterion
                            latched = true;
                            break;
                        }
                    }
                    look = nextLook;
                    cr = nr;
                    cc = nc;
                    misses = 0;
                    seen[(size_t)nr * cols + nc] = 1;
                } else {
                    look = 1 + (look % 8);
                    if (misses > 8) break;  // isolated pixel
                    ++misses;
                }
            }
        }
    }
    return out;
}

std::vector<Plan> single_level_circling_plans(const ViewpointGrid& g) {
    std::vector<Plan> levels;
    for (int z = 0; z < g.nz; ++z) {
        // layer[a][b] with a = first grid index (x loop), b = second (y loop); the reference calls them y and x
        std::vector<std::vector<bool>> rim((size_t)g.nx, std::vector<bool>((size_t)g.ny, false));
        auto touches = [&](int a, int b) {
            for (int db = -1; db <= 1; ++db)
                for (int da = -1; da <= 1; ++da) {
                    int aa = a + da, bb = b + db;
                    if (aa < 0 || aa >= g.nx || bb < 0 || bb >= g.ny) continue;
                    if (g.at(z, aa, bb) == -1) return true;
                }
            return false;
        };
        for (int a = 0; a < g.nx; ++a)
            for (int b = 0; b < g.ny; ++b) rim[a][b] = g.at(z, a, b) == -1 || touches(a, b);
        std::vector<std::pair<int, int>> tr = trace_boundary(rim);
        if (tr.empty()) continue;
        Plan plan;
        auto traced = [&](int b, int a) { return std::find(tr.begin(), tr.end(), std::make_pair(b, a)) != tr.end(); };
        for (const auto& t : tr) {
            int b = t.first, a = t.second;
            bool straightB = traced(b - 1, a) && traced(b + 1, a);
            bool straightA = traced(b, a - 1) && traced(b, a + 1);
            if (!(straightB || straightA)) plan.push_back(std::vector<double>{(double)g.at(z, a, b)});
        }
        levels.push_back(plan);
    }
    return levels;
}

std::vector<Plan> complete_circling_plans(const std::vector<Plan>& levels) {
    std::vector<Plan> out;
    const int total = (int)levels.size();
    for (int step = 1; step <= total; ++step) {
        std::vector<int> pick;
        for (int lv = 0; lv < total; lv += step) pick.push_back(lv);
        // shift the selection so that it is centred on the middle of the structure
        const double middle = (double)total / 2.0;
        const int mid = pick[pick.size() / 2];
        int skew;
        if (pick.size() % 2 == 1) skew = (int)(middle - mid);
        else skew = (int)(middle - (double)(mid + pick[pick.size() / 2 - 1]) / 2.0);
        Plan plan;
        for (int lv : pick) {
            int l = lv + skew;
            if (l < 0 || l >= total) throw std::logic_error("circling plan level out of range");
            plan.insert(plan.end(), levels[l].begin(), levels[l].end());
        }
        out.push_back(plan);
    }
    return out;
}

}  // namespace ipp
