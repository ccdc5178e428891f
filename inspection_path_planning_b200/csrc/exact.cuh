// exact.cuh -- fp64 geometry of the reference, written once for host and device.
//
// Translation units that include this header are compiled with -fmad=false (device)
// and -ffp-contract=off (host) so every expression is the plain IEEE sequence of
// the reference's x86-64 build; results (collision flags, box classification,
// energy, energy gate) are then bit-identical to a CPU evaluation of the same
// formulas.  Citations: UF/ = plan_evaluator/Utility_Functions/src/,
// PE/ = plan_evaluator/Evaluator/src/ below /root/reference.
#pragma once
#include <math.h>

#include "common.cuh"

#ifdef __CUDACC__
#define IPP_HD __host__ __device__ __forceinline__
#else
#define IPP_HD inline
#endif

namespace ipp {

struct D3 {
    double x, y, z;
};
IPP_HD D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
IPP_HD D3 operator+(const D3& a, const D3& b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
IPP_HD D3 operator-(const D3& a, const D3& b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
IPP_HD D3 operator-(const D3& a) { return d3(-a.x, -a.y, -a.z); }
IPP_HD D3 operator*(const D3& a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
IPP_HD D3 operator/(const D3& a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
// osg::Vec3d::operator* (dot), operator^ (cross), length(), normalize()
IPP_HD double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
IPP_HD D3 cross(const D3& a, const D3& b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
IPP_HD double length(const D3& a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
IPP_HD void normalize(D3& a) {
    double n = length(a);
    if (n > 0.0) {
        double inv = 1.0 / n;
        a.x *= inv;
        a.y *= inv;
        a.z *= inv;
    }
}
IPP_HD bool equal(const D3& a, const D3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// osg::Vec3 (float) round trip
IPP_HD D3 round_to_float(const D3& a) { return d3((double)(float)a.x, (double)(float)a.y, (double)(float)a.z); }

struct PlaneD {
    double a, b, c, d;
};
// osg::Plane::set(v1, v2, v3)
IPP_HD PlaneD plane_from_points(const D3& v1, const D3& v2, const D3& v3) {
    D3 n = cross(v2 - v1, v3 - v2);
    double len = length(n);
    if (len > 1e-6) n = n / len; else n = d3(0, 0, 0);
    PlaneD p; p.a = n.x; p.b = n.y; p.c = n.z; p.d = -(dot(v1, n));
    return p;
}
// getPlane, UF/OsgHelpers.cpp:284-292
IPP_HD PlaneD get_plane(const D3& p1, const D3& p2, const D3& p3, const D3& dir) {
    PlaneD p = plane_from_points(p1, p2, p3);
    double dn = p.a * dir.x + p.b * dir.y + p.c * dir.z;
    if (dn < 0) { p.a = -p.a; p.b = -p.b; p.c = -p.c; p.d = -p.d; }
    return p;
}
IPP_HD double plane_dist(const PlaneD& p, const D3& v) { return p.a * v.x + p.b * v.y + p.c * v.z + p.d; }

// generatePolytopeFromPlanEdge, UF/OsgHelpers.cpp:195-244.  Also returns the 8 float-rounded corners'
// bounding box (for conservative BVH pruning).
IPP_HD void edge_prism(const D3& edgeStart, const D3& edgeEnd, PlaneD out[6], double lo[3], double hi[3]) {
    D3 dir = edgeEnd - edgeStart;
    normalize(dir);
    D3 up = d3(0, 0, 1);
    if (edgeEnd.x == edgeStart.x && edgeEnd.y == edgeStart.y) up = d3(1, 0, 0);
    D3 right = cross(dir, up);
    const double B = kEdgeSafetyBuffer;
    D3 c[8];
    c[0] = round_to_float(edgeStart - (up * B) - (right * B));  // startLowerLeft
    c[1] = round_to_float(edgeStart - (up * B) + (right * B));  // startLowerRight
    c[2] = round_to_float(edgeStart + (up * B) - (right * B));  // startUpperLeft
    c[3] = round_to_float(edgeStart + (up * B) + (right * B));  // startUpperRight
    c[4] = round_to_float(edgeEnd - (up * B) - (right * B));    // endLowerLeft
    c[5] = round_to_float(edgeEnd - (up * B) + (right * B));    // endLowerRight
    c[6] = round_to_float(edgeEnd + (up * B) - (right * B));    // endUpperLeft
    c[7] = round_to_float(edgeEnd + (up * B) + (right * B));    // endUpperRight
    out[0] = get_plane(c[4], c[5], c[7], -dir);   // far
    out[1] = get_plane(c[1], c[0], c[2], dir);    // near
    out[2] = get_plane(c[0], c[2], c[6], right);  // left
    out[3] = get_plane(c[1], c[3], c[7], -right); // right
    out[4] = get_plane(c[2], c[3], c[7], -up);    // top
    out[5] = get_plane(c[1], c[0], c[4], up);     // bottom
    lo[0] = lo[1] = lo[2] = 1e300;
    hi[0] = hi[1] = hi[2] = -1e300;
    for (int k = 0; k < 8; ++k) {
        lo[0] = fmin(lo[0], c[k].x); lo[1] = fmin(lo[1], c[k].y); lo[2] = fmin(lo[2], c[k].z);
        hi[0] = fmax(hi[0], c[k].x); hi[1] = fmax(hi[1], c[k].y); hi[2] = fmax(hi[2], c[k].z);
    }
}

// generateBoxPolytope, UF/OsgHelpers.cpp:295-318
IPP_HD void box_polytope(const D3& c, double s, PlaneD out[6], double lo[3], double hi[3]) {
    D3 r = d3(0, 1, 0), f = d3(1, 0, 0), u = d3(0, 0, 1);
    out[0] = get_plane(c + (f * s), c + (f * s) + r, c + (f * s) + u, -(f * s));
    out[1] = get_plane(c - (f * s), c - (f * s) + r, c - (f * s) + u, (f * s));
    out[2] = get_plane(c - (r * s), c + f - (r * s), c + u - (r * s), (r * s));
    out[3] = get_plane(c + (r * s), c + f + (r * s), c + u + (r * s), -(r * s));
    out[4] = get_plane(c + (u * s), c + f + (u * s), c + (u * s) - r, -(u * s));
    out[5] = get_plane(c - (u * s), c + f - (u * s), c - (u * s) - r, (u * s));
    lo[0] = c.x - s; lo[1] = c.y - s; lo[2] = c.z - s;
    hi[0] = c.x + s; hi[1] = c.y + s; hi[2] = c.z + s;
}

// Triangle vs convex polytope {x : dist_k(x) >= 0}, standing in for
// osgUtil::PolytopeIntersector::containsIntersections() (UF/OsgHelpers.cpp:249-252, 323-326).
// Sutherland-Hodgman clipping in fp64 after two vertex shortcuts.
IPP_HD bool polytope_hits_triangle(const PlaneD* pl, const D3& A, const D3& B, const D3& C) {
    bool all_in = true;
    for (int k = 0; k < 6; ++k) {
        double da = plane_dist(pl[k], A), db = plane_dist(pl[k], B), dc = plane_dist(pl[k], C);
        if (da < 0 && db < 0 && dc < 0) return false;
        if (!(da >= 0 && db >= 0 && dc >= 0)) all_in = false;
    }
    if (all_in) return true;
    D3 poly[10];
    D3 outp[10];
    double dist[10];
    int n = 3;
    poly[0] = A; poly[1] = B; poly[2] = C;
    for (int k = 0; k < 6; ++k) {
        for (int i = 0; i < n; ++i) dist[i] = plane_dist(pl[k], poly[i]);
        int m = 0;
        for (int i = 0; i < n; ++i) {
            int j = (i + 1 == n) ? 0 : i + 1;
            bool in_i = dist[i] >= 0, in_j = dist[j] >= 0;
            if (in_i) outp[m++] = poly[i];
            if (in_i != in_j) {
                double t = dist[i] / (dist[i] - dist[j]);
                outp[m++] = poly[i] + (poly[j] - poly[i]) * t;
            }
        }
        n = m;
        if (n == 0) return false;
        for (int i = 0; i < n; ++i) poly[i] = outp[i];
    }
    return true;
}

// calculateViewingDirection, UF/OsgHelpers.cpp:421-446
IPP_HD D3 viewing_direction(const D3& a, const D3& b, const D3& center) {
    D3 d;
    if (a.x == b.x && a.y == b.y) {
        d = center - ((b + a) / 2.0);
        d.z = 0;
    } else {
        d = cross(d3(0, 0, 1), b - a);
        D3 mid = b + (b - a) / 2.0;
        D3 want = center - mid;
        want.z = 0;
        if (dot(d, want) < 0) d = -d;
    }
    return d;
}
// calculateSensorDirectionAndNormalize, PE/PlanInterpreterBoxOrder.cpp:258-276.  offset = +infinity is the sentinel of the
// post-processing heading search (OsgHelpers.cpp:451-489): the other candidate, i.e. the quick heading reversed, no offset.
IPP_HD D3 sensor_direction(const D3& prev, const D3& next, const D3& center, float offset) {
    D3 d = viewing_direction(prev, next, center);
    if (offset > 3.0e38f) {
        d = -d;
        offset = 0.f;
    }
    normalize(d);
    double rad = atan2(d.y, d.x);
    rad += offset;
    d = d3(cos(rad), sin(rad), 0);
    normalize(d);
    return d;
}

// calculateEnergyUsageOneMove without the collision term, UF/OsgHelpers.cpp:131-151
IPP_HD double move_energy(const D3& s, const D3& e, const D3* prev) {
    double energy = kTranslationEnergyFactor * length(e - s);
    if (prev) {
        D3 in = s - *prev, out = e - s;
        double angleCos = dot(in, out) / (length(in) * length(out));
        double turn = 1.0 - angleCos;
        energy += turn * kRotationEnergyFactor;
    }
    return energy;
}

}  // namespace ipp
