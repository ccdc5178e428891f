// exact.cu -- fp64 kernels whose results must equal the reference arithmetic bit for bit:
// mesh preparation (K0), polytope-vs-mesh queries through the BVH (K2 box classification,
// K5 edge collision), edge decoding into camera poses (K3) and plan energy + gate.
// Compiled with -fmad=false.  See exact.cuh for the formulas and their citations.
#include "exact.cuh"
#include "kernels.h"

namespace ipp {

// ------------------------------------------------------------------------------------------
// K0 mesh_prepare: Heron area (TriangleData.cpp:56-61), degenerate flag, canonical vertex order,
// centroid for the Morton sort.
// ------------------------------------------------------------------------------------------
__global__ void k_mesh_prepare(const float* __restrict__ v, int T, double* __restrict__ area, uint8_t* __restrict__ degenerate,
                               float* __restrict__ canon, float* __restrict__ centroid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    float p[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = v[(size_t)i * 9 + k];
    D3 v1 = d3(p[0], p[1], p[2]), v2 = d3(p[3], p[4], p[5]), v3 = d3(p[6], p[7], p[8]);
    double s1 = length(v2 - v1), s2 = length(v3 - v1), s3 = length(v3 - v2);
    double hp = (s1 + s2 + s3) / 2.0;
    area[i] = sqrt(hp * (hp - s1) * (hp - s2) * (hp - s3));
    // canonical order: sort the three vertices lexicographically (x, y, z)
    int o0 = 0, o1 = 1, o2 = 2;
    auto less = [&](int a, int b) {
        if (p[3 * a] != p[3 * b]) return p[3 * a] < p[3 * b];
        if (p[3 * a + 1] != p[3 * b + 1]) return p[3 * a + 1] < p[3 * b + 1];
        return p[3 * a + 2] < p[3 * b + 2];
    };
    if (less(o1, o0)) { int t = o0; o0 = o1; o1 = t; }
    if (less(o2, o1)) { int t = o1; o1 = o2; o2 = t; }
    if (less(o1, o0)) { int t = o0; o0 = o1; o1 = t; }
    int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) canon[(size_t)i * 9 + 3 * k + c] = p[3 * ord[k] + c];
    D3 c0 = d3(p[3 * o0], p[3 * o0 + 1], p[3 * o0 + 2]), c1 = d3(p[3 * o1], p[3 * o1 + 1], p[3 * o1 + 2]),
       c2 = d3(p[3 * o2], p[3 * o2 + 1], p[3 * o2 + 2]);
    D3 n = cross(c1 - c0, c2 - c0);
    degenerate[i] = (n.x == 0 && n.y == 0 && n.z == 0) ? 1 : 0;
    centroid[(size_t)i * 3 + 0] = (p[0] + p[3] + p[6]) * (1.0f / 3.0f);
    centroid[(size_t)i * 3 + 1] = (p[1] + p[4] + p[7]) * (1.0f / 3.0f);
    centroid[(size_t)i * 3 + 2] = (p[2] + p[5] + p[8]) * (1.0f / 3.0f);
}

void launch_mesh_prepare(const float* d_verts, int T, double* d_area, uint8_t* d_degenerate, float* d_canon, float* d_centroid,
                         cudaStream_t st) {
    k_mesh_prepare<<<(T + 255) / 256, 256, 0, st>>>(d_verts, T, d_area, d_degenerate, d_canon, d_centroid);
}

// ------------------------------------------------------------------------------------------
// Polytope-vs-mesh any-hit query through the BVH (one thread per query, early exit).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool box_may_hit(const PlaneD pl[6], const double lo[3], const double hi[3], float bx0, float by0, float bz0,
                                            float bx1, float by1, float bz1) {
    const double pad = 1e-3;  // the polytope may exceed the bounding box of its float-rounded corners by rounding only
    if ((double)bx1 < lo[0] - pad || (double)bx0 > hi[0] + pad || (double)by1 < lo[1] - pad || (double)by0 > hi[1] + pad ||
        (double)bz1 < lo[2] - pad || (double)bz0 > hi[2] + pad)
        return false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double px = pl[k].a >= 0 ? (double)bx1 : (double)bx0;
        double py = pl[k].b >= 0 ? (double)by1 : (double)by0;
        double pz = pl[k].c >= 0 ? (double)bz1 : (double)bz0;
        if (pl[k].a * px + pl[k].b * py + pl[k].c * pz + pl[k].d < 0) return false;
    }
    return true;
}

__device__ bool polytope_hits_mesh(const DeviceScene& sc, const PlaneD pl[6], const double lo[3], const double hi[3], int root = 0) {
    if (sc.T == 0) return false;
    int stack[kMaxDepth];
    int sp = 0;
    int node = root;
    while (true) {
        const BvhNode* np = sc.nodes + node;
        float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
        int link[2] = {__float_as_int(q3.x), __float_as_int(q3.y)};
        bool hit[2];
        hit[0] = box_may_hit(pl, lo, hi, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
        hit[1] = box_may_hit(pl, lo, hi, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
        int next = -1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!hit[c]) continue;
            if (link[c] < 0) {
                int code = ~link[c];
                int first = code >> kLeafShift, cnt = code & kLeafMask;
                for (int k = 0; k < cnt; ++k) {
                    float4 a = __ldg(sc.oa + first + k), b = __ldg(sc.ob + first + k), cc = __ldg(sc.oc + first + k);
                    if (polytope_hits_triangle(pl, d3(a.x, a.y, a.z), d3(b.x, b.y, b.z), d3(cc.x, cc.y, cc.z))) return true;
                }
            } else {
                if (next < 0) next = link[c];
                else if (sp < kMaxDepth) stack[sp++] = link[c];
            }
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return false;
}

// K5 edge_collision for memo keys: key = prevIdx * N + currIdx, prevIdx == N means the start location
__global__ void k_collide_keys(DeviceScene sc, const int* __restrict__ keys, int n, const double* __restrict__ pos, int N,
                               uint8_t* __restrict__ coll) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keys[i];
    int a = key / N, b = key - a * N;
    D3 A = d3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]), B = d3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
    uint8_t r = 0;
    if (!equal(A, B)) {
        PlaneD pl[6];
        double lo[3], hi[3];
        edge_prism(A, B, pl, lo, hi);
        r = polytope_hits_mesh(sc, pl, lo, hi) ? 1 : 0;
    }
    coll[key] = r;
}
// The same query with one warp per key, for the launches that have only a handful of keys (a generation of an NSGA-II run brings a
// few new edges at a time, and one thread walking the BVH alone is all such a launch would last): the warp keeps a frontier of
// BVH nodes in shared memory, every lane takes one node per round, tests both child boxes, tests the triangles of a leaf child
// itself and pushes an inner child back.  Any-hit: the answer does not depend on the order, so it is the one-thread answer.
constexpr int kCollideWarps = 4;
constexpr int kFrontierCap = 512;
__global__ void __launch_bounds__(kCollideWarps * 32) k_collide_keys_warp(DeviceScene sc, const int* __restrict__ keys, int n,
                                                                          const double* __restrict__ pos, int N, uint8_t* __restrict__ coll) {
    __shared__ int s_frontier[kCollideWarps][kFrontierCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kCollideWarps + warp;
    if (i >= n) return;
    const int key = keys[i];
    const int a = key / N, b = key - a * N;
    const D3 A = d3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]), B = d3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
    bool found = false;
    if (!equal(A, B) && sc.T > 0) {
        PlaneD pl[6];
        double lo[3], hi[3];
        edge_prism(A, B, pl, lo, hi);
        int* fr = s_frontier[warp];
        if (lane == 0) fr[0] = 0;
        int count = 1;
        __syncwarp();
        while (count > 0 && !found) {
            const int take = min(count, 32);
            const int node = lane < take ? fr[count - 1 - lane] : -1;
            count -= take;
            __syncwarp();
            int child[2] = {-1, -1};
            bool hit = false;
            if (node >= 0) {
                const BvhNode* np = sc.nodes + node;
                const float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
                const int link[2] = {__float_as_int(q3.x), __float_as_int(q3.y)};
                const bool bh[2] = {box_may_hit(pl, lo, hi, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y),
                                    box_may_hit(pl, lo, hi, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w)};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (!bh[c]) continue;
                    if (link[c] < 0) {
                        const int code = ~link[c];
                        const int first = code >> kLeafShift, cnt = code & kLeafMask;
                        for (int k = 0; k < cnt && !hit; ++k) {
                            const float4 va = __ldg(sc.oa + first + k), vb = __ldg(sc.ob + first + k), vc = __ldg(sc.oc + first + k);
                            hit = polytope_hits_triangle(pl, d3(va.x, va.y, va.z), d3(vb.x, vb.y, vb.z), d3(vc.x, vc.y, vc.z));
                        }
                    } else {
                        child[c] = link[c];
                    }
                }
            }
            found = __any_sync(0xffffffffu, hit);
            if (found) break;
            // push the inner children; a frontier that would overflow is walked by its lanes on their own instead
            const unsigned m0 = __ballot_sync(0xffffffffu, child[0] >= 0), m1 = __ballot_sync(0xffffffffu, child[1] >= 0);
            const int add = __popc(m0) + __popc(m1);
            if (count + add > kFrontierCap) {
                bool h = false;
                if (child[0] >= 0) h = polytope_hits_mesh(sc, pl, lo, hi, child[0]);
                if (!h && child[1] >= 0) h = polytope_hits_mesh(sc, pl, lo, hi, child[1]);
                found = __any_sync(0xffffffffu, h);
            } else {
                const unsigned lt = (1u << lane) - 1u;
                if (child[0] >= 0) fr[count + __popc(m0 & lt)] = child[0];
                if (child[1] >= 0) fr[count + __popc(m0) + __popc(m1 & lt)] = child[1];
                count += add;
            }
            __syncwarp();
        }
    }
    if (lane == 0) coll[key] = found ? 1 : 0;
}
__global__ void k_collide_xyz(DeviceScene sc, const double* __restrict__ edges, int64_t n, uint8_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 A = d3(edges[6 * i], edges[6 * i + 1], edges[6 * i + 2]), B = d3(edges[6 * i + 3], edges[6 * i + 4], edges[6 * i + 5]);
    PlaneD pl[6];
    double lo[3], hi[3];
    edge_prism(A, B, pl, lo, hi);
    out[i] = polytope_hits_mesh(sc, pl, lo, hi) ? 1 : 0;
}
// K2 classify_boxes: cube of half side s around each centre
__global__ void k_boxes_hit(DeviceScene sc, const double* __restrict__ centres, int64_t n, double s, uint8_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    PlaneD pl[6];
    double lo[3], hi[3];
    box_polytope(d3(centres[3 * i], centres[3 * i + 1], centres[3 * i + 2]), s, pl, lo, hi);
    out[i] = polytope_hits_mesh(sc, pl, lo, hi) ? 1 : 0;
}

void launch_collide_keys(const DeviceScene& sc, const int* d_keys, int n, const double* d_pos, int N, uint8_t* d_coll, cudaStream_t st) {
    if (n <= 0) return;
    // a few keys: a warp each (same answers, a fraction of the latency); many: a thread each fills the machine
    if (n <= 2048) k_collide_keys_warp<<<(n + kCollideWarps - 1) / kCollideWarps, kCollideWarps * 32, 0, st>>>(sc, d_keys, n, d_pos, N, d_coll);
    else k_collide_keys<<<(n + 63) / 64, 64, 0, st>>>(sc, d_keys, n, d_pos, N, d_coll);
}
void launch_collide_xyz(const DeviceScene& sc, const double* d_edges, int64_t n, uint8_t* d_out, cudaStream_t st) {
    if (n > 0) k_collide_xyz<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(sc, d_edges, n, d_out);
}
void launch_boxes_hit(const DeviceScene& sc, const double* d_centres, int64_t n, double s, uint8_t* d_out, cudaStream_t st) {
    if (n > 0) k_boxes_hit<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(sc, d_centres, n, s, d_out);
}

// ------------------------------------------------------------------------------------------
// K3 decode_edges: memo key -> heading, sample positions (CameraEstimator.cpp:83-97), camera poses
// (:149-167) and the macro-tile rectangle that can see the scene (camera-frustum culling).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int edge_samples(const D3& a, const D3& b, double interval, float* out /*nullable*/, int cap) {
    if (equal(a, b)) {
        if (out) { out[0] = (float)a.x; out[1] = (float)a.y; out[2] = (float)a.z; }
        return 1;
    }
    int n = 0;
    D3 mv = b - a;
    if (out && n < cap) { out[0] = (float)a.x; out[1] = (float)a.y; out[2] = (float)a.z; }
    n++;
    D3 next = a + (mv * interval) / double(length(mv));
    while (length(next - a) < length(mv)) {
        if (out && n < cap) { out[3 * n] = (float)next.x; out[3 * n + 1] = (float)next.y; out[3 * n + 2] = (float)next.z; }
        n++;
        next = next + (mv * interval) / length(mv);
    }
    if (out && n < cap) { out[3 * n] = (float)b.x; out[3 * n + 1] = (float)b.y; out[3 * n + 2] = (float)b.z; }
    n++;
    return n;
}

__device__ void make_snapshot(Snapshot& S, const float eye[3], const D3& center, const D3& up, int slot, const CameraParams& cam,
                              const float* bbmin, const float* bbmax) {
    D3 e = d3(eye[0], eye[1], eye[2]);
    // osg::Matrixd::makeLookAt
    D3 f = center - e;
    normalize(f);
    D3 s = cross(f, up);
    normalize(s);
    D3 u = cross(s, f);
    normalize(u);
    S.f[0] = f.x; S.f[1] = f.y; S.f[2] = f.z;
    S.s[0] = s.x; S.s[1] = s.y; S.s[2] = s.z;
    S.u[0] = u.x; S.u[1] = u.y; S.u[2] = u.z;
    S.eye[0] = eye[0]; S.eye[1] = eye[1]; S.eye[2] = eye[2];
    S.slot = slot;
    S.slotCount = 1;
    // conservative screen rectangle of the (slightly padded) scene box
    const int mtx = (cam.W + 31) / 32, mty = (cam.H + 31) / 32;
    double pad = 1e-3;
    bool anyBehind = false, allBehind = true, allBeyond = true;
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int k = 0; k < 8; ++k) {
        D3 p = d3((k & 1) ? bbmax[0] + pad : bbmin[0] - pad, (k & 2) ? bbmax[1] + pad : bbmin[1] - pad,
                  (k & 4) ? bbmax[2] + pad : bbmin[2] - pad) - e;
        double x = dot(p, s), y = dot(p, u), z = dot(p, f);
        double zn = (double)cam.zNear * 0.999;
        if (z < zn) anyBehind = true; else allBehind = false;
        if (z <= (double)cam.zFar * 1.001) allBeyond = false;
        if (z >= zn) {
            double X = x / z, Y = y / z;
            xmin = fmin(xmin, X); xmax = fmax(xmax, X);
            ymin = fmin(ymin, Y); ymax = fmax(ymax, Y);
        }
    }
    S.rx0 = S.ry0 = 0;
    S.rnx = S.rny = 0;
    if (allBehind || allBeyond) return;
    int px0 = 0, px1 = cam.W - 1, py0 = 0, py1 = cam.H - 1;
    if (!anyBehind) {
        double fx0 = (xmin / cam.tanX + 1.0) * 0.5 * cam.W - 0.5, fx1 = (xmax / cam.tanX + 1.0) * 0.5 * cam.W - 0.5;
        double fy0 = (ymin / cam.tanY + 1.0) * 0.5 * cam.H - 0.5, fy1 = (ymax / cam.tanY + 1.0) * 0.5 * cam.H - 0.5;
        if (fx1 < -2.0 || fy1 < -2.0 || fx0 > cam.W + 1.0 || fy0 > cam.H + 1.0) return;
        px0 = max(0, (int)floor(fx0) - 1);
        px1 = min(cam.W - 1, (int)ceil(fx1) + 1);
        py0 = max(0, (int)floor(fy0) - 1);
        py1 = min(cam.H - 1, (int)ceil(fy1) + 1);
        if (px0 > px1 || py0 > py1) return;
    }
    S.rx0 = px0 / 32;
    S.ry0 = py0 / 32;
    S.rnx = min(mtx - 1, px1 / 32) - S.rx0 + 1;
    S.rny = min(mty - 1, py1 / 32) - S.ry0 + 1;
}

__device__ __forceinline__ void key_endpoints(int key, float angle, const double* pos, int N, D3& A, D3& B) {
    int a = key / N, b = key - a * N;
    A = d3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]);
    B = d3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]);
}

__global__ void k_count_snapshots(const int* __restrict__ keys, int n, const double* __restrict__ pos, int N, CameraParams cam,
                                  int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 A, B;
    key_endpoints(keys[i], 0.f, pos, N, A, B);
    int ns = edge_samples(A, B, cam.sampleInterval, nullptr, 0);
    counts[i] = ns * ((cam.front ? 1 : 0) + (cam.down ? 1 : 0));
}

__global__ void k_fill_snapshots(const int* __restrict__ keys, const float* __restrict__ angles, int n, const double* __restrict__ pos,
                                 int N, D3 center, CameraParams cam, DeviceScene sc, const int* __restrict__ snapOffset,
                                 Snapshot* __restrict__ snaps, int* __restrict__ ntiles) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 A, B;
    key_endpoints(keys[i], 0.f, pos, N, A, B);
    float off = angles ? angles[i] : 0.f;
    D3 heading = sensor_direction(A, B, center, off);
    float hf[3] = {(float)heading.x, (float)heading.y, (float)heading.z};
    int o = snapOffset[i];
    // regenerate the sample positions one at a time (same arithmetic as edge_samples)
    auto emit = [&](const D3& p) {
        float eye[3] = {(float)p.x, (float)p.y, (float)p.z};
        if (cam.front) {
            // point + robotHeading is an osg::Vec3 (float) sum, CameraEstimator.cpp:151
            float c[3] = {eye[0] + hf[0], eye[1] + hf[1], eye[2] + hf[2]};
            make_snapshot(snaps[o], eye, d3(c[0], c[1], c[2]), d3(0, 0, 1), i, cam, sc.bbmin, sc.bbmax);
            ntiles[o] = snaps[o].rnx * snaps[o].rny;
            o++;
        }
        if (cam.down) {
            float c[3] = {eye[0] - 0.0f, eye[1] - 0.0f, eye[2] - 1.0f};  // point - UP_VECTOR in float, :162
            make_snapshot(snaps[o], eye, d3(c[0], c[1], c[2]), heading, i, cam, sc.bbmin, sc.bbmax);
            ntiles[o] = snaps[o].rnx * snaps[o].rny;
            o++;
        }
    };
    if (equal(A, B)) {
        emit(A);
        return;
    }
    D3 mv = B - A;
    emit(A);
    D3 next = A + (mv * cam.sampleInterval) / double(length(mv));
    while (length(next - A) < length(mv)) {
        emit(next);
        next = next + (mv * cam.sampleInterval) / length(mv);
    }
    emit(B);
}

__global__ void k_single_snapshot(Snapshot* snap, D3 eye, D3 center, D3 up, CameraParams cam, DeviceScene sc, int* ntiles) {
    float e[3] = {(float)eye.x, (float)eye.y, (float)eye.z};
    make_snapshot(*snap, e, center, up, 0, cam, sc.bbmin, sc.bbmax);
    // test hook renders every macro tile so that background pixels are reported too
    snap->rx0 = 0;
    snap->ry0 = 0;
    snap->rnx = (cam.W + 31) / 32;
    snap->rny = (cam.H + 31) / 32;
    *ntiles = snap->rnx * snap->rny;
}

void launch_count_snapshots(const int* d_keys, int n, const double* d_pos, int N, const CameraParams& cam, int* d_counts, cudaStream_t st) {
    if (n > 0) k_count_snapshots<<<(n + 127) / 128, 128, 0, st>>>(d_keys, n, d_pos, N, cam, d_counts);
}
void launch_fill_snapshots(const int* d_keys, const float* d_angles, int n, const double* d_pos, int N, const double center[3],
                           const CameraParams& cam, const DeviceScene& sc, const int* d_snapOffset, Snapshot* d_snaps, int* d_ntiles,
                           cudaStream_t st) {
    if (n > 0)
        k_fill_snapshots<<<(n + 63) / 64, 64, 0, st>>>(d_keys, d_angles, n, d_pos, N, d3(center[0], center[1], center[2]), cam, sc,
                                                       d_snapOffset, d_snaps, d_ntiles);
}
void launch_single_snapshot(Snapshot* d_snap, const double eye[3], const double center[3], const double up[3], const CameraParams& cam,
                            const DeviceScene& sc, int* d_ntiles, cudaStream_t st) {
    k_single_snapshot<<<1, 1, 0, st>>>(d_snap, d3(eye[0], eye[1], eye[2]), d3(center[0], center[1], center[2]), d3(up[0], up[1], up[2]),
                                       cam, sc, d_ntiles);
}

// ------------------------------------------------------------------------------------------
// Plan-level bookkeeping: claim unknown collision flags / edge rows, energy + gate.
// Position sequence of plan p: [start] g_0 .. g_{n-1} [g_0 if planLoopsAround]
// (PlanCoverageEstimator.cpp:216-220, PlanInterpreterBoxOrder.cpp:301-324).
// ------------------------------------------------------------------------------------------
struct PlanIter {
    const int* ids;
    int64_t lo, hi;
    int N;
    bool hasStart, loops;
    int64_t k;  // position cursor
    __device__ int count() const {
        int64_t n = hi - lo;
        if (n == 0) return 0;
        return (int)(n + (hasStart ? 1 : 0) + (loops ? 1 : 0));
    }
    // box index of position k (N = start location)
    __device__ int at(int k) const {
        int64_t n = hi - lo;
        if (hasStart) {
            if (k == 0) return N;
            k -= 1;
        }
        if (k < n) return ids[lo + k];
        return ids[lo];  // loop closure
    }
};

__global__ void k_mark_collisions(PlanMarkArgs a) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.pop) return;
    PlanIter it{a.ids, a.offsets[p], a.offsets[p + 1], a.N, a.hasStart, a.loops, 0};
    int m = it.count();
    // box ids are checked before anything is claimed (assert at PlanInterpreterBoxOrder.cpp:282): a plan with a bad id claims
    // nothing, and the host resolves what the other plans claimed before it reports the error -- no flag is left pending
    for (int64_t g = a.offsets[p]; g < a.offsets[p + 1]; ++g)
        if ((unsigned)a.ids[g] >= (unsigned)a.N) { *a.errFlag = 1; return; }
    int prev = m > 0 ? it.at(0) : 0;
    for (int k = 1; k < m; ++k) {
        int cur = it.at(k);
        int key = prev * a.N + cur;
        if (a.coll[key] == 0xFF) {
            // claim: 0xFF unknown -> 0xFE pending (byte CAS through the containing word)
            unsigned* w = (unsigned*)(a.coll + (key & ~3));
            unsigned sh = (key & 3) * 8;
            unsigned old = *w;
            while (((old >> sh) & 0xFF) == 0xFF) {
                unsigned assumed = old;
                unsigned nv = (old & ~(0xFFu << sh)) | (0xFEu << sh);
                old = atomicCAS(w, assumed, nv);
                if (old == assumed) {
                    int slot = atomicAdd(a.newCount, 1);
                    a.newKeys[slot] = key;
                    break;
                }
            }
        }
        prev = cur;
    }
}

__global__ void k_energy(PlanEnergyArgs a) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.pop) return;
    PlanIter it{a.ids, a.offsets[p], a.offsets[p + 1], a.N, a.hasStart, a.loops, 0};
    double* fit = a.fitness + 4 * p;
    int m = it.count();
    if (m == 0) {  // evaluateEmptyPlan, PlanCoverageEstimator.cpp:120-126
        fit[0] = 1.0; fit[1] = 0; fit[2] = 0; fit[3] = 0;
        a.active[p] = 0;
        return;
    }
    double energy = 0, pathLen = 0;
    int nColl = 0;
    int i0 = it.at(0);
    D3 p2 = d3(0, 0, 0), p1 = d3(a.pos[3 * i0], a.pos[3 * i0 + 1], a.pos[3 * i0 + 2]);
    int prevIdx = i0;
    for (int k = 1; k < m; ++k) {
        int ci = it.at(k);
        D3 c = d3(a.pos[3 * ci], a.pos[3 * ci + 1], a.pos[3 * ci + 2]);
        double e = 0.0;
        if (!equal(p1, c)) {
            e = move_energy(p1, c, k > 1 ? &p2 : nullptr);
            if (a.collisionPenalty > 0 && a.coll[prevIdx * a.N + ci] == 1) {
                e += a.collisionPenalty;
                nColl++;
            }
        }
        energy += e;
        pathLen += length(c - p1);
        p2 = p1;
        p1 = c;
        prevIdx = ci;
    }
    bool within = (a.maxEnergy == -1) ? true : (energy <= a.maxEnergy);  // PlanEnergyEvaluator.cpp:28-36
    bool act = a.disableLimit || within;
    fit[0] = 1.0;  // overwritten by plan_reduce for plans that pass the gate
    fit[1] = energy;
    fit[2] = pathLen;
    fit[3] = (double)nColl;
    a.active[p] = act ? 1 : 0;
}

__global__ void k_mark_rows(PlanMarkArgs a) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.pop) return;
    if (a.active && !a.active[p]) return;
    PlanIter it{a.ids, a.offsets[p], a.offsets[p + 1], a.N, a.hasStart, a.loops, 0};
    int m = it.count();
    int prev = m > 0 ? it.at(0) : 0;
    for (int k = 1; k < m; ++k) {
        int cur = it.at(k);
        int key = prev * a.N + cur;
        if (a.used) a.used[key] = 1;
        if (a.rowmap && a.rowmap[key] == -1) {
            if (atomicCAS(a.rowmap + key, -1, -2) == -1) {
                int slot = atomicAdd(a.newCount, 1);
                a.newKeys[slot] = key;
            }
        }
        prev = cur;
    }
}

__global__ void k_probe_rows(PlanMarkArgs a) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.pop) return;
    if (a.active && !a.active[p]) return;
    PlanIter it{a.ids, a.offsets[p], a.offsets[p + 1], a.N, a.hasStart, a.loops, 0};
    int m = it.count();
    int prev = m > 0 ? it.at(0) : 0;
    int unseen = 0;
    for (int k = 1; k < m; ++k) {
        int cur = it.at(k);
        if (a.rowmap[prev * a.N + cur] == -1) unseen++;
        prev = cur;
    }
    if (unseen) atomicAdd(a.newCount, unseen);
}

void launch_mark_collisions(const PlanMarkArgs& a, cudaStream_t st) {
    if (a.pop > 0) k_mark_collisions<<<(unsigned)((a.pop + 127) / 128), 128, 0, st>>>(a);
}
void launch_energy(const PlanEnergyArgs& a, cudaStream_t st) {
    if (a.pop > 0) k_energy<<<(unsigned)((a.pop + 127) / 128), 128, 0, st>>>(a);
}
void launch_probe_rows(const PlanMarkArgs& a, cudaStream_t st) {
    if (a.pop > 0) k_probe_rows<<<(unsigned)((a.pop + 127) / 128), 128, 0, st>>>(a);
}
void launch_mark_rows(const PlanMarkArgs& a, cudaStream_t st) {
    if (a.pop > 0) k_mark_rows<<<(unsigned)((a.pop + 127) / 128), 128, 0, st>>>(a);
}

// memo garbage collection (updateMemoisedEdges, PlanCoverageEstimator.cpp:284-327): rows whose key is
// not marked used go back to the free stack
__global__ void k_gc_rows(int* __restrict__ rowmap, const uint8_t* __restrict__ used, int nKeys, int* __restrict__ freeStack,
                          int* __restrict__ freeTop, int* __restrict__ live) {
    int key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nKeys) return;
    int r = rowmap[key];
    if (r == -2) rowmap[key] = -1;  // a claim that was never published (defensive: failed calls roll their claims back themselves)
    if (r < 0) return;
    if (used[key]) {
        atomicAdd(live, 1);
    } else {
        rowmap[key] = -1;
        int t = atomicAdd(freeTop, 1);
        freeStack[t] = r;
    }
}
void launch_gc_rows(int* d_rowmap, const uint8_t* d_used, int nKeys, int* d_freeStack, int* d_freeTop, int* d_live, cudaStream_t st) {
    if (nKeys > 0) k_gc_rows<<<(nKeys + 255) / 256, 256, 0, st>>>(d_rowmap, d_used, nKeys, d_freeStack, d_freeTop, d_live);
}

// rows for freshly claimed keys: pop from the free stack
__global__ void k_assign_rows(const int* __restrict__ keys, int n, int* __restrict__ rowmapPending, int* __restrict__ rowsOut,
                              const int* __restrict__ freeStack, int* __restrict__ freeTop, int scratchRow, int* __restrict__ errFlag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = atomicSub(freeTop, 1) - 1;
    if (t < 0) {  // the caller checks the capacity first; never index below the stack: render into the scratch row and report
        atomicAdd(freeTop, 1);
        *errFlag = 2;
        rowsOut[i] = scratchRow;
        return;
    }
    rowsOut[i] = freeStack[t];
}
void launch_assign_rows(const int* d_keys, int n, int* d_rowmap, int* d_rowsOut, const int* d_freeStack, int* d_freeTop, int scratchRow,
                        int* d_errFlag, cudaStream_t st) {
    if (n > 0) k_assign_rows<<<(n + 255) / 256, 256, 0, st>>>(d_keys, n, d_rowmap, d_rowsOut, d_freeStack, d_freeTop, scratchRow, d_errFlag);
}
// claims that will not be rendered after all (the call failed or is retried in smaller pieces): pending -> unknown
__global__ void k_unclaim_rows(const int* __restrict__ keys, int n, int* __restrict__ rowmap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && rowmap[keys[i]] == -2) rowmap[keys[i]] = -1;
}
void launch_unclaim_rows(const int* d_keys, int n, int* d_rowmap, cudaStream_t st) {
    if (n > 0) k_unclaim_rows<<<(n + 255) / 256, 256, 0, st>>>(d_keys, n, d_rowmap);
}
// memoization == false: the rows rendered for this call go straight back to the free stack
__global__ void k_unpublish_rows(const int* __restrict__ keys, int n, int* __restrict__ rowmap, int* __restrict__ freeStack,
                                 int* __restrict__ freeTop) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rowmap[keys[i]];
    rowmap[keys[i]] = -1;
    if (r >= 0) freeStack[atomicAdd(freeTop, 1)] = r;
}
void launch_unpublish_rows(const int* d_keys, int n, int* d_rowmap, int* d_freeStack, int* d_freeTop, cudaStream_t st) {
    if (n > 0) k_unpublish_rows<<<(n + 255) / 256, 256, 0, st>>>(d_keys, n, d_rowmap, d_freeStack, d_freeTop);
}
__global__ void k_publish_rows(const int* __restrict__ keys, const int* __restrict__ rows, int n, int* __restrict__ rowmap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rowmap[keys[i]] = rows[i];
}
void launch_publish_rows(const int* d_keys, const int* d_rows, int n, int* d_rowmap, cudaStream_t st) {
    if (n > 0) k_publish_rows<<<(n + 255) / 256, 256, 0, st>>>(d_keys, d_rows, n, d_rowmap);
}


__global__ void k_rows_gather(const int* __restrict__ keys, int n, const int* __restrict__ rowmap, const uint4* __restrict__ rows,
                              int q4, uint4* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n * q4) return;
    int k = (int)(i / q4), w = (int)(i - (long long)k * q4);
    out[i] = rows[(size_t)rowmap[keys[k]] * q4 + w];
}
__global__ void k_rows_scatter(const uint4* __restrict__ in, const int* __restrict__ rowOfSlot, int n, uint4* __restrict__ rows, int q4) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n * q4) return;
    int k = (int)(i / q4), w = (int)(i - (long long)k * q4);
    rows[(size_t)rowOfSlot[k] * q4 + w] = in[i];
}
void launch_rows_gather(const int* d_keys, int n, const int* d_rowmap, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_out, cudaStream_t st) {
    long long t = (long long)n * (rowWords / 4);
    if (t > 0) k_rows_gather<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(d_keys, n, d_rowmap, (const uint4*)d_rows, (int)(rowWords / 4), (uint4*)d_out);
}
void launch_rows_scatter(const uint32_t* d_in, const int* d_rowOfSlot, int n, uint32_t* d_rows, int64_t rowWords, cudaStream_t st) {
    long long t = (long long)n * (rowWords / 4);
    if (t > 0) k_rows_scatter<<<(unsigned)((t + 255) / 256), 256, 0, st>>>((const uint4*)d_in, d_rowOfSlot, n, (uint4*)d_rows, (int)(rowWords / 4));
}

}  // namespace ipp
