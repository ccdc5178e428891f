// visibility.cu -- K4 edge_visibility: the CUDA replacement of the reference's OpenGL item buffer
// (viewer.frame() + glReadPixels + getAllColorsInFrame; ImageViewerCaptureTool.cpp:59-70,
// CameraEstimator.cpp:99-129,131-178).
//
// Per camera pose: the scene box is culled against the camera frustum into a rectangle of 32x32-pixel
// macro tiles (exact.cu); one CTA takes a macro tile, each warp an 8x4-pixel packet; the 32 pixel-centre
// rays of a packet walk the BVH together with one warp-uniform stack in shared memory (ballot-voted
// child order, leaves of <= 4 triangles set up once by <= 4 lanes into shared memory and tested by all
// 32 rays).  The nearest triangle inside planar depth [zNear, zFar] (ties: lower id, = GL_LESS in
// draw order) gets its byte flag set with a plain store -- no atomics; k_pack_rows then builds the
// bitmap words with byte->bit multiplies and clears the flags.
#include "kernels.h"

namespace ipp {

namespace {

constexpr int kVisWarps = 8;
constexpr int kVisThreads = kVisWarps * 32;
constexpr int kChunk = 8;  // macro tiles per scheduling step
// Depth resolution of the item buffer (GL_LESS on a fixed-point depth buffer, triangles drawn in id order): fragments whose planar
// depth is within a relative 2^-16 of the nearest one are ties and the lowest id keeps the pixel. 
constexpr float kDepthTie = 1.0f + 1.0f / 65536.0f;

struct Hit {
    float t;
    int id;
};

// 4 float4 per triangle: (Na, vol) (Nb, id) (Nc, -) (Ng, -) with the eye at the origin.
//   e_a = D.Na, e_b = D.Nb, e_c = D.Nc are the homogeneous edge functions of the pixel ray D; the ray covers the triangle
//   iff they share a sign; planar depth t = vol / (D.Ng) with vol = (A - eye).Ng.
//   Every edge normal is evaluated from the lexicographically lower endpoint of its edge (the vertices are stored sorted), so two
//   triangles sharing an edge get bit-identical normals up to sign and no pixel ray can slip between them.
constexpr int kTriVec = 4;
__device__ __forceinline__ void tri_setup(const float4 a, const float4 b, const float4 c, const float4 n, float ox, float oy, float oz,
                                          float4& s0, float4& s1, float4& s2, float4& s3) {
    if (b.w != 0.f) {  // degenerate triangle: never produces a fragment
        s0 = make_float4(0.f, 0.f, 0.f, 0.f);
        s1 = make_float4(0.f, 0.f, 0.f, a.w);
        s2 = make_float4(0.f, 0.f, 0.f, 0.f);
        s3 = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    float ax = a.x - ox, ay = a.y - oy, az = a.z - oz;
    float bx = b.x - ox, by = b.y - oy, bz = b.z - oz;
    // edges from the original coordinates (well conditioned), not from the translated ones
    float ebcx = c.x - b.x, ebcy = c.y - b.y, ebcz = c.z - b.z;
    float eacx = c.x - a.x, eacy = c.y - a.y, eacz = c.z - a.z;
    float eabx = b.x - a.x, eaby = b.y - a.y, eabz = b.z - a.z;
    s0.x = by * ebcz - bz * ebcy; s0.y = bz * ebcx - bx * ebcz; s0.z = bx * ebcy - by * ebcx;           // B' x (C-B)
    s1.x = -(ay * eacz - az * eacy); s1.y = -(az * eacx - ax * eacz); s1.z = -(ax * eacy - ay * eacx);  // C' x (A-C) = -(A' x (C-A))
    s2.x = ay * eabz - az * eaby; s2.y = az * eabx - ax * eabz; s2.z = ax * eaby - ay * eabx;           // A' x (B-A)
    s0.w = ax * n.x + ay * n.y + az * n.z;  // vol = A'.Ng
    s1.w = a.w;                            // original triangle id (int bits)
    s2.w = 0.f;
    s3 = make_float4(n.x, n.y, n.z, 0.f);
}

__device__ __forceinline__ void tri_test(const float4 s0, const float4 s1, const float4 s2, const float4 s3, float dx, float dy, float dz,
                                         float zNear, float zFar, Hit& h) {
    float eu = dx * s0.x + dy * s0.y + dz * s0.z;
    float ev = dx * s1.x + dy * s1.y + dz * s1.z;
    float ew = dx * s2.x + dy * s2.y + dz * s2.z;
    bool in = (eu >= 0.f && ev >= 0.f && ew >= 0.f) || (eu <= 0.f && ev <= 0.f && ew <= 0.f);
    if (in) {
        float det = dx * s3.x + dy * s3.y + dz * s3.z;
        float t = s0.w / det;  // det == 0: t is inf or NaN and fails the range test
        int id = __float_as_int(s1.w);
        // depth-tie band (kDepthTie): fragments within a relative 2^-16 of the nearest are ties and the lowest id wins
        if (t >= zNear && t <= zFar && t <= h.t * kDepthTie) {
            h.id = (t * kDepthTie < h.t) ? id : min(h.id, id);  // clearly nearer: replace; inside the band: lowest id
            h.t = fminf(h.t, t);
        }
    }
}

// Ray/box slab test with the subtraction done first: (b - o) * inv keeps the relative error of every plane distance at
// ~1e-7 even when a direction component is close to zero (inv huge), where b * inv - o * inv cancels catastrophically.
__device__ __forceinline__ bool slab(float bx0, float by0, float bz0, float bx1, float by1, float bz1, float ix, float iy, float iz,
                                     float ox, float oy, float oz, float tmin, float tmax, float& tn) {
    float x0 = (bx0 - ox) * ix, x1 = (bx1 - ox) * ix;
    float y0 = (by0 - oy) * iy, y1 = (by1 - oy) * iy;
    float z0 = (bz0 - oz) * iz, z1 = (bz1 - oz) * iz;
    tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    return tn <= tf * 1.0000005f + 1e-7f;  // conservative against rounding (the boxes are padded as well)
}

__device__ __forceinline__ float safe_rcp(float d) {
    return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d));
}

// Warp-cooperative closest-hit traversal: all 32 rays share the origin and one stack.
__device__ __forceinline__ void trace_packet(const DeviceScene& sc, float ox, float oy, float oz, float dx, float dy, float dz, bool valid,
                                             float zNear, float zFar, int* stack, float4* s_tri, int lane, Hit& h) {
    h.t = valid ? zFar : -1.f;  // an invalid lane has an empty interval
    h.id = 0x7fffffff;
    const float tmin = zNear;
    float ix = safe_rcp(dx), iy = safe_rcp(dy), iz = safe_rcp(dz);
    int sp = 0;
    int node = 0;
    while (true) {
        const BvhNode* np = sc.nodes + node;
        float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
        float tn0, tn1;
        // a fragment can still matter up to the end of the depth-tie band behind the nearest hit so far
        const float tmax = h.t * kDepthTie;
        bool h0 = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ix, iy, iz, ox, oy, oz, tmin, tmax, tn0);
        bool h1 = slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ix, iy, iz, ox, oy, oz, tmin, tmax, tn1);
        unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
        int l0 = __float_as_int(q3.x), l1 = __float_as_int(q3.y);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            unsigned m = c ? m1 : m0;
            int l = c ? l1 : l0;
            if (m && l < 0) {
                int code = ~l;
                int first = code >> 3, cnt = code & 7;
                __syncwarp();
                if (lane < cnt) {
                    float4 a = __ldg(sc.ta + first + lane), b = __ldg(sc.tb + first + lane), cc = __ldg(sc.tc + first + lane),
                           nn = __ldg(sc.tn + first + lane);
                    float4 s0, s1, s2, s3;
                    tri_setup(a, b, cc, nn, ox, oy, oz, s0, s1, s2, s3);
                    s_tri[lane * kTriVec + 0] = s0;
                    s_tri[lane * kTriVec + 1] = s1;
                    s_tri[lane * kTriVec + 2] = s2;
                    s_tri[lane * kTriVec + 3] = s3;
                }
                __syncwarp();
                for (int k = 0; k < cnt; ++k)
                    tri_test(s_tri[k * kTriVec], s_tri[k * kTriVec + 1], s_tri[k * kTriVec + 2], s_tri[k * kTriVec + 3], dx, dy, dz, zNear,
                             zFar, h);
            }
        }
        bool i0 = m0 && l0 >= 0, i1 = m1 && l1 >= 0;
        if (i0 && i1) {
            // visit first the child that more rays enter first
            int v1 = __popc(__ballot_sync(0xffffffffu, h1 && (!h0 || tn1 < tn0)));
            int v0 = __popc(__ballot_sync(0xffffffffu, h0 && (!h1 || tn0 <= tn1)));
            int nearL = (v1 > v0) ? l1 : l0, farL = (v1 > v0) ? l0 : l1;
            if (lane == 0) stack[sp] = farL;
            sp++;
            node = nearL;
        } else if (i0) {
            node = l0;
        } else if (i1) {
            node = l1;
        } else {
            if (sp == 0) break;
            __syncwarp();
            node = stack[--sp];
        }
    }
}

__global__ void __launch_bounds__(kVisThreads) k_visibility(DeviceScene sc, CameraParams cam, const Snapshot* __restrict__ snaps,
                                                            const int* __restrict__ tileOffset, int nSnaps, long long totalTiles,
                                                            uint8_t* __restrict__ flags, int32_t* __restrict__ idImage,
                                                            unsigned long long* __restrict__ counters) {
    __shared__ int s_stack[kVisWarps][kMaxDepth];
    __shared__ float4 s_tri[kVisWarps][kLeafMax * kTriVec];
    __shared__ Snapshot s_snap;
    __shared__ long long s_chunk;
    __shared__ int s_snapIdx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long myRays = 0;
    while (true) {
        if (threadIdx.x == 0) s_chunk = (long long)atomicAdd(&counters[1], (unsigned long long)kChunk);
        __syncthreads();
        long long c0 = s_chunk;
        if (c0 >= totalTiles) break;
        long long c1 = c0 + kChunk < totalTiles ? c0 + kChunk : totalTiles;
        if (threadIdx.x == 0) {
            // last snapshot whose first tile is <= c0 (tileOffset has nSnaps + 1 entries)
            int lo = 0, hi = nSnaps - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if ((long long)tileOffset[mid] <= c0) lo = mid; else hi = mid - 1;
            }
            s_snapIdx = lo;
        }
        __syncthreads();
        int si = s_snapIdx;
        int loaded = -1;
        for (long long task = c0; task < c1; ++task) {
            while ((long long)tileOffset[si + 1] <= task) si++;  // skips culled (empty) snapshots
            if (si != loaded) {
                __syncthreads();
                const int* src = (const int*)(snaps + si);
                int* dst = (int*)&s_snap;
                for (int k = threadIdx.x; k < (int)(sizeof(Snapshot) / 4); k += kVisThreads) dst[k] = src[k];
                loaded = si;
                __syncthreads();
            }
            int local = (int)(task - tileOffset[si]);
            int mx = s_snap.rx0 + local % s_snap.rnx, my = s_snap.ry0 + local / s_snap.rnx;
            float ox = s_snap.eye[0], oy = s_snap.eye[1], oz = s_snap.eye[2];
            for (int sub = warp; sub < 32; sub += kVisWarps) {
                int px = mx * 32 + (sub & 3) * 8 + (lane & 7);
                int py = my * 32 + (sub >> 2) * 4 + (lane >> 3);
                bool valid = px < cam.W && py < cam.H;
                // pixel centre in image-plane units; the ray is scaled so that t is the planar depth (D.f = 1)
                double X = ((px + 0.5) / cam.W * 2.0 - 1.0) * cam.tanX;
                double Y = ((py + 0.5) / cam.H * 2.0 - 1.0) * cam.tanY;
                float dx = (float)(s_snap.f[0] + s_snap.s[0] * X + s_snap.u[0] * Y);
                float dy = (float)(s_snap.f[1] + s_snap.s[1] * X + s_snap.u[1] * Y);
                float dz = (float)(s_snap.f[2] + s_snap.s[2] * X + s_snap.u[2] * Y);
                Hit h;
                trace_packet(sc, ox, oy, oz, dx, dy, dz, valid, cam.zNear, cam.zFar, s_stack[warp], s_tri[warp], lane, h);
                bool hit = valid && h.id != 0x7fffffff;
                if (hit) flags[(size_t)s_snap.slot * sc.Tpad + h.id] = 1;
                if (idImage && valid) idImage[(size_t)py * cam.W + px] = hit ? h.id : -1;
                if (valid) myRays++;
            }
        }
        __syncthreads();
    }
    // rays issued: one atomic per warp
    for (int o = 16; o; o >>= 1) myRays += __shfl_down_sync(0xffffffffu, myRays, o);
    if (lane == 0 && myRays) atomicAdd(&counters[0], myRays);
}

// flags [slot][Tpad] bytes (0/1) -> bitmap row words; thread = one word = 32 flag bytes; clears the flags
__global__ void k_pack_rows(uint8_t* __restrict__ flags, int nSlots, int Tpad, const int* __restrict__ rowOfSlot,
                            uint32_t* __restrict__ rows, long long rowWords) {
    int wordsPerSlot = Tpad / 32;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nSlots * wordsPerSlot) return;
    int slot = (int)(i / wordsPerSlot), w = (int)(i - (long long)slot * wordsPerSlot);
    uint4* p = (uint4*)(flags + (size_t)slot * Tpad + (size_t)w * 32);
    uint4 a = p[0], b = p[1];
    const unsigned M = 0x01020408u;  // gathers the low bit of 4 bytes into bits 24..27
    unsigned word = ((a.x * M) >> 24 & 0xF) | (((a.y * M) >> 24 & 0xF) << 4) | (((a.z * M) >> 24 & 0xF) << 8) |
                    (((a.w * M) >> 24 & 0xF) << 12) | (((b.x * M) >> 24 & 0xF) << 16) | (((b.y * M) >> 24 & 0xF) << 20) |
                    (((b.z * M) >> 24 & 0xF) << 24) | (((b.w * M) >> 24 & 0xF) << 28);
    rows[(size_t)rowOfSlot[slot] * rowWords + w] = word;
    if (a.x | a.y | a.z | a.w) p[0] = make_uint4(0, 0, 0, 0);
    if (b.x | b.y | b.z | b.w) p[1] = make_uint4(0, 0, 0, 0);
}

// zero the padding words of freshly written rows (words between Tpad/32 and rowWords)
__global__ void k_zero_row_tail(int nSlots, int wordsPerSlot, const int* __restrict__ rowOfSlot, uint32_t* __restrict__ rows,
                                long long rowWords) {
    int tail = (int)(rowWords - wordsPerSlot);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tail <= 0 || i >= (long long)nSlots * tail) return;
    int slot = (int)(i / tail), w = (int)(i - (long long)slot * tail);
    rows[(size_t)rowOfSlot[slot] * rowWords + wordsPerSlot + w] = 0u;
}

// ---- BVH self test: every thread traces one ray through the BVH and by brute force ----
__device__ void trace_single(const DeviceScene& sc, float ox, float oy, float oz, float dx, float dy, float dz, Hit& h) {
    h.t = 1e30f;
    h.id = 0x7fffffff;
    float ix = safe_rcp(dx), iy = safe_rcp(dy), iz = safe_rcp(dz);
    int stack[kMaxDepth];
    int sp = 0, node = 0;
    while (true) {
        const BvhNode* np = sc.nodes + node;
        float4 q0 = np->q0, q1 = np->q1, q2 = np->q2, q3 = np->q3;
        float tn0, tn1;
        bool h0 = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ix, iy, iz, ox, oy, oz, 0.f, h.t * kDepthTie, tn0);
        bool h1 = slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ix, iy, iz, ox, oy, oz, 0.f, h.t * kDepthTie, tn1);
        int l[2] = {__float_as_int(q3.x), __float_as_int(q3.y)};
        bool hh[2] = {h0, h1};
        int next = -1;
        for (int c = 0; c < 2; ++c) {
            if (!hh[c]) continue;
            if (l[c] < 0) {
                int code = ~l[c];
                int first = code >> 3, cnt = code & 7;
                for (int k = 0; k < cnt; ++k) {
                    float4 s0, s1, s2, s3;
                    tri_setup(sc.ta[first + k], sc.tb[first + k], sc.tc[first + k], sc.tn[first + k], ox, oy, oz, s0, s1, s2, s3);
                    tri_test(s0, s1, s2, s3, dx, dy, dz, 0.f, 1e30f, h);
                }
            } else if (next < 0) {
                next = l[c];
            } else if (sp < kMaxDepth) {
                stack[sp++] = l[c];
            }
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
}

__global__ void k_bvh_selftest(DeviceScene sc, const float* __restrict__ rays, long long n, int32_t* __restrict__ idsBvh,
                               int32_t* __restrict__ idsBrute, float* __restrict__ tOut) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ox = rays[6 * i], oy = rays[6 * i + 1], oz = rays[6 * i + 2], dx = rays[6 * i + 3], dy = rays[6 * i + 4], dz = rays[6 * i + 5];
    Hit h;
    trace_single(sc, ox, oy, oz, dx, dy, dz, h);
    idsBvh[i] = h.id == 0x7fffffff ? -1 : h.id;
    Hit b;
    b.t = 1e30f;
    b.id = 0x7fffffff;
    for (int k = 0; k < sc.T; ++k) {
        float4 s0, s1, s2, s3;
        tri_setup(sc.ta[k], sc.tb[k], sc.tc[k], sc.tn[k], ox, oy, oz, s0, s1, s2, s3);
        tri_test(s0, s1, s2, s3, dx, dy, dz, 0.f, 1e30f, b);
    }
    // near-ties (two surfaces within rounding of each other along the ray) may be culled by the box test: same depth counts as agreement
    bool same = (b.id == h.id) || (b.id != 0x7fffffff && h.id != 0x7fffffff && fabsf(b.t - h.t) <= 1e-5f * fmaxf(1.f, fabsf(b.t)));
    idsBrute[i] = same ? idsBvh[i] : (b.id == 0x7fffffff ? -1 : b.id);
    if (tOut) {
        tOut[2 * i] = h.t;
        tOut[2 * i + 1] = b.t;
    }
}

}  // namespace

int visibility_grid_blocks() {
    int dev = 0, sms = 148, perSM = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_visibility, kVisThreads, 0);
    if (perSM < 1) perSM = 1;
    return sms * perSM;
}

void launch_visibility(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, const int* d_tileOffset, int nSnaps,
                       int64_t totalTiles, uint8_t* d_flags, int32_t* d_idImage, unsigned long long* d_counters, cudaStream_t st) {
    if (totalTiles <= 0 || nSnaps <= 0) return;
    static int blocks = 0;
    if (!blocks) blocks = visibility_grid_blocks();
    long long chunks = (totalTiles + kChunk - 1) / kChunk;
    int grid = (int)std::min<long long>(blocks, chunks);
    // counters[1] is the chunk cursor of this launch
    cudaMemsetAsync(d_counters + 1, 0, sizeof(unsigned long long), st);
    k_visibility<<<grid, kVisThreads, 0, st>>>(sc, cam, d_snaps, d_tileOffset, nSnaps, (long long)totalTiles, d_flags, d_idImage, d_counters);
}

void launch_pack_rows(uint8_t* d_flags, int nSlots, int Tpad, const int* d_rowOfSlot, uint32_t* d_rows, int64_t rowWords, cudaStream_t st) {
    if (nSlots <= 0) return;
    int wordsPerSlot = Tpad / 32;
    long long n = (long long)nSlots * wordsPerSlot;
    k_pack_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_flags, nSlots, Tpad, d_rowOfSlot, d_rows, (long long)rowWords);
    long long tail = (long long)nSlots * (rowWords - wordsPerSlot);
    if (tail > 0) k_zero_row_tail<<<(unsigned)((tail + 255) / 256), 256, 0, st>>>(nSlots, wordsPerSlot, d_rowOfSlot, d_rows, (long long)rowWords);
}

void launch_bvh_selftest(const DeviceScene& sc, const float* d_rays, int64_t n, int32_t* d_idsBvh, int32_t* d_idsBrute, float* d_t,
                         cudaStream_t st) {
    if (n > 0) k_bvh_selftest<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(sc, d_rays, (long long)n, d_idsBvh, d_idsBrute, d_t);
}

}  // namespace ipp
