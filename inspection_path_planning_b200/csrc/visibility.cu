// visibility.cu -- K4 edge_visibility: the CUDA replacement of the reference's OpenGL item buffer
// (viewer.frame() + glReadPixels + getAllColorsInFrame; ImageViewerCaptureTool.cpp:59-70,
// CameraEstimator.cpp:99-129,131-178).
//
// Tile-packet ray casting.  Per camera pose the scene box is culled against the camera frustum into a rectangle of
// 32x32-pixel tiles (exact.cu).  One warp takes one tile: its 1024 pixel-centre rays walk the BVH together as one frustum
// packet -- the warp tests both child boxes of a node against the four side planes of the tile frustum, the near plane and
// the tile's current farthest hit (12 plane tests spread over 12 lanes, one ballot), front child first, on a warp-uniform
// stack in shared memory.  At a leaf up to 8 lanes set up one triangle each into shared memory (homogeneous edge functions
// of the pixel ray D = f + X s + Y u: e(X, Y) = N.f + (N.s) X + (N.u) Y, depth t = (A - eye).Ng / D.Ng, projected bounding
// rectangle); the warp then runs the ray/triangle test for the rays inside that rectangle, 8x4 pixels per step, every lane
// owning the same 32 pixels of the tile throughout, and keeps the nearest hit per ray in shared memory: no atomics.
// The nearest triangle inside planar depth [zNear, zFar] gets its byte flag set with a plain store; k_pack_rows then builds
// the bitmap words with byte->bit multiplies and clears the flags.
//
// Numerics that matter for parity:
//  * every edge normal is evaluated from the lexicographically lower endpoint of its edge (vertices are stored sorted), so two
//    triangles sharing an edge get bit-identical coefficients up to sign and no pixel ray slips between them;
//  * Ng comes from an fp64 cross product (lbvh.cu), the fp32 one loses most digits on long thin triangles;
//  * depth resolution: fragments within a relative 2^-16 of the nearest are ties and the lowest id wins (GL_LESS on a
//    fixed-point depth buffer, triangles drawn in id order).
#include "raster.cuh"

namespace ipp {

namespace {

using namespace raster;

constexpr int kVisWarps = 4;
constexpr int kVisThreads = kVisWarps * 32;
constexpr int kGrab = 4;        // tiles a warp claims per scheduling step

struct Hit {
    float t;
    int id;
};

// a*b - c*d and a.b with a fixed rounding sequence.  Explicit intrinsics: the compiler may neither contract nor reassociate them, so
// the same inputs give the same bits everywhere and negated inputs give exactly negated results -- the property the shared-edge
// argument below rests on (-(x*y - z*w) silently rewritten as fma(z, w, -(x*y)) rounds the other product first and breaks it).
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) { return __fmaf_rn(a, b, -__fmul_rn(c, d)); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}

// Edge normals with the eye at the origin.  e_a = D.Na, e_b = D.Nb, e_c = D.Nc are the homogeneous edge functions of a ray D.
// Each normal is lowerEndpoint' x (upperEndpoint - lowerEndpoint) of its edge (vertices are stored lexicographically sorted), up to
// sign: two triangles sharing an edge compute bit-identical normals up to sign.
__device__ __forceinline__ void edge_normals(const float4 a, const float4 b, const float4 c, float ox, float oy, float oz, float3& na,
                                             float3& nb, float3& nc, float3& ap) {
    float ax = a.x - ox, ay = a.y - oy, az = a.z - oz;
    float bx = b.x - ox, by = b.y - oy, bz = b.z - oz;
    // edges from the original coordinates (well conditioned), not from the translated ones
    float ebcx = c.x - b.x, ebcy = c.y - b.y, ebcz = c.z - b.z;
    float eacx = c.x - a.x, eacy = c.y - a.y, eacz = c.z - a.z;
    float eabx = b.x - a.x, eaby = b.y - a.y, eabz = b.z - a.z;
    // B' x (C-B)
    na = make_float3(diff_of_products(by, ebcz, bz, ebcy), diff_of_products(bz, ebcx, bx, ebcz), diff_of_products(bx, ebcy, by, ebcx));
    // C' x (A-C) = -(A' x (C-A))
    nb = make_float3(-diff_of_products(ay, eacz, az, eacy), -diff_of_products(az, eacx, ax, eacz), -diff_of_products(ax, eacy, ay, eacx));
    // A' x (B-A)
    nc = make_float3(diff_of_products(ay, eabz, az, eaby), diff_of_products(az, eabx, ax, eabz), diff_of_products(ax, eaby, ay, eabx));
    ap = make_float3(ax, ay, az);
}

__device__ __forceinline__ void depth_update(float t, int id, float zNear, float zFar, Hit& h) {
    // depth-tie band: fragments within a relative 2^-16 of the nearest are ties and the lowest id wins
    if (t >= zNear && t <= zFar && t <= h.t * kDepthTie) {
        h.id = (t * kDepthTie < h.t) ? id : min(h.id, id);  // clearly nearer: replace; inside the band: lowest id
        h.t = fminf(h.t, t);
    }
}

// ray form of the test, used by the BVH self test
__device__ __forceinline__ void ray_tri_test(const float4 a, const float4 b, const float4 c, const float4 n, float ox, float oy, float oz,
                                             float dx, float dy, float dz, float zNear, float zFar, Hit& h) {
    if (b.w != 0.f) return;  // degenerate triangle: never produces a fragment
    float3 na, nb, nc, ap;
    edge_normals(a, b, c, ox, oy, oz, na, nb, nc, ap);
    float eu = dot3(na.x, na.y, na.z, dx, dy, dz);
    float ev = dot3(nb.x, nb.y, nb.z, dx, dy, dz);
    float ew = dot3(nc.x, nc.y, nc.z, dx, dy, dz);
    bool in = (eu >= 0.f && ev >= 0.f && ew >= 0.f) || (eu <= 0.f && ev <= 0.f && ew <= 0.f);
    if (in) {
        float det = dx * n.x + dy * n.y + dz * n.z;
        float vol = ap.x * n.x + ap.y * n.y + ap.z * n.z;
        depth_update(vol / det, __float_as_int(a.w), zNear, zFar, h);  // det == 0: inf / NaN fails the range test
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

__device__ __forceinline__ float safe_rcp(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

// ------------------------------------------------------------------------------------------------------------------
// K4
// ------------------------------------------------------------------------------------------------------------------
struct WarpShared {
    float t[kTilePixels];            // nearest depth per ray, index = block * 32 + lane (block = 8x4 pixels)
    int id[kTilePixels];             // its triangle
    float tri[kPending][kTriWords];  // staged triangles
    int stack[2 * kMaxDepth];        // (link, zmin bits) pairs
    Snapshot snap;                   // the camera pose of the current tile
};

__global__ void __launch_bounds__(kVisThreads, 5) k_visibility(DeviceScene sc, CameraParams cam, const Snapshot* __restrict__ snaps,
                                                            const int* __restrict__ tileOffset, int nSnaps, long long totalTiles,
                                                            uint8_t* __restrict__ flags, const int* __restrict__ slotList, int32_t* __restrict__ idImage,
                                                            unsigned long long* __restrict__ counters) {
    __shared__ WarpShared s_all[kVisWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpShared& S = s_all[warp];
    unsigned long long myRays = 0;
    unsigned myNodes = 0, myTris = 0, myTiles = 0;  // traversal statistics (lane 0 only): nodes fetched, triangles staged, tiles
    // pixel-centre image-plane coordinates: X(px) = ((px + 0.5) / W * 2 - 1) tanX = px kx + cx
    const float kx = (float)(2.0 * cam.tanX / cam.W), cx = (float)((1.0 / cam.W - 1.0) * cam.tanX);
    const float ky = (float)(2.0 * cam.tanY / cam.H), cy = (float)((1.0 / cam.H - 1.0) * cam.tanY);
    const float zNear = cam.zNear, zFar = cam.zFar;
    // this lane's role in the node test: child = lane / 6, quantity = lane % 6 (lanes >= 12 idle)
    const int role = lane < 12 ? lane % 6 : 0;
    const bool second = lane >= 6 && lane < 12;

    while (true) {
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(&counters[1], (unsigned long long)kGrab);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= totalTiles) break;
        const long long end = base + kGrab < totalTiles ? base + kGrab : totalTiles;
        int si = find_snapshot(tileOffset, nSnaps, base, lane);
        int loaded = -1;
        float ox = 0, oy = 0, oz = 0;
        for (long long task = base; task < end; ++task) {
            while ((long long)__ldg(tileOffset + si + 1) <= task) si++;  // skips culled (empty) snapshots
            if (si != loaded) {
                __syncwarp();
                const int* src = (const int*)(snaps + si);
                int* dst = (int*)&S.snap;
                if (lane < (int)(sizeof(Snapshot) / 4)) dst[lane] = __ldg(src + lane);
                __syncwarp();
                ox = S.snap.eye[0]; oy = S.snap.eye[1]; oz = S.snap.eye[2];
                loaded = si;
            }
            const float fx = (float)S.snap.f[0], fy = (float)S.snap.f[1], fz = (float)S.snap.f[2];
            const float sx = (float)S.snap.s[0], sy = (float)S.snap.s[1], sz = (float)S.snap.s[2];
            const float ux = (float)S.snap.u[0], uy = (float)S.snap.u[1], uz = (float)S.snap.u[2];
            const int rx0 = S.snap.rx0, ry0 = S.snap.ry0, rnx = S.snap.rnx;
            const int local = (int)(task - __ldg(tileOffset + si));
            const int tpx0 = (rx0 + local % rnx) * kTile, tpy0 = (ry0 + local / rnx) * kTile;
            // ---- tile state ----
            const int vw = min(kTile, cam.W - tpx0), vh = min(kTile, cam.H - tpy0);  // valid pixels of a border tile
            tile_init(S.t, S.id, lane, vw, vh, zFar);
            float zcull = zFar * kDepthTie;  // nothing beyond this depth can matter to any ray of the tile
            // lane b keeps the depth beyond which nothing can matter to any ray of 8x4 block b (0: block outside the image)
            float bzReg = ((lane & 3) * 8 < vw && (lane >> 2) * 4 < vh) ? zcull : 0.f;
            // ---- tile frustum: the rays through the pixel centres, widened by half a pixel ----
            const float XL = (tpx0 - 0.5f) * kx + cx, XR = (tpx0 + vw - 0.5f) * kx + cx;
            const float YB = (tpy0 - 0.5f) * ky + cy, YT = (tpy0 + vh - 0.5f) * ky + cy;
            float nx, ny, nz, thr;
            switch (role) {
                case 0: nx = sx - XL * fx; ny = sy - XL * fy; nz = sz - XL * fz; thr = -1e-4f; break;  // x >= XL z
                case 1: nx = XR * fx - sx; ny = XR * fy - sy; nz = XR * fz - sz; thr = -1e-4f; break;  // x <= XR z
                case 2: nx = ux - YB * fx; ny = uy - YB * fy; nz = uz - YB * fz; thr = -1e-4f; break;  // y >= YB z
                case 3: nx = YT * fx - ux; ny = YT * fy - uy; nz = YT * fz - uz; thr = -1e-4f; break;  // y <= YT z
                case 4: nx = fx; ny = fy; nz = fz; thr = zNear * 0.9999f - 1e-4f; break;                 // zmax >= zNear
                default: nx = -fx; ny = -fy; nz = -fz; thr = 0.f; break;                                 // zmin <= zcull (limit set per node)
            }
            TileGeom tg;
            tile_geometry(tg, tpx0, tpy0, lane, kx, cx, ky, cy);
            __syncwarp();

            // ---- frustum-packet traversal ----
            // Leaf triangles are not tested one leaf at a time: they queue up (lane l holds the l-th pending triangle) until the
            // queue could overflow at the next node, then all pending triangles are set up in parallel, one per lane (the fp64
            // setup then runs on up to 32 lanes instead of <= 8), and tested one after the other.
            int sp_ = 0;
            int node = 0;
            bool more = sc.T > 0;
            int pend = 0, myTri = 0;
            bool tileHit = false;  // some ray of the tile has a hit
            for (;;) {
                if (pend > kPending - 2 * kLeafMax || (!more && pend > 0)) {
                    // ---- flush the pending triangles ----
                    __syncwarp();
                    bool mine = false;
                    if (lane < pend) mine = setup_triangle(sc, &S.snap, myTri, S.tri[lane], ox, oy, oz, zNear, zcull, tg);
                    const unsigned live = __ballot_sync(0xffffffffu, mine);  // triangles that can produce a fragment in this tile
                    __syncwarp();
                    const bool dirty = raster_staged(live, S.tri, S.t, S.id, tg, lane, zNear, zFar, zcull, bzReg);
                    pend = 0;
                    if (__any_sync(0xffffffffu, dirty)) {
                        tileHit = true;
                        __syncwarp();
                        refresh_depth_bounds(S.t, lane, bzReg, zcull);
                    }
                }
                if (!more) break;
                const BvhNode* np = sc.nodes + node;
                const float4 q0 = __ldg(&np->q0), q1 = __ldg(&np->q1), q2 = __ldg(&np->q2), q3 = __ldg(&np->q3);
                myNodes++;
                // this lane's child box, relative to the eye
                const float bx0 = (second ? q1.z : q0.x) - ox, by0 = (second ? q1.w : q0.y) - oy, bz0 = (second ? q2.x : q0.z) - oz;
                const float bx1 = (second ? q2.y : q0.w) - ox, by1 = (second ? q2.z : q1.x) - oy, bz1 = (second ? q2.w : q1.y) - oz;
                // maximum of n.P over the box
                const float m = (nx >= 0.f ? bx1 : bx0) * nx + (ny >= 0.f ? by1 : by0) * ny + (nz >= 0.f ? bz1 : bz0) * nz;
                const float limit = role == 5 ? -(zcull * 1.000001f + 1e-4f) : thr;
                const unsigned ok = __ballot_sync(0xffffffffu, m >= limit);
                const bool h0 = (ok & 0x3fu) == 0x3fu, h1 = (ok & 0xfc0u) == 0xfc0u;
                const float zmin0 = -__shfl_sync(0xffffffffu, m, 5), zmin1 = -__shfl_sync(0xffffffffu, m, 11);
                const int l0 = __float_as_int(q3.x), l1 = __float_as_int(q3.y);
                // ---- leaves: queue their triangles ----
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int l = c ? l1 : l0;
                    if (!(c ? h1 : h0) || l >= 0) continue;
                    const int code = ~l;
                    const int first = code >> kLeafShift, cnt = code & kLeafMask;
                    if (lane >= pend && lane < pend + cnt) myTri = first + lane - pend;
                    pend += cnt;
                    myTris += cnt;
                }
                // ---- inner children: nearer one first ----
                const bool i0 = h0 && l0 >= 0, i1 = h1 && l1 >= 0;
                if (i0 && i1) {
                    const bool swap = zmin1 < zmin0;
                    if (lane == 0) {
                        S.stack[2 * sp_] = swap ? l0 : l1;
                        S.stack[2 * sp_ + 1] = __float_as_int(swap ? zmin0 : zmin1);
                        // the far child is visited later: start pulling its node towards L1 now
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.nodes + (swap ? l0 : l1)));
                    }
                    sp_++;
                    node = swap ? l1 : l0;
                } else if (i0) {
                    node = l0;
                } else if (i1) {
                    node = l1;
                } else {
                    more = false;
                    __syncwarp();
                    while (sp_ > 0) {
                        --sp_;
                        const int cand = S.stack[2 * sp_];
                        const float zmin = __int_as_float(S.stack[2 * sp_ + 1]);
                        if (zmin <= zcull * 1.000001f + 1e-4f) {  // still in front of the farthest hit
                            node = cand;
                            more = true;
                            break;
                        }
                    }
                }
            }
            __syncwarp();
            // ---- the winners ----
            if (tileHit || idImage) tile_output(S.id, lane, flags, (size_t)sc.Tpad, slotList + S.snap.slot, S.snap.slotCount, idImage, tpx0, tpy0, cam.W, cam.H);
            if (lane == 0) myRays += (unsigned long long)(vw * vh);
            myTiles++;
            __syncwarp();
        }
    }
    if (lane == 0 && myRays) {
        atomicAdd(&counters[0], myRays);
        atomicAdd(&counters[2], (unsigned long long)myNodes);
        atomicAdd(&counters[3], (unsigned long long)myTris);
        atomicAdd(&counters[4], (unsigned long long)myTiles);
    }
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
    h.id = kNoId;
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
                int first = code >> kLeafShift, cnt = code & kLeafMask;
                for (int k = 0; k < cnt; ++k)
                    ray_tri_test(sc.ta[first + k], sc.tb[first + k], sc.tc[first + k], sc.tn[first + k], ox, oy, oz, dx, dy, dz, 0.f, 1e30f, h);
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
    idsBvh[i] = h.id == kNoId ? -1 : h.id;
    Hit b;
    b.t = 1e30f;
    b.id = kNoId;
    for (int k = 0; k < sc.T; ++k) ray_tri_test(sc.ta[k], sc.tb[k], sc.tc[k], sc.tn[k], ox, oy, oz, dx, dy, dz, 0.f, 1e30f, b);
    // near-ties (two surfaces within rounding of each other along the ray) may be culled by the box test: same depth counts as agreement
    bool same = (b.id == h.id) || (b.id != kNoId && h.id != kNoId && fabsf(b.t - h.t) <= 1e-5f * fmaxf(1.f, fabsf(b.t)));
    idsBrute[i] = same ? idsBvh[i] : (b.id == kNoId ? -1 : b.id);
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
    cudaFuncSetAttribute(k_visibility, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_visibility, kVisThreads, 0);
    if (perSM < 1) perSM = 1;
    return sms * perSM;
}

void launch_visibility(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, const int* d_tileOffset, int nSnaps,
                       int64_t totalTiles, uint8_t* d_flags, const int* d_slotList, int32_t* d_idImage, unsigned long long* d_counters,
                       cudaStream_t st) {
    if (totalTiles <= 0 || nSnaps <= 0) return;
    static int blocks = 0;
    if (!blocks) blocks = visibility_grid_blocks();
    long long groups = (totalTiles + (long long)kVisWarps * kGrab - 1) / ((long long)kVisWarps * kGrab);
    int grid = (int)std::min<long long>(blocks, groups);
    // counters[1] is the tile cursor of this launch
    cudaMemsetAsync(d_counters + 1, 0, sizeof(unsigned long long), st);
    k_visibility<<<grid, kVisThreads, 0, st>>>(sc, cam, d_snaps, d_tileOffset, nSnaps, (long long)totalTiles, d_flags, d_slotList, d_idImage, d_counters);
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
