// visibility_bins.cu -- K4b: edge visibility with depth-sorted leaf bins instead of a BVH walk per tile.  Same item-buffer sampling
// and the same per-tile rasteriser as visibility.cu (raster.cuh); what changes is how a tile finds its triangles.
//
//   k_leaf_table  one CTA per rendered camera pose: every cluster box (a BVH leaf, or a subtree of <= kMaxClusterTris triangles for
//                 meshes with more than kMaxBinLeaves leaves; a few thousand for the reference's meshes) is projected -- 8 corners,
//                 screen rectangle in 32x32-pixel tile units with a one-pixel margin, nearest depth -- culled against the view
//                 volume, packed into one 64-bit word [depth:16 | tx0 tx1 ty0 ty1:20 | cluster:28]; the survivors are compacted and
//                 block-radix sorted by depth (cub::BlockRadixSort, stable, so the order is deterministic).  The CTA also ORs the
//                 rectangles into a 32x32 tile occupancy mask and an 8x8 supertile mask: empty tiles are never scheduled.
//   k_visibility_bins  persistent grid, one warp per occupied 128x128 supertile (or per quadrant / tile of one when a launch has
//                 too few supertiles to fill the grid): one coalesced pass over the pose's table keeps the entries that touch the
//                 supertile (per-warp list in a scratch).  The tile-independent half of the triangle set-up then runs once per
//                 triangle of that list: triangles that can produce a fragment in the supertile become 64-byte records (a second
//                 per-warp scratch) with a 4-byte header each [depth key of the cluster:16 | tiles of the supertile the triangle
//                 can touch:16] -- the exact tile-level edge test, not the screen box.  The warp then rasterises its <= 16 tiles
//                 one after the other: a tile scans the headers front to back, 32 per load, stages the records that carry its bit
//                 as soon as a chunk has produced any (so that the depth bounds stay fresh), and stops at the first chunk behind
//                 its farthest hit -- strict front-to-back order makes the hierarchical depth culling of raster.cuh bite much
//                 earlier than the approximate order of a BVH walk, no dependent node fetch is left on the critical path, and a
//                 tile never looks at a triangle that cannot touch it.  A supertile with more surviving triangles than the record
//                 scratch holds is split into halves (down to single tiles, which take their triangles in several batches).
//
// Used when the mesh has <= kMaxBinLeaves clusters of <= kMaxClusterTris triangles (meshes up to ~2 M triangles) and the image is
// <= 1024 pixels a side; other inputs take the BVH-walking kernel of visibility.cu.
#include <cstdlib>

#include <cub/block/block_radix_sort.cuh>

#include "raster.cuh"

namespace ipp {

namespace {

using namespace raster;

constexpr int kBinWarps = 4;
constexpr int kBinThreads = kBinWarps * 32;
constexpr unsigned long long kNoEntry = ~0ull;
constexpr int kSortRadixBits = 6;  // 12 depth bits = 2 passes
constexpr int kSmallItems = 2;     // keys per thread of the short sort a pose with few surviving leaves takes
constexpr int kBinRecords = 4096;   // set-up triangles (64 B each) a warp can keep for one supertile (or part of one)
#ifndef IPP_STAGE_MIN
#define IPP_STAGE_MIN 1
#endif
constexpr int kStageMin = IPP_STAGE_MIN;  // waiting records that make a staging round worth running at the end of a header chunk
#ifndef IPP_FILTER_UNROLL
#define IPP_FILTER_UNROLL 4
#endif
constexpr int kFilterUnroll = IPP_FILTER_UNROLL;
#ifndef IPP_SUB_ROUND
#define IPP_SUB_ROUND 8
#endif
constexpr int kSubRound = IPP_SUB_ROUND;  // staged triangles rasterised between two refreshes of the tile's depth bounds    // table loads a lane keeps in flight while a supertile's list is filtered
static_assert(kBinRecords <= 65536, "tile-list entries carry 16 bits of record index");
static_assert(kBinRecords >= 64, "a batch of set-up records ends at a round of 32 triangles and resumes there");

struct __align__(16) BinShared {
    float t[kTilePixels];
    int id[kTilePixels];
    float tri[kPending][kTriWords];
    uint16_t queue[96];  // records of the tile being rasterised that wait for a staging round
    float tileX[8], tileY[8];  // extreme pixel centres of the supertile's tile columns and rows ([k] min, [4 + k] max)
    Snapshot snap;
};

// ---------------------------------------------------------------------------------------------------------------------
// leaf table
// ---------------------------------------------------------------------------------------------------------------------
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 512 ? 2 : 1)) k_leaf_table(DeviceScene sc, CameraParams cam, const Snapshot* __restrict__ snaps, int stride,
                                                        unsigned long long* __restrict__ table, int* __restrict__ tableCount,
                                                        uint32_t* __restrict__ tileMask, unsigned long long* __restrict__ superMask,
                                                        int* __restrict__ nSuper) {
    using Sort = cub::BlockRadixSort<unsigned long long, THREADS, ITEMS, cub::NullType, kSortRadixBits>;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(smemRaw);
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(smemRaw);  // keys before the sort (same bytes as temp)
    __shared__ unsigned rowMask[32];
    __shared__ int nValid;
    __shared__ int warpTotal[32];
    __shared__ unsigned superHalf[2];

    const int si = blockIdx.x, tid = threadIdx.x;
    const Snapshot& sp = snaps[si];
    if (sp.rnx == 0 || sp.rny == 0) {  // the scene box is outside the view volume (make_snapshot)
        if (tid < 32) tileMask[(size_t)si * 32 + tid] = 0u;
        if (tid == 0) { tableCount[si] = 0; superMask[si] = 0ull; nSuper[si] = 0; }
        return;
    }
    if (tid < 32) rowMask[tid] = 0u;
    if (tid == 0) nValid = 0;
    __syncthreads();

    const float ox = sp.eye[0], oy = sp.eye[1], oz = sp.eye[2];
    const float fx = (float)sp.f[0], fy = (float)sp.f[1], fz = (float)sp.f[2];
    const float sx = (float)sp.s[0], sy = (float)sp.s[1], sz = (float)sp.s[2];
    const float ux = (float)sp.u[0], uy = (float)sp.u[1], uz = (float)sp.u[2];
    const float kx = (float)(2.0 * cam.tanX / cam.W), cx = (float)((1.0 / cam.W - 1.0) * cam.tanX);
    const float ky = (float)(2.0 * cam.tanY / cam.H), cy = (float)((1.0 / cam.H - 1.0) * cam.tanY);
    const float ikx = 1.f / kx, iky = 1.f / ky;
    const float zFarCull = cam.zFar * kDepthTie;
    const float qscale = 65535.f / (zFarCull * 1.000001f + 1e-4f);
    // fragments exist only at depth >= zNear: boxes are clipped at a plane just inside it before they are projected (a near plane
    // closer than 1 cm leaves no headroom for the fp32 projection: such boxes take the whole image when they straddle it)
    const float zClip = cam.zNear * 0.999f, izClip = 1.f / zClip;
    const bool clipOk = zClip >= 0.01f;
    const int txMax = (cam.W - 1) >> 5, tyMax = (cam.H - 1) >> 5;

    int mine = 0;
    // one leaf per thread and step, not unrolled: the body is long and ITEMS copies of it would not fit the instruction cache
#pragma unroll 1
    for (int leaf = tid; leaf < THREADS * ITEMS; leaf += THREADS) {
        stage[leaf] = kNoEntry;
        if (leaf >= sc.nLeaves) continue;
        const float4 lo4 = __ldg(sc.leafLo + leaf), hi4 = __ldg(sc.leafHi + leaf);
        const float lx = lo4.x - ox, ly = lo4.y - oy, lz = lo4.z - oz, hx = hi4.x - ox, hy = hi4.y - oy, hz = hi4.z - oz;
        // camera-space coordinates of the corners are sums of one term per axis
        const float zx0 = fx * lx, zx1 = fx * hx, zy0 = fy * ly, zy1 = fy * hy, zz0 = fz * lz, zz1 = fz * hz;
        const float zmin = fminf(zx0, zx1) + fminf(zy0, zy1) + fminf(zz0, zz1);
        const float zmax = fmaxf(zx0, zx1) + fmaxf(zy0, zy1) + fmaxf(zz0, zz1);
        if (zmax < cam.zNear * 0.9999f - 1e-4f || zmin > zFarCull * 1.000001f + 1e-4f) continue;
        int px0 = 0, px1 = cam.W - 1, py0 = 0, py1 = cam.H - 1;
        if (zmin >= zClip || clipOk) {
            // screen rectangle of the part of the box in front of the plane z = zClip (just inside the near plane): its corners in
            // front of the plane and the points where its edges cross the plane
            const float xx0 = sx * lx, xx1 = sx * hx, xy0 = sy * ly, xy1 = sy * hy, xz0 = sz * lz, xz1 = sz * hz;
            const float yx0 = ux * lx, yx1 = ux * hx, yy0 = uy * ly, yy1 = uy * hy, yz0 = uz * lz, yz1 = uz * hz;
            float Xmin = 1e30f, Xmax = -1e30f, Ymin = 1e30f, Ymax = -1e30f;
            float cxs[8], cys[8], czs[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                czs[c] = ((c & 1) ? zx1 : zx0) + ((c & 2) ? zy1 : zy0) + ((c & 4) ? zz1 : zz0);
                cxs[c] = ((c & 1) ? xx1 : xx0) + ((c & 2) ? xy1 : xy0) + ((c & 4) ? xz1 : xz0);
                cys[c] = ((c & 1) ? yx1 : yx0) + ((c & 2) ? yy1 : yy0) + ((c & 4) ? yz1 : yz0);
                if (czs[c] >= zClip) {
                    const float iz = rcp_approx(czs[c]);  // czs >= zClip >= 0.01; the rectangle is widened by a pixel below
                    const float X = cxs[c] * iz, Y = cys[c] * iz;
                    Xmin = fminf(Xmin, X); Xmax = fmaxf(Xmax, X);
                    Ymin = fminf(Ymin, Y); Ymax = fmaxf(Ymax, Y);
                }
            }
            if (zmin < zClip) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int ax = 1; ax <= 4; ax <<= 1) {
                        if (c & ax) continue;  // every edge once, from its lower corner
                        const int d = c | ax;
                        if ((czs[c] >= zClip) != (czs[d] >= zClip)) {
                            const float t = (zClip - czs[c]) / (czs[d] - czs[c]);
                            const float X = (cxs[c] + t * (cxs[d] - cxs[c])) * izClip, Y = (cys[c] + t * (cys[d] - cys[c])) * izClip;
                            Xmin = fminf(Xmin, X); Xmax = fmaxf(Xmax, X);
                            Ymin = fminf(Ymin, Y); Ymax = fmaxf(Ymax, Y);
                        }
                    }
                }
            }
            // pixel centres X(px) = px kx + cx inside [Xmin, Xmax], widened by one pixel against rounding
            const float a0 = fminf(fmaxf((Xmin - cx) * ikx, -1e6f), 1e6f), a1 = fminf(fmaxf((Xmax - cx) * ikx, -1e6f), 1e6f);
            const float b0 = fminf(fmaxf((Ymin - cy) * iky, -1e6f), 1e6f), b1 = fminf(fmaxf((Ymax - cy) * iky, -1e6f), 1e6f);
            px0 = (int)floorf(a0) - 1; px1 = (int)ceilf(a1) + 1;
            py0 = (int)floorf(b0) - 1; py1 = (int)ceilf(b1) + 1;
            if (px1 < 0 || py1 < 0 || px0 > cam.W - 1 || py0 > cam.H - 1) continue;
            px0 = max(px0, 0); py0 = max(py0, 0);
            px1 = min(px1, cam.W - 1); py1 = min(py1, cam.H - 1);
        }
        const unsigned tx0 = px0 >> 5, tx1 = min(px1 >> 5, txMax), ty0 = py0 >> 5, ty1 = min(py1 >> 5, tyMax);
        // depth in 2^-12 steps of the culling range (the low 4 of the 16 key bits stay 0: fewer sort passes)
        // (capped one step below the top: the 12 sorted bits of a real entry never equal those of the empty key)
        const unsigned zq = min((unsigned)fminf(65534.f, floorf(fmaxf(zmin, 0.f) * qscale)) & 0xfff0u, 0xffe0u);
        const unsigned ref = (unsigned)leaf;  // cluster index (first triangle and count: leafLo.w, leafHi.w)
        stage[leaf] = ((unsigned long long)zq << 48) | ((unsigned long long)tx0 << 43) | ((unsigned long long)tx1 << 38) |
                   ((unsigned long long)ty0 << 33) | ((unsigned long long)ty1 << 28) | ref;
        const unsigned bits = ((2u << tx1) - 1u) & ~((1u << tx0) - 1u);
        for (unsigned ty = ty0; ty <= ty1; ++ty) atomicOr(&rowMask[ty], bits);
        mine++;
    }
    if (mine) atomicAdd(&nValid, mine);
    __syncthreads();
    if (nValid == 0) {  // nothing of the scene in this view: no sort
        if (tid < 32) tileMask[(size_t)si * 32 + tid] = 0u;
        if (tid == 0) { tableCount[si] = 0; superMask[si] = 0ull; nSuper[si] = 0; }
        return;
    }
    unsigned long long keys[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) keys[it] = stage[it * THREADS + tid];
    __syncthreads();
    unsigned long long* out = table + (size_t)si * stride;
    const int n = nValid;
    if (ITEMS > kSmallItems && n <= THREADS * kSmallItems) {
        // A pose keeps a fraction of the leaves (a tenth on the bench workload): the surviving keys are compacted, in the order
        // the full sort takes its input in (thread, item), and a sort of kSmallItems keys per thread finishes the job -- same
        // keys, same relative order of equal depths, same table.
        int cnt = 0;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) cnt += keys[it] != kNoEntry ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((tid & 31) >= d) incl += v;
        }
        if ((tid & 31) == 31) warpTotal[tid >> 5] = incl;
        __syncthreads();
        int base = incl - cnt;
        for (int w = 0; w < (tid >> 5); ++w) base += warpTotal[w];
        unsigned long long* compact = stage;  // the keys are in registers: the staging area is free
#pragma unroll
        for (int it = 0; it < ITEMS; ++it)
            if (keys[it] != kNoEntry) compact[base++] = keys[it];
        __syncthreads();
        using SortS = cub::BlockRadixSort<unsigned long long, THREADS, kSmallItems, cub::NullType, kSortRadixBits>;
        unsigned long long ks[kSmallItems];
#pragma unroll
        for (int j = 0; j < kSmallItems; ++j) ks[j] = tid * kSmallItems + j < n ? compact[tid * kSmallItems + j] : kNoEntry;
        __syncthreads();
        SortS(*reinterpret_cast<typename SortS::TempStorage*>(smemRaw)).Sort(ks, 52, 64);
#pragma unroll
        for (int j = 0; j < kSmallItems; ++j) {
            const int pos = tid * kSmallItems + j;
            if (pos < n) out[pos] = ks[j];
        }
    } else {
        Sort(temp).Sort(keys, 52, 64);  // stable: entries of equal depth keep their (fixed) input order
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int pos = tid * ITEMS + it;
            if (pos < n) out[pos] = keys[it];
        }
    }
    __syncthreads();
    if (tid < 64) {
        const int ssx = tid & 7, ssy = tid >> 3;
        const unsigned any = (rowMask[4 * ssy] | rowMask[4 * ssy + 1] | rowMask[4 * ssy + 2] | rowMask[4 * ssy + 3]) & (0xFu << (4 * ssx));
        const unsigned b = __ballot_sync(0xffffffffu, any != 0u);
        if ((tid & 31) == 0) superHalf[tid >> 5] = b;
    }
    if (tid < 32) tileMask[(size_t)si * 32 + tid] = rowMask[tid];
    __syncthreads();
    if (tid == 0) {
        const unsigned long long m = (unsigned long long)superHalf[0] | ((unsigned long long)superHalf[1] << 32);
        tableCount[si] = n;
        superMask[si] = m;
        nSuper[si] = __popcll(m);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K4b
// ---------------------------------------------------------------------------------------------------------------------
// A region of a supertile: tile columns x0..x1 and rows y0..y1 (0..3), packed into 8 bits.
__device__ __forceinline__ unsigned region_pack(unsigned x0, unsigned x1, unsigned y0, unsigned y1) { return x0 | (x1 << 2) | (y0 << 4) | (y1 << 6); }
__device__ __forceinline__ unsigned region_mask(unsigned r) {  // bit ty * 4 + tx of every tile of the region
    const unsigned x0 = r & 3u, x1 = (r >> 2) & 3u, y0 = (r >> 4) & 3u, y1 = (r >> 6) & 3u;
    return (((2u << x1) - (1u << x0)) * 0x1111u) & ((16u << (4 * y1)) - 1u) & ~((1u << (4 * y0)) - 1u);
}

__global__ void __launch_bounds__(kBinThreads, 5) k_visibility_bins(DeviceScene sc, CameraParams cam, const Snapshot* __restrict__ snaps,
                                                                 const int* __restrict__ superOffset, int nSnaps,
                                                                 const unsigned long long* __restrict__ table, int stride,
                                                                 const int* __restrict__ tableCount, const uint32_t* __restrict__ tileMask,
                                                                 const unsigned long long* __restrict__ superMask,
                                                                 unsigned long long* scratch, float4* recordsAll, uint8_t* __restrict__ flags,
                                                                 const int* __restrict__ slotList, int32_t* __restrict__ idImage, unsigned long long* __restrict__ counters,
                                                                 int splitLog2) {
    __shared__ BinShared s_all[kBinWarps];
    const int warp = threadIdx.x >> 5;
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // volatile: kept in a register instead of re-read (S2R) inside the loops
    BinShared& S = s_all[warp];
    unsigned long long* list = scratch + (size_t)(blockIdx.x * kBinWarps + warp) * stride;
    float4* rec = recordsAll + (size_t)(blockIdx.x * kBinWarps + warp) * (kBinRecords * 17 / 4);
    unsigned* hdr = reinterpret_cast<unsigned*>(rec + (size_t)kBinRecords * 4);  // one word per record: cluster depth key:16 | tiles:16
    uint16_t* q16 = S.queue;
    unsigned long long myRays = 0, myEntries = 0;
    unsigned myTris = 0, myTiles = 0, mySuperRec = 0, mySuperSplit = 0;  // statistics (lane 0 only matters)
#ifdef IPP_VIS_DIAG
    unsigned long long dLive = 0, dOpen = 0, dBlocks = 0, dStaged = 0;  // instrumented builds only (tools/build_variant.py -DIPP_VIS_DIAG)
#endif
    const float kx = (float)(2.0 * cam.tanX / cam.W), cx = (float)((1.0 / cam.W - 1.0) * cam.tanX);
    const float ky = (float)(2.0 * cam.tanY / cam.H), cy = (float)((1.0 / cam.H - 1.0) * cam.tanY);
    const float zNear = cam.zNear, zFar = cam.zFar;
    const float iq = (zFar * kDepthTie * 1.000001f + 1e-4f) / 65535.f;  // depth of a quantised table entry (a lower bound)
    // a task = one occupied supertile, or -- when a launch has too few of them to fill the grid (the small batches of an NSGA-II
    // generation) -- one quadrant (splitLog2 = 2) or one tile (4) of it: the same work in finer pieces
    const long long totalSuper = (long long)__ldg(superOffset + nSnaps) << splitLog2;
    const unsigned ltMask = (1u << lane) - 1u;

    while (true) {
        long long task = 0;
        if (lane == 0) task = (long long)atomicAdd(&counters[1], 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= totalSuper) break;
        const unsigned part = (unsigned)task & ((1u << splitLog2) - 1u);
        task >>= splitLog2;
        const int si = find_snapshot(superOffset, nSnaps, task, lane);
        __syncwarp();
        {
            const int* src = (const int*)(snaps + si);
            int* dst = (int*)&S.snap;
            if (lane < (int)(sizeof(Snapshot) / 4)) dst[lane] = __ldg(src + lane);
        }
        __syncwarp();
        const float ox = S.snap.eye[0], oy = S.snap.eye[1], oz = S.snap.eye[2];
        // the k-th occupied supertile of this pose
        int sbit;
        {
            unsigned long long m = __ldg(superMask + si);
            for (int k = (int)(task - (long long)__ldg(superOffset + si)); k > 0; --k) m &= m - 1;
            sbit = __ffsll((long long)m) - 1;
        }
        const unsigned ssx = sbit & 7, ssy = sbit >> 3;
        const uint32_t* poseMask = tileMask + (size_t)si * 32;

        // ---- phase 1: the table entries that touch this supertile, still front to back ----
        const int nEnt = __ldg(tableCount + si);
        const unsigned long long* tab = table + (size_t)si * stride;
        int L = 0;
        for (int base = 0; base < nEnt; base += 32 * kFilterUnroll) {  // kFilterUnroll loads in flight per lane
            unsigned long long ev[kFilterUnroll];
#pragma unroll
            for (int u = 0; u < kFilterUnroll; ++u) {
                const int i = base + 32 * u + lane;
                ev[u] = i < nEnt ? __ldg(tab + i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kFilterUnroll; ++u) {
                const int i = base + 32 * u + lane;
                const unsigned long long e = ev[u];
                const unsigned r = (unsigned)(e >> 28);
                const unsigned tx0 = (r >> 15) & 31u, tx1 = (r >> 10) & 31u, ty0 = (r >> 5) & 31u, ty1 = r & 31u;
                const bool ok = i < nEnt && (tx0 >> 2) <= ssx && (tx1 >> 2) >= ssx && (ty0 >> 2) <= ssy && (ty1 >> 2) >= ssy;
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) list[L + __popc(m & ltMask)] = e;
                L += __popc(m);
            }
        }
        myEntries += (unsigned long long)nEnt;
        __syncwarp();

        SuperGeom sgm;
        sgm.X0 = __fmaf_rn((float)((int)ssx * 4 * kTile), kx, cx);
        sgm.Y0 = __fmaf_rn((float)((int)ssy * 4 * kTile), ky, cy);
        sgm.ipx = 1.f / kx;
        sgm.ipy = 1.f / ky;
        if (lane < 16) {
            // the expressions of tile_geometry(): tXmin / tYmin, and tXmax / tYmax through the last 8x4 block of the tile
            const int k = lane & 3;
            const bool isY = (lane & 8) != 0, isMax = (lane & 4) != 0;
            const int p0 = ((int)(isY ? ssy : ssx) * 4 + k) * kTile;
            const float kk = isY ? ky : kx, cc = isY ? cy : cx;
            const float lo = __fmaf_rn((float)p0, kk, cc);
            const float hi = isY ? __fmaf_rn(7.f, 4.f * ky, __fmaf_rn((float)(p0 + 3), ky, cy)) : __fmaf_rn(3.f, 8.f * kx, __fmaf_rn((float)(p0 + 7), kx, cx));
            (isY ? S.tileY : S.tileX)[(isMax ? 4 : 0) + k] = isMax ? hi : lo;
        }
        sgm.tX = S.tileX;
        sgm.tY = S.tileY;
        __syncwarp();
        const float zFarCull = zFar * kDepthTie;

        // ---- regions of the supertile: the whole of it, or, when its triangles do not fit the record scratch, halves of it ----
        unsigned long long stack = (unsigned long long)(splitLog2 == 0   ? region_pack(0, 3, 0, 3)
                                                        : splitLog2 == 2 ? region_pack(2 * (part & 1u), 2 * (part & 1u) + 1, 2 * (part >> 1), 2 * (part >> 1) + 1)
                                                                         : region_pack(part & 3u, part & 3u, part >> 2, part >> 2));
        int depth = 1;
        mySuperRec++;
        while (depth > 0) {
            const unsigned region = (unsigned)(stack & 0xffull);
            stack >>= 8;
            depth--;
            const unsigned rx0 = region & 3u, rx1 = (region >> 2) & 3u, ry0 = (region >> 4) & 3u, ry1 = (region >> 6) & 3u;
            const unsigned rmask = region_mask(region);
            const bool single = rx0 == rx1 && ry0 == ry1;
            const int atx0 = (int)ssx * 4 + (int)rx0, atx1 = (int)ssx * 4 + (int)rx1, aty0 = (int)ssy * 4 + (int)ry0, aty1 = (int)ssy * 4 + (int)ry1;
            // a region without an occupied tile has nothing to do (the test hook visits every tile)
            if (!idImage) {
                bool any = false;
                for (int ty = aty0; ty <= aty1; ++ty) any = any || ((__ldg(poseMask + ty) >> atx0) & ((2u << (atx1 - atx0)) - 1u)) != 0u;
                if (!any) continue;
            }

            // what a single-tile region keeps across its batches: the tile's farthest hit and whether its hit buffers are in use
            // (everything else about a tile is rebuilt when its rasterisation starts or resumes: registers are scarce in phase 1.5)
            float zcullKeep = zFarCull;
            bool inited = false, tileHit = false;

            int resumeBase = 0, resumeP0 = 0;
            bool overflow = false;
            do {  // batches (only a single-tile region runs more than one)
                __syncwarp();  // the records and tile lists of the previous batch / region have been read by every lane
                // ---- phase 1.5: the pose half of the set-up, once per triangle of the list instead of once per tile that stages it.
                // A triangle that can produce a fragment in the region becomes a 64-byte record with a 4-byte header (depth key of its
                // cluster, tiles of the supertile its screen box touches); the others are dropped here, so that the tiles scan
                // fewer headers.
                int nRecs = 0;
                bool stop = false;
                for (int base = resumeBase; base < L && !stop; base += 32) {
                    const int i = base + lane;
                    const unsigned long long e = i < L ? list[i] : 0ull;
                    const unsigned r = (unsigned)(e >> 28);
                    const int tx0 = (r >> 15) & 31u, tx1 = (r >> 10) & 31u, ty0 = (r >> 5) & 31u, ty1 = r & 31u;
                    const bool inRegion = i < L && tx0 <= atx1 && tx1 >= atx0 && ty0 <= aty1 && ty1 >= aty0;
                    const unsigned ref = (unsigned)e & 0x0fffffffu;
                    const int first = inRegion ? __float_as_int(__ldg(&sc.leafLo[ref].w)) : 0;
                    const int cnt = inRegion ? __float_as_int(__ldg(&sc.leafHi[ref].w)) : 0;
                    const unsigned zq = (unsigned)(e >> 48);
                    if (single && base > resumeBase) {
                        // front to back: nothing behind the tile's farthest hit can matter any more
                        const float zl0 = (float)__shfl_sync(0xffffffffu, zq, 0) * iq;
                        if (inited && zl0 > zcullKeep * 1.000001f + 1e-4f) { resumeBase = L; resumeP0 = 0; break; }
                    }
                    int incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    const int excl = incl - cnt;
                    for (int p0 = (base == resumeBase ? resumeP0 : 0); p0 < total; p0 += 32) {
                        if (nRecs + 32 > kBinRecords) {  // the scratch is full: this batch ends here
                            stop = true;
                            resumeBase = base;
                            resumeP0 = p0;
                            break;
                        }
                        const int p = p0 + lane;
                        // the list entry that holds position p: the number of entries whose inclusive prefix is <= p
                        int src = 0;
#pragma unroll
                        for (int sft = 16; sft; sft >>= 1) {
                            const int v = __shfl_sync(0xffffffffu, incl, min(src + sft - 1, 31));
                            if (v <= p) src += sft;
                        }
                        src = min(src, 31);
                        const int tri = __shfl_sync(0xffffffffu, first, src) + (p - __shfl_sync(0xffffffffu, excl, src));
                        const unsigned zqe = __shfl_sync(0xffffffffu, zq, src);
                        unsigned tiles = 0u;
                        bool keep = false;
                        if (p < total) {
                            keep = setup_triangle_t<true>(sc, &S.snap, tri, S.tri[lane], ox, oy, oz, zNear, zFarCull, TileGeom(), &sgm, &tiles);
                            tiles &= rmask;
                            keep = keep && tiles != 0u;
                        }
                        myTris += (unsigned)min(32, total - p0);
                        const unsigned km = __ballot_sync(0xffffffffu, keep);
                        if (km == 0u) continue;
                        const int pos = nRecs + __popc(km & ltMask);
                        if (keep) {
                            const float4* srcRow = reinterpret_cast<const float4*>(S.tri[lane]);
                            float4* dst = rec + (size_t)pos * 4;
                            dst[0] = srcRow[0]; dst[1] = srcRow[1]; dst[2] = srcRow[2]; dst[3] = srcRow[3];
                        }
                        if (keep) hdr[pos] = (zqe << 16) | tiles;
                        nRecs += __popc(km);
                    }
                    if (!stop && base + 32 >= L) { resumeBase = L; resumeP0 = 0; }
                }
                if (L == 0 || resumeBase >= L) stop = false;
                overflow = stop;
                if (overflow && !single) break;  // the region is split instead
                __syncwarp();

                // ---- phase 2: the occupied tiles of the region ----
                for (int tl = 0; tl < 16; ++tl) {
                    if (!((rmask >> tl) & 1u)) continue;
                    const int tx = (int)ssx * 4 + (tl & 3), ty = (int)ssy * 4 + (tl >> 2);
                    if (tx * kTile >= cam.W || ty * kTile >= cam.H) continue;
                    const bool occupied = (__ldg(poseMask + ty) >> tx) & 1u;
                    if (!occupied && !idImage) continue;
                    const int tpx0 = tx * kTile, tpy0 = ty * kTile;
                    const int vw = min(kTile, cam.W - tpx0), vh = min(kTile, cam.H - tpy0);  // valid pixels of a border tile
                    if (nRecs == 0 && !idImage && !(single && inited)) {  // nothing of the region's triangles can touch any tile
                        myRays += (unsigned long long)(vw * vh);
                        myTiles++;
                        continue;
                    }
                    TileGeom tg;
                    tile_geometry(tg, tpx0, tpy0, lane, kx, cx, ky, cy);
                    tg.sox = (tl & 3) * kTile;
                    tg.soy = (tl >> 2) * kTile;
                    // zcull: nothing beyond this depth can matter to any ray of the tile; lane b keeps the same for 8x4 block b in bzReg
                    // (0: block outside the image)
                    float zcull = zFarCull, bzReg = ((lane & 3) * 8 < vw && (lane >> 2) * 4 < vh) ? zFarCull : 0.f;
                    if (single && inited) refresh_depth_bounds(S.t, lane, bzReg, zcull);  // a later batch of the same tile
                    else { inited = false; tileHit = false; }
                    // ---- flush: stage up to 32 records in parallel, one per lane, then test them kSubRound at a time ----
                    auto flush = [&](bool mineValid, int myRec) {
                        __syncwarp();
                        bool mineOk = false;
                        if (mineValid) mineOk = stage_record(rec + (size_t)myRec * 4, S.tri[lane], zcull, tg);
                        const unsigned live = __ballot_sync(0xffffffffu, mineOk);  // triangles that can produce a fragment in this tile
#ifdef IPP_VIS_DIAG
                        dLive += __popc(live);
                        dStaged += __popc(__ballot_sync(0xffffffffu, mineValid));
#if IPP_VIS_DIAG == 2
                        // how many of the rejected ones fall to the depth test alone (reported instead of the open tiles)
                        dOpen += __popc(__ballot_sync(0xffffffffu, mineValid && !mineOk && rec[(size_t)myRec * 4 + 3].z > zcull * 1.000001f));
#endif
#endif
                        if (live) {
                            if (!inited) {
                                tile_init(S.t, S.id, lane, vw, vh, zFar);
                                inited = true;
                            }
                            __syncwarp();
                            // kSubRound triangles at a time: the farthest hits of the tile and of its 8x4 blocks are refreshed in
                            // between, so that what the front-most triangles of the round cover already culls the ones behind them
#pragma unroll 1
                            for (int s0 = 0; s0 < 32; s0 += kSubRound) {
                                const unsigned sub = live & (kSubRound >= 32 ? 0xffffffffu : (((1u << (kSubRound & 31)) - 1u) << s0));
                                if (sub == 0u) continue;
#ifdef IPP_VIS_DIAG
                                unsigned nb = 0;
                                const bool dirty = raster_staged<true>(sub, S.tri, S.t, S.id, tg, lane, zNear, zFar, zcull, bzReg, &nb);
                                dBlocks += nb;
#else
                                const bool dirty = raster_staged<true>(sub, S.tri, S.t, S.id, tg, lane, zNear, zFar, zcull, bzReg);
#endif
                                if (__any_sync(0xffffffffu, dirty)) {
                                    tileHit = true;
                                    __syncwarp();
                                    refresh_depth_bounds(S.t, lane, bzReg, zcull);
                                }
                            }
                        }
                    };
                    // the record headers, front to back: a lane keeps the records whose triangle can touch this tile; 32 of them make a
                    // staging round (the next chunk of headers is requested before the current one is worked on)
                    int nq = 0;
                    unsigned hNext = (occupied && lane < nRecs) ? hdr[lane] : 0xffff0000u;
                    for (int base = 0; occupied && base < nRecs; base += 32) {
                        const float bound = zcull * 1.000001f + 1e-4f;
                        const int i = base + lane;
                        const unsigned h = hNext;
                        hNext = (i + 32 < nRecs) ? hdr[i + 32] : 0xffff0000u;
                        const float zl = (float)(h >> 16) * iq;
                        // sorted by cluster depth: when the first record of the chunk is behind the farthest hit, so is everything after it
                        if (__shfl_sync(0xffffffffu, zl, 0) > bound) break;
                        myEntries += 16ull;  // 32 headers of 4 bytes
                        const bool hit = i < nRecs && ((h >> tl) & 1u) != 0u && zl <= bound;
                        const unsigned m = __ballot_sync(0xffffffffu, hit);
                        if (m == 0u) continue;
                        myTris += __popc(m);
                        // the record is staged a few header chunks from now: have it in L1 by then (A/B: -0.9 % on config 2, -1.2 % on luis4)
                        if (hit) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)i * 4));
                        if (hit) q16[nq + __popc(m & ltMask)] = (uint16_t)i;
                        nq += __popc(m);
                        __syncwarp();
                        if (nq >= 32) {
                            const int myRec = q16[lane];
                            const uint16_t moved = q16[32 + lane];
                            __syncwarp();
                            q16[lane] = moved;
                            nq -= 32;
                            flush(true, myRec);
                        } else if (nq >= kStageMin) {  // a short round now keeps the depth bound, and with it the early exit, fresh
                            const int myRec = q16[lane];
                            const bool mine = lane < nq;
                            nq = 0;
                            flush(mine, myRec);
                        }
                    }
                    if (nq > 0) flush(lane < nq, q16[lane]);
                    __syncwarp();
                    zcullKeep = zcull;
#ifdef IPP_VIS_DIAG
#if IPP_VIS_DIAG != 2
                    if (zcull >= zFarCull) dOpen++;
#endif
#endif
                    if (single && overflow) continue;  // more batches of this tile's triangles follow
                    // ---- the winners ----
                    if (idImage && !inited) tile_init(S.t, S.id, lane, vw, vh, zFar);  // test hook: background pixels are reported too
                    __syncwarp();
                    if (tileHit || idImage) tile_output(S.id, lane, flags, (size_t)sc.Tpad, slotList + S.snap.slot, S.snap.slotCount, idImage, tpx0, tpy0, cam.W, cam.H);
                    myRays += (unsigned long long)(vw * vh);
                    myTiles++;
                    __syncwarp();
                }
            } while (single && overflow);
            if (overflow && !single) {
                // halves, the longer side first
                mySuperSplit++;
                if (rx1 - rx0 >= ry1 - ry0) {
                    const unsigned xm = (rx0 + rx1) >> 1;
                    stack = (stack << 16) | ((unsigned long long)region_pack(xm + 1, rx1, ry0, ry1) << 8) | region_pack(rx0, xm, ry0, ry1);
                } else {
                    const unsigned ym = (ry0 + ry1) >> 1;
                    stack = (stack << 16) | ((unsigned long long)region_pack(rx0, rx1, ym + 1, ry1) << 8) | region_pack(rx0, rx1, ry0, ym);
                }
                depth += 2;
            }
        }
    }
    if (lane == 0 && (myTiles || mySuperRec)) {
        atomicAdd(&counters[0], myRays);
        atomicAdd(&counters[3], (unsigned long long)myTris);
        atomicAdd(&counters[4], (unsigned long long)myTiles);
        atomicAdd(&counters[5], myEntries);
#ifdef IPP_VIS_DIAG
        atomicAdd(&counters[6], dLive);
        atomicAdd(&counters[7], dOpen);
        atomicAdd(&counters[8], dBlocks);   // reported as "supertiles" by ipp_visibility_stats in such a build
        atomicAdd(&counters[9], dStaged);   // ... and "supertile_splits"
#else
        atomicAdd(&counters[8], (unsigned long long)mySuperRec);
        atomicAdd(&counters[9], (unsigned long long)mySuperSplit);
#endif
    }
}

template <int THREADS, int ITEMS>
void launch_leaf_table_t(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, int nSnaps, const BinBuffers& b,
                         cudaStream_t st) {
    using Sort = cub::BlockRadixSort<unsigned long long, THREADS, ITEMS, cub::NullType, kSortRadixBits>;
    const size_t smem = std::max(sizeof(typename Sort::TempStorage), (size_t)THREADS * ITEMS * sizeof(unsigned long long));
    // per launch: the attribute belongs to the current device's context, and a process may own evaluators on several devices
    IPP_CUDA_CHECK(cudaFuncSetAttribute(k_leaf_table<THREADS, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_leaf_table<THREADS, ITEMS><<<nSnaps, THREADS, smem, st>>>(sc, cam, d_snaps, b.stride, b.table, b.tableCount, b.tileMask, b.superMask,
                                                               b.nSuper);
}

}  // namespace

bool bins_supported(const DeviceScene& sc, const CameraParams& cam) {
    return sc.nLeaves > 0 && sc.nLeaves <= kMaxBinLeaves && cam.W <= 1024 && cam.H <= 1024 && sc.T < (1 << 24);
}

int bins_grid_blocks() {
    int dev = 0, sms = 148, perSM = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_visibility_bins, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_visibility_bins, kBinThreads, 0);
    if (perSM < 1) perSM = 1;
    return sms * perSM;
}
int bins_warps_per_block() { return kBinWarps; }
size_t bins_record_floats_per_warp() { return (size_t)kBinRecords * 17; }  // 64-byte records + 4-byte headers

// size classes of the per-pose sort (threads x keys per thread); stride of a pose's table = capacity of its class
struct SortClass { int capacity; };
static const int kSortCaps[] = {1024, 2048, 3072, 4096, 6144, 8192, 12288, 16384};
int bins_table_stride(int nLeaves) {
    for (int c : kSortCaps)
        if (nLeaves <= c) return c;
    return kMaxBinLeaves;
}

void launch_leaf_table(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, int nSnaps, const BinBuffers& b,
                       cudaStream_t st) {
    if (nSnaps <= 0) return;
    switch (bins_table_stride(sc.nLeaves)) {
        case 1024: launch_leaf_table_t<256, 4>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 2048: launch_leaf_table_t<256, 8>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 3072: launch_leaf_table_t<256, 12>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 4096: launch_leaf_table_t<256, 16>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 6144: launch_leaf_table_t<512, 12>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 8192: launch_leaf_table_t<512, 16>(sc, cam, d_snaps, nSnaps, b, st); break;
        case 12288: launch_leaf_table_t<1024, 12>(sc, cam, d_snaps, nSnaps, b, st); break;
        default: launch_leaf_table_t<1024, 16>(sc, cam, d_snaps, nSnaps, b, st); break;
    }
}

void launch_visibility_bins(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, const int* d_superOffset, int nSnaps,
                            int64_t totalSuper, const BinBuffers& b, uint8_t* d_flags, const int* d_slotList, int32_t* d_idImage,
                            unsigned long long* d_counters, cudaStream_t st) {
    if (totalSuper <= 0 || nSnaps <= 0) return;
    // fewer supertiles than two per resident warp: hand them out as quadrants or single tiles (a supertile is up to 16 tiles of serial
    // work for one warp, which is all a small launch would last).  The pieces repeat the table filter and part of the set-up, so
    // this only pays while most warps would otherwise idle -- measured on luis4: 60 poses 3.0 -> 1.3 ms as single tiles, 300 poses
    // 4.4 ms whole against 5.0 ms as quadrants, 1 500 poses 13.8 / 14.3 / 23.4 ms whole / quadrants / tiles.
    const long long warps = (long long)b.gridBlocks * kBinWarps;
    int splitLog2 = totalSuper >= 2 * warps ? 0 : (totalSuper * 4 >= 2 * warps ? 2 : 4);
    if (const char* e = getenv("IPP_BINS_SPLIT")) {  // tests and A/B runs: 0, 2 or 4
        const int v = atoi(e);
        if (v == 0 || v == 2 || v == 4) splitLog2 = v;
    }
    long long groups = ((totalSuper << splitLog2) + kBinWarps - 1) / kBinWarps;
    int grid = (int)std::min<long long>(b.gridBlocks, groups);
    // counters[1] is the work cursor of this launch
    cudaMemsetAsync(d_counters + 1, 0, sizeof(unsigned long long), st);
    k_visibility_bins<<<grid, kBinThreads, 0, st>>>(sc, cam, d_snaps, d_superOffset, nSnaps, b.table, b.stride, b.tableCount, b.tileMask,
                                                   b.superMask, b.scratch, b.records, d_flags, d_slotList, d_idImage, d_counters, splitLog2);
}

}  // namespace ipp
