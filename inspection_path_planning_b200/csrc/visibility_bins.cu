// visibility_bins.cu -- K4b: edge visibility with depth-sorted leaf bins instead of a BVH walk per tile.  Same item-buffer sampling
// and the same per-tile rasteriser as visibility.cu (raster.cuh); what changes is how a tile finds its triangles.
//
//   k_leaf_table  one CTA per camera pose: every BVH leaf box (a few thousand for the reference's meshes) is projected -- 8 corners,
//                 screen rectangle in 32x32-pixel tile units with a one-pixel margin, nearest depth -- culled against the view
//                 volume, packed into one 64-bit word [depth:16 | tx0 tx1 ty0 ty1:20 | first triangle, count:28] and block-radix
//                 sorted by depth (cub::BlockRadixSort, stable, so the order is deterministic).  The CTA also ORs the rectangles
//                 into a 32x32 tile occupancy mask and an 8x8 supertile mask: empty tiles are never scheduled.
//   k_visibility_bins  persistent grid, one warp per occupied 128x128 supertile: one coalesced pass over the pose's table keeps the
//                 entries that touch the supertile (per-warp list in an L2-resident scratch); the tile-independent half of the
//                 triangle set-up then runs once per triangle of that list into 64-byte records (a second per-warp scratch) with
//                 a 4-byte header each (depth key of the leaf, tiles of the supertile the triangle's screen box touches), and
//                 the warp rasterises its <= 16 tiles one after the other: a tile scans the headers front to back, stages the
//                 records that can touch it, and stops at the first header behind its farthest hit (a list with more triangles
//                 than the record scratch holds keeps the older scheme: walk the leaves, set their triangles up per tile) --
//                 strict front-to-back order makes
//                 the hierarchical depth culling of raster.cuh bite much earlier than the approximate order of a BVH walk, and
//                 there is no dependent node fetch left on the critical path.
//
// Used when the mesh has <= kMaxLeaves leaves and the image is <= 1024 pixels a side (every asset of the reference); larger
// inputs take the BVH-walking kernel of visibility.cu.
#include <cub/block/block_radix_sort.cuh>

#include "raster.cuh"

namespace ipp {

namespace {

using namespace raster;

constexpr int kBinWarps = 4;
constexpr int kBinThreads = kBinWarps * 32;
constexpr unsigned long long kNoEntry = ~0ull;
constexpr int kSortRadixBits = 6;  // 12 depth bits = 2 passes
#ifndef IPP_STAGE_MIN
#define IPP_STAGE_MIN 8
#endif
constexpr int kStageMin = IPP_STAGE_MIN;  // waiting records that make a staging round worth running at the end of a header chunk
constexpr int kBinRecords = 4096;   // set-up triangles (64 B each) a warp can keep for one supertile
#ifndef IPP_FILTER_UNROLL
#define IPP_FILTER_UNROLL 4
#endif
constexpr int kFilterUnroll = IPP_FILTER_UNROLL;    // table loads a lane keeps in flight while a supertile's list is filtered

struct __align__(16) BinShared {
    float t[kTilePixels];
    int id[kTilePixels];
    float tri[kPending][kTriWords];
    uint8_t queue[32 * kLeafMax];  // triangles of one list chunk waiting for set-up: (source lane << 3) | triangle of its leaf
    Snapshot snap;
};
static_assert(kLeafMax <= 8, "queue entries hold 3 bits of triangle index");

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
        const unsigned ref = (unsigned)__float_as_int(lo4.w) & 0x0fffffffu;
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
    Sort(temp).Sort(keys, 52, 64);  // stable: entries of equal depth keep their (fixed) input order
    __syncthreads();
    unsigned long long* out = table + (size_t)si * stride;
    const int n = nValid;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int pos = tid * ITEMS + it;
        if (pos < n) out[pos] = keys[it];
    }
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
#ifdef IPP_VIS_DIAG
__device__ unsigned long long g_diag[1];  // diagnostic build: triangles staged in tiles that stayed open
#endif
__global__ void __launch_bounds__(kBinThreads, 5) k_visibility_bins(DeviceScene sc, CameraParams cam, const Snapshot* __restrict__ snaps,
                                                                 const int* __restrict__ superOffset, int nSnaps,
                                                                 const unsigned long long* __restrict__ table, int stride,
                                                                 const int* __restrict__ tableCount, const uint32_t* __restrict__ tileMask,
                                                                 const unsigned long long* __restrict__ superMask,
                                                                 unsigned long long* scratch, float4* recordsAll, uint8_t* __restrict__ flags,
                                                                 int32_t* __restrict__ idImage, unsigned long long* __restrict__ counters) {
    __shared__ BinShared s_all[kBinWarps];
    const int warp = threadIdx.x >> 5;
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // volatile: kept in a register instead of re-read (S2R) inside the loops
    BinShared& S = s_all[warp];
    unsigned long long* list = scratch + (size_t)(blockIdx.x * kBinWarps + warp) * stride;
    float4* rec = recordsAll + (size_t)(blockIdx.x * kBinWarps + warp) * ((kBinRecords + 32) * 17 / 4);
    unsigned* hdr = reinterpret_cast<unsigned*>(rec + (size_t)(kBinRecords + 32) * 4);  // one word per record: leaf depth key:16 | tiles:16
    uint16_t* q16 = reinterpret_cast<uint16_t*>(S.queue);                                // phase 2: records waiting for a staging round
    unsigned long long myRays = 0, myEntries = 0;
    unsigned myTris = 0, myTiles = 0, mySuperRec = 0, mySuperWalk = 0;  // statistics (lane 0 only matters)
#ifdef IPP_VIS_DIAG
    unsigned myLive = 0, myOpen = 0, myOpenTris = 0;
#endif
    const float kx = (float)(2.0 * cam.tanX / cam.W), cx = (float)((1.0 / cam.W - 1.0) * cam.tanX);
    const float ky = (float)(2.0 * cam.tanY / cam.H), cy = (float)((1.0 / cam.H - 1.0) * cam.tanY);
    const float zNear = cam.zNear, zFar = cam.zFar;
    const float iq = (zFar * kDepthTie * 1.000001f + 1e-4f) / 65535.f;  // depth of a quantised table entry (a lower bound)
    const long long totalSuper = (long long)__ldg(superOffset + nSnaps);
    const unsigned ltMask = (1u << lane) - 1u;

    while (true) {
        long long task = 0;
        if (lane == 0) task = (long long)atomicAdd(&counters[1], 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= totalSuper) break;
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

        // ---- phase 1.5: the pose half of the set-up, once per triangle of the list instead of once per tile that stages it ----
        // (a leaf rectangle covers most of the 16 tiles, so a triangle was set up by ~10 of them).  The 64-byte records and their
        // 4-byte headers go to this warp's part of `recordsAll` (L2-resident) in list order; a list with more than kBinRecords
        // triangles keeps the leaf walk with its per-tile set-up.
        bool cached;
        int nRec = 0;
        {
            for (int i = lane; i < L; i += 32) nRec += (int)((unsigned)list[i] & kLeafMask);
#pragma unroll
            for (int d = 16; d; d >>= 1) nRec += __shfl_xor_sync(0xffffffffu, nRec, d);
            cached = nRec <= kBinRecords;
        }
        if (cached) mySuperRec++;
        else mySuperWalk++;
        if (cached) {
            TileGeom none = {};
            SuperGeom sgm;
            sgm.X0 = __fmaf_rn((float)((int)ssx * 4 * kTile), kx, cx);
            sgm.Y0 = __fmaf_rn((float)((int)ssy * 4 * kTile), ky, cy);
            sgm.ipx = 1.f / kx;
            sgm.ipy = 1.f / ky;
            const float zFarCull = zFar * kDepthTie;
            int recBase = 0;
            for (int base = 0; base < L; base += 32) {
                const int i = base + lane;
                const unsigned long long e = i < L ? list[i] : 0ull;
                const int first = (int)(((unsigned)e & 0x0fffffffu) >> kLeafShift);
                const int cnt = i < L ? (int)((unsigned)e & kLeafMask) : 0;
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
                for (int q = 0; q < cnt; ++q) S.queue[incl - cnt + q] = (uint8_t)((lane << 3) | q);
                __syncwarp();
                myTris += total;
                for (int p0 = 0; p0 < total; p0 += 32) {
                    const int p = p0 + lane;
                    const unsigned b = p < total ? S.queue[p] : 0u;
                    const int tri = __shfl_sync(0xffffffffu, first, b >> 3) + (int)(b & 7u);
                    const unsigned zq = __shfl_sync(0xffffffffu, (unsigned)(e >> 48), b >> 3);
                    unsigned tiles = 0u;  // a rejected triangle is never staged
                    if (p < total) {
                        setup_triangle_t<true>(sc, &S.snap, tri, S.tri[lane], ox, oy, oz, zNear, zFarCull, none, &sgm, &tiles);
                        hdr[recBase + p] = (zq << 16) | tiles;
                    }
                    __syncwarp();
                    // 32 rows of 64 B in shared memory -> 2 KB of records, coalesced (rows past `total` are never read)
                    const float4* src = reinterpret_cast<const float4*>(&S.tri[0][0]);
                    float4* dst = rec + (size_t)(recBase + p0) * 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[k * 32 + lane] = src[k * 32 + lane];
                    __syncwarp();
                }
                recBase += total;
            }
        }
        __syncwarp();

        // ---- phase 2: the occupied tiles of the supertile ----
        for (int tl = 0; tl < 16; ++tl) {
            const int tx = (int)ssx * 4 + (tl & 3), ty = (int)ssy * 4 + (tl >> 2);
            const int tpx0 = tx * kTile, tpy0 = ty * kTile;
            if (tpx0 >= cam.W || tpy0 >= cam.H) continue;
            const bool occupied = (__ldg(tileMask + (size_t)si * 32 + ty) >> tx) & 1u;
            if (!occupied && !idImage) continue;
            const int vw = min(kTile, cam.W - tpx0), vh = min(kTile, cam.H - tpy0);  // valid pixels of a border tile
            float zcull = zFar * kDepthTie;  // nothing beyond this depth can matter to any ray of the tile
            // lane b keeps the depth beyond which nothing can matter to any ray of 8x4 block b (0: block outside the image)
            float bzReg = ((lane & 3) * 8 < vw && (lane >> 2) * 4 < vh) ? zcull : 0.f;
            TileGeom tg;
            tile_geometry(tg, tpx0, tpy0, lane, kx, cx, ky, cy);
            bool inited = false, tileHit = false;
#ifdef IPP_VIS_DIAG
            const unsigned trisBefore = myTris;
#endif
            // Triangles wait in lane order: lanes < pend carry one each (myTri, its leaf's depth myZ) from the previous chunk.
            int pend = 0, myTri = 0;
            float myZ = 0.f;
            // ---- flush: set up to 32 triangles up in parallel, one per lane, then test them one after the other ----
            auto flush = [&](bool mineValid) {
                __syncwarp();
                bool mineOk = false;
                if (mineValid) {
                    if (cached) mineOk = stage_record(rec + (size_t)myTri * 4, S.tri[lane], zcull, tg);
                    else mineOk = setup_triangle(sc, &S.snap, myTri, S.tri[lane], ox, oy, oz, zNear, zcull, tg);
                }
                const unsigned live = __ballot_sync(0xffffffffu, mineOk);  // triangles that can produce a fragment in this tile
#ifdef IPP_VIS_DIAG
                myLive += __popc(live);
#endif
                if (live) {
                    if (!inited) {
                        tile_init(S.t, S.id, lane, vw, vh, zFar);
                        inited = true;
                    }
                    __syncwarp();
                    const bool dirty = raster_staged(live, S.tri, S.t, S.id, tg, lane, zNear, zFar, zcull, bzReg);
                    if (__any_sync(0xffffffffu, dirty)) {
                        tileHit = true;
                        __syncwarp();
                        refresh_depth_bounds(S.t, lane, bzReg, zcull);
                    }
                }
            };
            if (cached) {
                // the record headers, front to back: a lane keeps the records whose triangle can touch this tile; 32 of them make a
                // staging round
                int nq = 0;
                for (int base = 0; occupied && base < nRec; base += 32) {
                    const float bound = zcull * 1.000001f + 1e-4f;
                    const int i = base + lane;
                    const unsigned h = i < nRec ? hdr[i] : 0xffff0000u;
                    const float zl = (float)(h >> 16) * iq;
                    // sorted by leaf depth: when the first record of the chunk is behind the farthest hit, so is everything after it
                    if (__shfl_sync(0xffffffffu, zl, 0) > bound) break;
                    myEntries += 16ull;  // 32 headers of 4 bytes
                    const bool hit = i < nRec && ((h >> tl) & 1u) != 0u && zl <= bound;
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m == 0u) continue;
                    myTris += __popc(m);
                    if (hit) q16[nq + __popc(m & ltMask)] = (uint16_t)i;
                    nq += __popc(m);
                    __syncwarp();
                    if (nq >= 32) {
                        myTri = q16[lane];
                        const uint16_t moved = q16[32 + lane];
                        __syncwarp();
                        q16[lane] = moved;
                        nq -= 32;
                        flush(true);
                    } else if (nq >= kStageMin) {  // a short round now keeps the depth bound, and with it the early exit, fresh
                        myTri = q16[lane];
                        const bool mine = lane < nq;
                        nq = 0;
                        flush(mine);
                    }
                }
                if (nq > 0) {
                    myTri = q16[lane];
                    flush(lane < nq);
                }
            } else {
            for (int base = 0; occupied && base < L; base += 32) {
                float bound = zcull * 1.000001f + 1e-4f;
                const int i = base + lane;
                const unsigned long long e = i < L ? list[i] : kNoEntry;
                const float zl = (float)(unsigned)(e >> 48) * iq;
                // sorted by depth: when the first entry of the chunk is behind the farthest hit, so is everything after it
                if (__shfl_sync(0xffffffffu, zl, 0) > bound) break;
                myEntries += 32ull;
                const unsigned r = (unsigned)(e >> 28);
                const int tx0 = (r >> 15) & 31u, tx1 = (r >> 10) & 31u, ty0 = (r >> 5) & 31u, ty1 = r & 31u;
                const bool hit = i < L && tx0 <= tx && tx <= tx1 && ty0 <= ty && ty <= ty1 && zl <= bound;
                if (!__any_sync(0xffffffffu, hit)) continue;
                // the triangles of the leaves that contain the tile, in list order: leaf k contributes positions [P_k, P_k + cnt_k)
                const int first = (int)(((unsigned)e & 0x0fffffffu) >> kLeafShift);
                const int cnt = hit ? (int)((unsigned)e & kLeafMask) : 0;
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
                for (int q = 0; q < cnt; ++q) S.queue[incl - cnt + q] = (uint8_t)((lane << 3) | q);  // (source lane, triangle of its leaf)
                __syncwarp();
                const int pend0 = pend, avail = pend0 + total;
                myTris += total;
                // position p of the waiting line: p < pend is lane p's carried triangle, otherwise queue[p - pend]
                int p = lane;
                for (; p - lane + 32 <= avail; p += 32) {
                    const int qp = p - pend0;
                    const unsigned b = qp >= 0 ? S.queue[qp] : 0u;
                    const int tri = __shfl_sync(0xffffffffu, first, b >> 3) + (int)(b & 7u);
                    const float z = __shfl_sync(0xffffffffu, zl, b >> 3);
                    if (qp >= 0) { myTri = tri; myZ = z; }
                    flush(myZ <= bound);  // a flush may have pulled the farthest hit in front of this leaf since it was queued
                    bound = zcull * 1.000001f + 1e-4f;
                }
                {   // what is left waits in lanes 0 .. avail - (p - lane) - 1
                    const int qp = p - pend0;
                    const bool take = p < avail && qp >= 0;
                    const unsigned b = take ? S.queue[qp] : 0u;
                    const int tri = __shfl_sync(0xffffffffu, first, b >> 3) + (int)(b & 7u);
                    const float z = __shfl_sync(0xffffffffu, zl, b >> 3);
                    if (take) { myTri = tri; myZ = z; }
                    pend = avail - (p - lane);
                }
            }
            if (pend > 0) flush(lane < pend && myZ <= zcull * 1.000001f + 1e-4f);
            }
            __syncwarp();
            // ---- the winners ----
            if (idImage && !inited) tile_init(S.t, S.id, lane, vw, vh, zFar);  // test hook: background pixels are reported too
            __syncwarp();
            if (tileHit || idImage) tile_output(S.id, lane, flags + (size_t)S.snap.slot * sc.Tpad, idImage, tpx0, tpy0, cam.W, cam.H);
            myRays += (unsigned long long)(vw * vh);
            myTiles++;
#ifdef IPP_VIS_DIAG
            if (zcull == zFar * kDepthTie) {  // some ray of the tile never hit: the list was walked to its end
                myOpen++;
                myOpenTris += myTris - trisBefore;
            }
#endif
            __syncwarp();
        }
    }
    if (lane == 0 && myTiles) {
        atomicAdd(&counters[0], myRays);
        atomicAdd(&counters[3], (unsigned long long)myTris);
        atomicAdd(&counters[4], (unsigned long long)myTiles);
        atomicAdd(&counters[5], myEntries);
        atomicAdd(&counters[8], (unsigned long long)mySuperRec);
        atomicAdd(&counters[9], (unsigned long long)mySuperWalk);
#ifdef IPP_VIS_DIAG
        atomicAdd(&counters[6], (unsigned long long)myLive);
        atomicAdd(&counters[7], (unsigned long long)myOpen);
        atomicAdd(&g_diag[0], (unsigned long long)myOpenTris);
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
size_t bins_record_floats_per_warp() { return (size_t)(kBinRecords + 32) * 17; }  // 64-byte records + 4-byte headers

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
                            int64_t totalSuper, const BinBuffers& b, uint8_t* d_flags, int32_t* d_idImage, unsigned long long* d_counters,
                            cudaStream_t st) {
    if (totalSuper <= 0 || nSnaps <= 0) return;
    long long groups = (totalSuper + kBinWarps - 1) / kBinWarps;
    int grid = (int)std::min<long long>(b.gridBlocks, groups);
    // counters[1] is the work cursor of this launch
    cudaMemsetAsync(d_counters + 1, 0, sizeof(unsigned long long), st);
    k_visibility_bins<<<grid, kBinThreads, 0, st>>>(sc, cam, d_snaps, d_superOffset, nSnaps, b.table, b.stride, b.tableCount, b.tileMask,
                                                   b.superMask, b.scratch, b.records, d_flags, d_idImage, d_counters);
#ifdef IPP_VIS_DIAG
    unsigned long long d[1];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(d, g_diag, sizeof(d));
    fprintf(stderr, "[vis diag, cumulative] triangles staged in open tiles %llu\n", d[0]);
#endif
}

}  // namespace ipp
