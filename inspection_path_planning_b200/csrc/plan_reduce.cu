// plan_reduce.cu -- K6 plan_reduce: the HBM-bound kernel of the warm path.
//
// Per plan that passed the energy gate: OR the memoised seen-triangle bitmaps of its E edges
// (evaluateMemoisedPlan, PlanCoverageEstimator.cpp:43-118) and turn the union into the coverage score
// 1 - sum(area[i] * bit[i]) / totalArea (TriangleData::calculateAndColorCoverage, TriangleData.cpp:155-194).
//
// Layout: one warp per plan, 8 plans per CTA marching together over tiles of 128 bitmap words (4096
// triangles).  Per tile a lane streams one uint4 (16 B) of each of the E rows straight from HBM --
// E independent 512-byte coalesced requests in flight per warp -- and ORs them in registers; the 4096
// fp64 areas of the tile are staged once per CTA in shared memory with cp.async (double buffered) and
// shared by the 8 plans; the union words are broadcast with shuffles, lane b adds area[32 w + b] when bit
// b of word w is set (ballot-selected non-zero words only), and a final shuffle tree gives the covered
// area.  No atomics, no per-individual bitmap ever touches memory.
// Algorithmic bytes per plan: E * 4 * ceil(T/32) + 4 * (V + 1) + 32  (SURVEY 8d).
// Compiled with -fmad=false (only fp64 adds and one division are involved).
#include <cstdlib>

#include "kernels.h"

namespace ipp {

namespace {

constexpr int kPlansPerCta = 8;
constexpr int kPlansPerCtaWide = 12;  // large populations (launch_plan_reduce)
constexpr int kBlockTiles = 8;  // tiles whose area sums a lane adds up before the warp reduces them (the unit a split respects)
constexpr int kThreads = kPlansPerCta * 32;
constexpr int kSmemBytes = 2 * kTileTris * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    // rows are read once per plan and never reused by this CTA: do not allocate in L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Adds the areas of the set bits of one 128-word tile of a plan's union bitmap.  m = this lane's four words of the tile; word w of
// lane q covers triangles 128 q + 32 w .. + 31 of the tile, and lane b of the warp adds area[...+ b] when bit b is set (the words are
// broadcast with shuffles, non-zero lanes only).  Four independent partial sums per lane (one per word) keep the fp64 add chain
// short -- a single sum made every tile wait for 128 dependent adds during which the warp had no row loads in flight; every K6
// kernel uses this routine and combine_area(), so they all return the same bits.
struct AreaAcc {
    double s[4];
    __device__ __forceinline__ void clear() { s[0] = s[1] = s[2] = s[3] = 0.0; }
};
__device__ __forceinline__ void add_tile_area(AreaAcc& acc, const uint4& m, const double* __restrict__ ar, int lane) {
    unsigned nz = __ballot_sync(0xffffffffu, (m.x | m.y | m.z | m.w) != 0u);
    while (nz) {
        const int q = __ffs(nz) - 1;
        nz &= nz - 1;
        const unsigned w0 = __shfl_sync(0xffffffffu, m.x, q), w1 = __shfl_sync(0xffffffffu, m.y, q),
                       w2 = __shfl_sync(0xffffffffu, m.z, q), w3 = __shfl_sync(0xffffffffu, m.w, q);
        const double* base = ar + q * 128 + lane;
        if ((w0 >> lane) & 1u) acc.s[0] += base[0];
        if ((w1 >> lane) & 1u) acc.s[1] += base[32];
        if ((w2 >> lane) & 1u) acc.s[2] += base[64];
        if ((w3 >> lane) & 1u) acc.s[3] += base[96];
    }
}
// this lane's share, then the warp's (butterfly: every lane ends with the same sum)
__device__ __forceinline__ double combine_area(const AreaAcc& acc) {
    double v = (acc.s[0] + acc.s[1]) + (acc.s[2] + acc.s[3]);
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The streaming kernel of the warm path: rows through rowmap[key], every CTA walks all tiles of its plans.  Kept apart from the
// extended kernel below on purpose: this loop is L2 / HBM bound and ptxas' schedule of it is fragile (one extra branch in the row
// lookup or a reduce inside the tile loop cost 7-9 % of its bandwidth when both lived in one template).
template <int kPlansPerCta>
__global__ void __launch_bounds__(kPlansPerCta * 32) k_plan_reduce(PlanReduceArgs a) {
    constexpr int kThreads = kPlansPerCta * 32;
    extern __shared__ __align__(16) double s_area_raw[];  // 2 x 32 KB
    double(*s_area)[kTileTris] = (double(*)[kTileTris])s_area_raw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowVec = (int)(a.rowWords / 4);  // uint4 per row
    for (long long p0 = (long long)blockIdx.x * kPlansPerCta; p0 < a.pop; p0 += (long long)gridDim.x * kPlansPerCta) {
        long long p = p0 + warp;
        // ---- this warp's plan: edge list -> row indices (lane e holds edge e of the current 32-edge chunk) ----
        int nEdges = 0;
        long long lo = 0;
        int nGenes = 0;
        bool act = false;
        if (p < a.pop && a.active[p]) {
            lo = a.offsets[p];
            nGenes = (int)(a.offsets[p + 1] - lo);
            int m = nGenes == 0 ? 0 : nGenes + (a.hasStart ? 1 : 0) + (a.loops ? 1 : 0);
            nEdges = m > 0 ? m - 1 : 0;
            act = true;
        }
        auto boxAt = [&](int k) -> int {  // position k of the plan (N = start location)
            if (a.hasStart) {
                if (k == 0) return a.N;
                k -= 1;
            }
            return k < nGenes ? a.ids[lo + k] : a.ids[lo];
        };
        auto rowOfEdge = [&](int e) -> int {  // edge e joins positions e and e + 1
            if (e >= nEdges) return -1;
            return a.rowmap[boxAt(e) * a.N + boxAt(e + 1)];
        };
        // an area tile is needed by the CTA if any of its plans is active
        int anyAct = __syncthreads_or(act ? 1 : 0);
        if (!anyAct) continue;

        AreaAcc acc;
        acc.clear();
        // prefetch area tile 0
        {
            const double* src = a.area;
            for (int k = threadIdx.x; k < kTileTris / 2; k += kThreads) cp_async16(&s_area[0][2 * k], src + 2 * k);
            cp_async_commit();
        }
        for (int tile = 0; tile < a.nTiles; ++tile) {
            if (tile + 1 < a.nTiles) {
                const double* src = a.area + (size_t)(tile + 1) * kTileTris;
                for (int k = threadIdx.x; k < kTileTris / 2; k += kThreads) cp_async16(&s_area[(tile + 1) & 1][2 * k], src + 2 * k);
            }
            cp_async_commit();
            // ---- OR of the E rows for this tile ----
            uint4 m = make_uint4(0, 0, 0, 0);
            int vec = tile * (kTileWords / 4) + lane;  // uint4 index within the row
            const bool inb = vec < rowVec;
            if (act) {  // warp-uniform
                for (int e0 = 0; e0 < nEdges; e0 += 32) {
                    int myRow = rowOfEdge(e0 + lane);
                    int cnt = min(32, nEdges - e0);
                    int e = 0;
                    for (; e + 4 <= cnt; e += 4) {
                        int r0 = __shfl_sync(0xffffffffu, myRow, e), r1 = __shfl_sync(0xffffffffu, myRow, e + 1),
                            r2 = __shfl_sync(0xffffffffu, myRow, e + 2), r3 = __shfl_sync(0xffffffffu, myRow, e + 3);
                        uint4 z = make_uint4(0, 0, 0, 0);
                        uint4 v0 = (inb && r0 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r0 * a.rowWords) + vec) : z;
                        uint4 v1 = (inb && r1 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r1 * a.rowWords) + vec) : z;
                        uint4 v2 = (inb && r2 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r2 * a.rowWords) + vec) : z;
                        uint4 v3 = (inb && r3 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r3 * a.rowWords) + vec) : z;
                        m.x |= v0.x | v1.x | v2.x | v3.x;
                        m.y |= v0.y | v1.y | v2.y | v3.y;
                        m.z |= v0.z | v1.z | v2.z | v3.z;
                        m.w |= v0.w | v1.w | v2.w | v3.w;
                    }
                    for (; e < cnt; ++e) {
                        int r = __shfl_sync(0xffffffffu, myRow, e);
                        if (inb && r >= 0) {
                            uint4 v = ldg_stream((const uint4*)(a.rows + (size_t)r * a.rowWords) + vec);
                            m.x |= v.x; m.y |= v.y; m.z |= v.z; m.w |= v.w;
                        }
                    }
                }
            }
            // ---- area tile of this step is ready ----
            cp_async_wait<1>();
            __syncthreads();
            add_tile_area(acc, m, s_area[tile & 1], lane);
            __syncthreads();  // everyone is done with this buffer before it is refilled two steps later
        }
        cp_async_wait<0>();
        const double covered = combine_area(acc);
        if (act && lane == 0) a.fitness[4 * p] = 1.0 - (covered / a.totalArea);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The Blackwell form of the streaming kernel: bitmap rows move from HBM into shared memory with 1-D bulk asynchronous copies
// (cp.async.bulk, the TMA unit; SASS UBLKCP) that complete on mbarriers, so the bytes a warp has in flight no longer depend on
// how many registers its loads pin or on how many warps an SM holds.  Per plan (warp): a ring of S stages, a stage = the same
// 128 * SUB-word chunk of up to R of the plan's rows (R * SUB * 512 B); the lanes issue one bulk copy each for step k + S - 1,
// then wait for step k, OR its rows into registers with conflict-free 16-byte shared-memory reads, and after the last row group
// of a chunk add the areas of the set bits as before.  The 32 KB area tiles are bulk copies too (one instruction per tile and
// CTA, double buffered).  Rows of plans longer than kBulkMaxEdges edges are read with plain loads (same sums).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBulkMaxEdges = 64;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smemDst, const void* gmemSrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}

template <int PLANS, int R, int S, int SUB>
struct BulkLayout {
    static constexpr int kChunkWords = kTileWords * SUB;
    static constexpr size_t kAreaBytes = 2 * (size_t)kTileTris * sizeof(double);
    static constexpr size_t kRingBytes = (size_t)PLANS * S * R * kChunkWords * 4;
    static constexpr size_t kRowIdxBytes = (size_t)PLANS * kBulkMaxEdges * sizeof(int);
    static constexpr size_t kBarBytes = (size_t)(2 + PLANS * S) * sizeof(unsigned long long);
    static constexpr size_t kBytes = kAreaBytes + kRingBytes + kRowIdxBytes + kBarBytes;
};

template <int PLANS, int R, int S, int SUB>
__global__ void __launch_bounds__(PLANS * 32, 1) k_plan_reduce_bulk(PlanReduceArgs a) {
    using LY = BulkLayout<PLANS, R, S, SUB>;
    constexpr int kChunkWords = LY::kChunkWords;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double(*s_area)[kTileTris] = (double(*)[kTileTris])smemRaw;
    uint32_t* ringAll = reinterpret_cast<uint32_t*>(smemRaw + LY::kAreaBytes);
    int* rowIdxAll = reinterpret_cast<int*>(smemRaw + LY::kAreaBytes + LY::kRingBytes);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smemRaw + LY::kAreaBytes + LY::kRingBytes + LY::kRowIdxBytes);
    unsigned long long* areaFull = bars;                 // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* full = bars + 2 + warp * S;      // [S] of this warp
    uint32_t* ring = ringAll + (size_t)warp * S * R * kChunkWords;
    int* rowIdx = rowIdxAll + warp * kBulkMaxEdges;
    if (threadIdx.x == 0) {
        mbar_init(areaFull, 1);
        mbar_init(areaFull + 1, 1);
    }
    if (lane == 0)
        for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const long long rowWords = a.rowWords;
    // chunks of nb <= SUB tiles, balanced (18 tiles with SUB = 16: two chunks of 9 rather than 16 + 2)
    const int nSuper = (a.nTiles + SUB - 1) / SUB;
    const int nb = (a.nTiles + nSuper - 1) / nSuper;
    const int chunkWords = nb * kTileWords;
    unsigned stepCount = 0;   // row steps this warp has consumed since the kernel started (stage = count % S, parity = (count / S) & 1)
    unsigned tileCount = 0;   // area tiles this CTA has consumed since the kernel started
    for (long long p0 = (long long)blockIdx.x * PLANS; p0 < a.pop; p0 += (long long)gridDim.x * PLANS) {
        const long long p = p0 + warp;
        int nEdges = 0, nGenes = 0;
        long long lo = 0;
        bool act = false;
        if (p < a.pop && a.active[p]) {
            lo = a.offsets[p];
            nGenes = (int)(a.offsets[p + 1] - lo);
            const int m = nGenes == 0 ? 0 : nGenes + (a.hasStart ? 1 : 0) + (a.loops ? 1 : 0);
            nEdges = m > 0 ? m - 1 : 0;
            act = true;
        }
        auto boxAt = [&](int k) -> int {  // position k of the plan (N = start location)
            if (a.hasStart) {
                if (k == 0) return a.N;
                k -= 1;
            }
            return k < nGenes ? a.ids[lo + k] : a.ids[lo];
        };
        auto rowOfEdge = [&](int e) -> int {  // edge e joins positions e and e + 1
            if (e >= nEdges) return -1;
            return a.rowmap[boxAt(e) * a.N + boxAt(e + 1)];
        };
        const int anyAct = __syncthreads_or(act ? 1 : 0);
        if (!anyAct) continue;
        const bool bulk = act && nEdges <= kBulkMaxEdges;
        if (bulk)
            for (int e = lane; e < nEdges; e += 32) rowIdx[e] = rowOfEdge(e);
        __syncwarp();
        const int nGroups = bulk ? (nEdges + R - 1) / R : 0;
        const unsigned totalSteps = (unsigned)(nSuper * nGroups);
        // issue the row step `k` of this plan (0-based within the plan): chunk k / nGroups of the rows of group k % nGroups
        auto issue = [&](unsigned k, unsigned slotCount) {
            const int sup = (int)(k / (unsigned)nGroups), g = (int)(k % (unsigned)nGroups);
            const int stage = (int)(slotCount % S);
            const long long w0 = (long long)sup * chunkWords;
            const unsigned bytes = (unsigned)(min((long long)chunkWords, rowWords - w0) * 4);
            const int e = g * R + lane;
            const int row = (lane < R && e < nEdges) ? rowIdx[e] : -1;
            const unsigned valid = __ballot_sync(0xffffffffu, row >= 0);
            if (lane == 0) mbar_expect_tx(full + stage, bytes * (unsigned)__popc(valid));
            __syncwarp();
            if (row >= 0) bulk_g2s(ring + ((size_t)stage * R + lane) * kChunkWords, a.rows + (size_t)row * rowWords + w0, bytes, full + stage);
        };
        unsigned issued = 0;
        for (; issued < totalSteps && issued < (unsigned)(S - 1); ++issued) issue(issued, stepCount + issued);
        // area tiles 0 and 1
        if (threadIdx.x == 0) {
            for (int t = 0; t < 2 && t < a.nTiles; ++t) {
                unsigned long long* bar = areaFull + ((tileCount + t) & 1u);
                mbar_expect_tx(bar, (unsigned)(kTileTris * sizeof(double)));
                bulk_g2s(&s_area[(tileCount + t) & 1u][0], a.area + (size_t)t * kTileTris, (unsigned)(kTileTris * sizeof(double)), bar);
            }
        }

        AreaAcc acc;
        acc.clear();
        unsigned consumed = 0;
        for (int sup = 0; sup < nSuper; ++sup) {
            uint4 m[SUB];
#pragma unroll
            for (int u = 0; u < SUB; ++u) m[u] = make_uint4(0, 0, 0, 0);
            const long long w0 = (long long)sup * chunkWords;
            if (bulk) {
                for (int g = 0; g < nGroups; ++g) {
                    if (issued < totalSteps) {
                        issue(issued, stepCount + issued);
                        ++issued;
                    }
                    const unsigned slot = stepCount + consumed;
                    mbar_wait(full + (slot % S), (slot / S) & 1u);
                    const uint32_t* st = ring + (size_t)(slot % S) * R * kChunkWords;
                    const int nrows = min(R, nEdges - g * R);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r < nrows && rowIdx[g * R + r] >= 0) {  // warp-uniform
#pragma unroll
                            for (int u = 0; u < SUB; ++u) {
                                if (u < nb && w0 + u * kTileWords + lane * 4 < rowWords) {
                                    const uint4 v = lds128(st + (size_t)r * kChunkWords + u * kTileWords + lane * 4);
                                    m[u].x |= v.x; m[u].y |= v.y; m[u].z |= v.z; m[u].w |= v.w;
                                }
                            }
                        }
                    }
                    ++consumed;
                    __syncwarp();  // every lane has read the stage before a later step's copies land in it
                }
            } else if (act) {  // a plan of more than kBulkMaxEdges edges: plain loads
                for (int e = 0; e < nEdges; ++e) {
                    const int r = rowOfEdge(e);
                    if (r < 0) continue;
#pragma unroll
                    for (int u = 0; u < SUB; ++u) {
                        const long long w = w0 + u * kTileWords + lane * 4;
                        if (u < nb && w < rowWords) {
                            const uint4 v = ldg_stream((const uint4*)(a.rows + (size_t)r * rowWords + w));
                            m[u].x |= v.x; m[u].y |= v.y; m[u].z |= v.z; m[u].w |= v.w;
                        }
                    }
                }
            }
            // ---- areas of the set bits, tile by tile ----
#pragma unroll
            for (int u = 0; u < SUB; ++u) {
                const int tile = sup * nb + u;
                if (u >= nb || tile >= a.nTiles) break;
                const unsigned tslot = tileCount + (unsigned)tile;
                mbar_wait(areaFull + (tslot & 1u), (tslot >> 1) & 1u);
                add_tile_area(acc, m[u], s_area[tslot & 1u], lane);
                __syncthreads();  // everyone is done with this buffer: tile + 2 may land in it
                if (threadIdx.x == 0 && tile + 2 < a.nTiles) {
                    unsigned long long* bar = areaFull + (tslot & 1u);
                    mbar_expect_tx(bar, (unsigned)(kTileTris * sizeof(double)));
                    bulk_g2s(&s_area[tslot & 1u][0], a.area + (size_t)(tile + 2) * kTileTris, (unsigned)(kTileTris * sizeof(double)), bar);
                }
            }
        }
        stepCount += totalSteps;
        tileCount += (unsigned)a.nTiles;
        const double covered = combine_area(acc);
        if (act && lane == 0) a.fitness[4 * p] = 1.0 - (covered / a.totalArea);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-major form of the streaming kernel: a warp takes the rows of its plan one after the other and reads a long contiguous piece
// of each (nb <= NB tiles = up to NB * 512 B) with NB independent 16-byte loads per lane in flight, ORing into NB uint4 registers;
// the areas of the set bits are added afterwards, tile by tile, from bulk-copied area tiles as in k_plan_reduce_bulk.  Long
// contiguous reads instead of 512-byte pieces of E rows interleaved.
// ---------------------------------------------------------------------------------------------------------------------
template <int PLANS, int NB, int EU, int MINB>
__global__ void __launch_bounds__(PLANS * 32, MINB) k_plan_reduce_rows(PlanReduceArgs a) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double(*s_area)[kTileTris] = (double(*)[kTileTris])smemRaw;
    unsigned long long* areaFull = reinterpret_cast<unsigned long long*>(smemRaw + 2 * (size_t)kTileTris * sizeof(double));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(areaFull, 1);
        mbar_init(areaFull + 1, 1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long rowWords = a.rowWords;
    const int nSuper = (a.nTiles + NB - 1) / NB;
    const int nb = (a.nTiles + nSuper - 1) / nSuper;
    unsigned tileCount = 0;
    for (long long p0 = (long long)blockIdx.x * PLANS; p0 < a.pop; p0 += (long long)gridDim.x * PLANS) {
        const long long p = p0 + warp;
        int nEdges = 0, nGenes = 0;
        long long lo = 0;
        bool act = false;
        if (p < a.pop && a.active[p]) {
            lo = a.offsets[p];
            nGenes = (int)(a.offsets[p + 1] - lo);
            const int m = nGenes == 0 ? 0 : nGenes + (a.hasStart ? 1 : 0) + (a.loops ? 1 : 0);
            nEdges = m > 0 ? m - 1 : 0;
            act = true;
        }
        auto boxAt = [&](int k) -> int {
            if (a.hasStart) {
                if (k == 0) return a.N;
                k -= 1;
            }
            return k < nGenes ? a.ids[lo + k] : a.ids[lo];
        };
        auto rowOfEdge = [&](int e) -> int {
            if (e >= nEdges) return -1;
            return a.rowmap[boxAt(e) * a.N + boxAt(e + 1)];
        };
        const int anyAct = __syncthreads_or(act ? 1 : 0);
        if (!anyAct) continue;
        if (threadIdx.x == 0) {
            for (int t = 0; t < 2 && t < a.nTiles; ++t) {
                unsigned long long* bar = areaFull + ((tileCount + t) & 1u);
                mbar_expect_tx(bar, (unsigned)(kTileTris * sizeof(double)));
                bulk_g2s(&s_area[(tileCount + t) & 1u][0], a.area + (size_t)t * kTileTris, (unsigned)(kTileTris * sizeof(double)), bar);
            }
        }
        AreaAcc acc;
        acc.clear();
        for (int sup = 0; sup < nSuper; ++sup) {
            uint4 m[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) m[u] = make_uint4(0, 0, 0, 0);
            const long long w0 = (long long)sup * nb * kTileWords + lane * 4;
            if (act) {
                for (int e0 = 0; e0 < nEdges; e0 += 32) {
                    const int myRow = rowOfEdge(e0 + lane);
                    const int cnt = min(32, nEdges - e0);
                    for (int e = 0; e < cnt; e += EU) {
                        uint4 v[EU][NB];
#pragma unroll
                        for (int q = 0; q < EU; ++q) {
                            const int r = __shfl_sync(0xffffffffu, myRow, min(e + q, 31));
                            const uint32_t* row = a.rows + (size_t)max(r, 0) * rowWords + w0;
                            const bool rowOk = e + q < cnt && r >= 0;
#pragma unroll
                            for (int u = 0; u < NB; ++u)
                                v[q][u] = (rowOk && u < nb && w0 + u * kTileWords < rowWords) ? ldg_stream((const uint4*)(row + u * kTileWords))
                                                                                              : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int q = 0; q < EU; ++q)
#pragma unroll
                            for (int u = 0; u < NB; ++u) {
                                m[u].x |= v[q][u].x; m[u].y |= v[q][u].y; m[u].z |= v[q][u].z; m[u].w |= v[q][u].w;
                            }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int tile = sup * nb + u;
                if (u >= nb || tile >= a.nTiles) break;
                const unsigned tslot = tileCount + (unsigned)tile;
                mbar_wait(areaFull + (tslot & 1u), (tslot >> 1) & 1u);
                add_tile_area(acc, m[u], s_area[tslot & 1u], lane);
                __syncthreads();
                if (threadIdx.x == 0 && tile + 2 < a.nTiles) {
                    unsigned long long* bar = areaFull + (tslot & 1u);
                    mbar_expect_tx(bar, (unsigned)(kTileTris * sizeof(double)));
                    bulk_g2s(&s_area[tslot & 1u][0], a.area + (size_t)(tile + 2) * kTileTris, (unsigned)(kTileTris * sizeof(double)), bar);
                }
            }
        }
        tileCount += (unsigned)a.nTiles;
        const double covered = combine_area(acc);
        if (act && lane == 0) a.fitness[4 * p] = 1.0 - (covered / a.totalArea);
    }
}

// The row-major kernel with the areas read straight from global memory (L1 / L2 resident: T doubles) instead of bulk-copied area
// tiles in shared memory: no barrier couples the warps of a CTA, so a warp that waits for a row does not hold the others back.
template <int PLANS, int NB, int MINB>
__global__ void __launch_bounds__(PLANS * 32, MINB) k_plan_reduce_rows_g(PlanReduceArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rowWords = a.rowWords;
    const int nSuper = (a.nTiles + NB - 1) / NB;
    const int nb = (a.nTiles + nSuper - 1) / nSuper;
    for (long long p = (long long)blockIdx.x * PLANS + warp; p < a.pop; p += (long long)gridDim.x * PLANS) {
        if (!a.active[p]) continue;
        const long long lo = a.offsets[p];
        const int nGenes = (int)(a.offsets[p + 1] - lo);
        const int mm = nGenes == 0 ? 0 : nGenes + (a.hasStart ? 1 : 0) + (a.loops ? 1 : 0);
        const int nEdges = mm > 0 ? mm - 1 : 0;
        auto boxAt = [&](int k) -> int {
            if (a.hasStart) {
                if (k == 0) return a.N;
                k -= 1;
            }
            return k < nGenes ? a.ids[lo + k] : a.ids[lo];
        };
        auto rowOfEdge = [&](int e) -> int {
            if (e >= nEdges) return -1;
            return a.rowmap[boxAt(e) * a.N + boxAt(e + 1)];
        };
        AreaAcc acc;
        acc.clear();
        for (int sup = 0; sup < nSuper; ++sup) {
            uint4 m[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) m[u] = make_uint4(0, 0, 0, 0);
            const long long w0 = (long long)sup * nb * kTileWords + lane * 4;
            for (int e0 = 0; e0 < nEdges; e0 += 32) {
                const int myRow = rowOfEdge(e0 + lane);
                const int cnt = min(32, nEdges - e0);
                for (int e = 0; e < cnt; ++e) {
                    const int r = __shfl_sync(0xffffffffu, myRow, e);
                    if (r < 0) continue;
                    const uint32_t* row = a.rows + (size_t)r * rowWords + w0;
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        if (u < nb && w0 + u * kTileWords < rowWords) {
                            const uint4 v = ldg_stream((const uint4*)(row + u * kTileWords));
                            m[u].x |= v.x; m[u].y |= v.y; m[u].z |= v.z; m[u].w |= v.w;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int tile = sup * nb + u;
                if (u >= nb || tile >= a.nTiles) break;
                add_tile_area(acc, m[u], a.area + (size_t)tile * kTileTris, lane);
            }
        }
        const double covered = combine_area(acc);
        if (lane == 0) a.fitness[4 * p] = 1.0 - (covered / a.totalArea);
    }
}

// The same reduce with two options.  kRows: rows come from the explicit per-edge list ([id, angle] genes, heading search) instead
// of rowmap[key]; kSplit: the tiles of a plan are spread over gridDim.y CTAs in blocks of kBlockTiles (small populations on large
// meshes), per-block area sums go to a.partial and k_plan_finish adds them up in block order.
template <bool kRows, bool kSplit>
__global__ void __launch_bounds__(kThreads) k_plan_reduce_ext(PlanReduceArgs a) {
    extern __shared__ __align__(16) double s_area_raw[];  // 2 x 32 KB
    double(*s_area)[kTileTris] = (double(*)[kTileTris])s_area_raw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowVec = (int)(a.rowWords / 4);  // uint4 per row
    for (long long p0 = (long long)blockIdx.x * kPlansPerCta; p0 < a.pop; p0 += (long long)gridDim.x * kPlansPerCta) {
        long long p = p0 + warp;
        // ---- this warp's plan: edge list -> row indices (lane e holds edge e of the current 32-edge chunk) ----
        int nEdges = 0;
        long long lo = 0;
        int nGenes = 0;
        bool act = false;
        if (p < a.pop && a.active[p]) {
            lo = a.offsets[p];
            nGenes = (int)(a.offsets[p + 1] - lo);
            int m = nGenes == 0 ? 0 : nGenes + (a.hasStart ? 1 : 0) + (a.loops ? 1 : 0);
            nEdges = m > 0 ? m - 1 : 0;
            act = true;
        }
        auto boxAt = [&](int k) -> int {  // position k of the plan (N = start location)
            if (a.hasStart) {
                if (k == 0) return a.N;
                k -= 1;
            }
            return k < nGenes ? a.ids[lo + k] : a.ids[lo];
        };
        auto rowOfEdge = [&](int e) -> int {  // edge e joins positions e and e + 1
            if (e >= nEdges) return -1;
            if (kRows) return a.edgeRows[a.edgeOffsets[p] + e];
            return a.rowmap[boxAt(e) * a.N + boxAt(e + 1)];
        };
        // an area tile is needed by the CTA if any of its plans is active
        int anyAct = __syncthreads_or(act ? 1 : 0);
        if (!anyAct) continue;

        // this CTA's share of the tiles, in whole blocks of kBlockTiles tiles (gridDim.y splits; 1 = all of them)
        const int nBlocks = (a.nTiles + kBlockTiles - 1) / kBlockTiles;
        const int tileLo = kSplit ? (int)(((long long)blockIdx.y * nBlocks) / gridDim.y) * kBlockTiles : 0;
        const int tileHi = kSplit ? min(a.nTiles, (int)(((long long)(blockIdx.y + 1) * nBlocks) / gridDim.y) * kBlockTiles) : a.nTiles;
        AreaAcc acc;  // this lane's share of the covered area (kSplit: of the current block of tiles)
        acc.clear();
        // prefetch the first area tile
        if (tileLo < tileHi) {
            const double* src = a.area + (size_t)tileLo * kTileTris;
            for (int k = threadIdx.x; k < kTileTris / 2; k += kThreads) cp_async16(&s_area[tileLo & 1][2 * k], src + 2 * k);
            cp_async_commit();
        }
        for (int tile = tileLo; tile < tileHi; ++tile) {
            if (tile + 1 < tileHi) {
                const double* src = a.area + (size_t)(tile + 1) * kTileTris;
                for (int k = threadIdx.x; k < kTileTris / 2; k += kThreads) cp_async16(&s_area[(tile + 1) & 1][2 * k], src + 2 * k);
            }
            cp_async_commit();
            // ---- OR of the E rows for this tile ----
            uint4 m = make_uint4(0, 0, 0, 0);
            int vec = tile * (kTileWords / 4) + lane;  // uint4 index within the row
            const bool inb = vec < rowVec;
            if (act) {  // warp-uniform
                for (int e0 = 0; e0 < nEdges; e0 += 32) {
                    int myRow = rowOfEdge(e0 + lane);
                    int cnt = min(32, nEdges - e0);
                    int e = 0;
                    for (; e + 4 <= cnt; e += 4) {
                        int r0 = __shfl_sync(0xffffffffu, myRow, e), r1 = __shfl_sync(0xffffffffu, myRow, e + 1),
                            r2 = __shfl_sync(0xffffffffu, myRow, e + 2), r3 = __shfl_sync(0xffffffffu, myRow, e + 3);
                        uint4 z = make_uint4(0, 0, 0, 0);
                        uint4 v0 = (inb && r0 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r0 * a.rowWords) + vec) : z;
                        uint4 v1 = (inb && r1 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r1 * a.rowWords) + vec) : z;
                        uint4 v2 = (inb && r2 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r2 * a.rowWords) + vec) : z;
                        uint4 v3 = (inb && r3 >= 0) ? ldg_stream((const uint4*)(a.rows + (size_t)r3 * a.rowWords) + vec) : z;
                        m.x |= v0.x | v1.x | v2.x | v3.x;
                        m.y |= v0.y | v1.y | v2.y | v3.y;
                        m.z |= v0.z | v1.z | v2.z | v3.z;
                        m.w |= v0.w | v1.w | v2.w | v3.w;
                    }
                    for (; e < cnt; ++e) {
                        int r = __shfl_sync(0xffffffffu, myRow, e);
                        if (inb && r >= 0) {
                            uint4 v = ldg_stream((const uint4*)(a.rows + (size_t)r * a.rowWords) + vec);
                            m.x |= v.x; m.y |= v.y; m.z |= v.z; m.w |= v.w;
                        }
                    }
                }
            }
            // ---- area tile of this step is ready ----
            cp_async_wait<1>();
            __syncthreads();
            add_tile_area(acc, m, s_area[tile & 1], lane);
            if (kSplit && ((tile + 1) % kBlockTiles == 0 || tile + 1 == a.nTiles)) {
                // covered area of this block of tiles
                const double blockSum = combine_area(acc);
                if (act && lane == 0) a.partial[(size_t)p * nBlocks + tile / kBlockTiles] = blockSum;
                acc.clear();
            }
            __syncthreads();  // everyone is done with this buffer before it is refilled two steps later
        }
        cp_async_wait<0>();
        if (!kSplit) {  // one reduce over all tiles, the summation order of the streaming kernel
            const double covered = combine_area(acc);
            if (act && lane == 0) a.fitness[4 * p] = 1.0 - (covered / a.totalArea);
        }
    }
}

// second step of a split reduce: per-block sums in block order -> coverage score
__global__ void k_plan_finish(PlanReduceArgs a) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= a.pop || !a.active[p]) return;
    const int nBlocks = (a.nTiles + kBlockTiles - 1) / kBlockTiles;
    double total = 0.0;
    for (int t = 0; t < nBlocks; ++t) total += a.partial[(size_t)p * nBlocks + t];
    a.fitness[4 * p] = 1.0 - (total / a.totalArea);
}

// OR of all rows whose key is marked used -> union bitmap (thread per word)
__global__ void k_union_rows(const uint8_t* __restrict__ used, int nKeys, const int* __restrict__ rowmap, const uint32_t* __restrict__ rows,
                             long long rowWords, uint32_t* __restrict__ out) {
    long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= rowWords) return;
    uint32_t acc = 0;
    for (int k = 0; k < nKeys; ++k) {
        if (!used[k]) continue;
        int r = rowmap[k];
        if (r >= 0) acc |= rows[(size_t)r * rowWords + w];
    }
    out[w] |= acc;  // accumulated: a population evaluated in several pieces adds the union of each piece
}

__global__ void k_union_row_list(const int* __restrict__ rowList, int n, const uint32_t* __restrict__ rows, long long rowWords,
                                 uint32_t* __restrict__ out) {
    long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (w >= rowWords) return;
    uint32_t acc = 0;
    for (int k = 0; k < n; ++k) acc |= rows[(size_t)rowList[k] * rowWords + w];
    out[w] |= acc;  // accumulated, like k_union_rows
}
__global__ void k_release_rows(const int* __restrict__ rowList, int n, int* __restrict__ freeStack, int* __restrict__ freeTop) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) freeStack[atomicAdd(freeTop, 1)] = rowList[i];
}

__global__ void k_fill_u8(uint8_t* p, long long n, uint8_t v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_i32(int* p, long long n, int v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_iota_i32(int* p, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)(n - 1 - i);  // stack order: row 0 is popped first
}
// read-bandwidth probe: every thread sums 16-byte loads over the buffer, `reps` passes (the sums keep the loads alive)
__global__ void k_read_probe(const uint4* __restrict__ p, long long n16, int reps, unsigned* __restrict__ sink) {
    unsigned acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void k_flush(uint32_t* p, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = (uint32_t)i;
}

}  // namespace

namespace {
template <int PLANS>
void launch_stream(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    // per call: the attributes belong to the current device's context
    cudaFuncSetAttribute(k_plan_reduce<PLANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(k_plan_reduce<PLANS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static int perSM = 0;
    if (!perSM) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_plan_reduce<PLANS>, PLANS * 32, kSmemBytes);
        if (perSM < 1) perSM = 1;
    }
    const long long groups = (a.pop + PLANS - 1) / PLANS;
    int grid = (int)std::min<long long>((long long)numSMs * perSM, groups);
    k_plan_reduce<PLANS><<<grid, PLANS * 32, kSmemBytes, st>>>(a);
}

template <bool kRows, bool kSplit>
void launch_plan_reduce_t(const PlanReduceArgs& a, int numSMs, int split, cudaStream_t st) {
    // per call: the attributes belong to the current device's context
    cudaFuncSetAttribute(k_plan_reduce_ext<kRows, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(k_plan_reduce_ext<kRows, kSplit>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static int perSM = 0;
    if (!perSM) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_plan_reduce_ext<kRows, kSplit>, kThreads, kSmemBytes);
        if (perSM < 1) perSM = 1;
    }
    long long groups = (a.pop + kPlansPerCta - 1) / kPlansPerCta;
    int grid = (int)std::min<long long>((long long)numSMs * perSM, groups);
    k_plan_reduce_ext<kRows, kSplit><<<dim3(grid, split), kThreads, kSmemBytes, st>>>(a);
}
}  // namespace

namespace {
template <int PLANS, int R, int S, int SUB>
void launch_bulk(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    using LY = BulkLayout<PLANS, R, S, SUB>;
    static_assert(LY::kBytes <= 227 * 1024, "the bulk-copy ring does not fit one SM");
    cudaFuncSetAttribute(k_plan_reduce_bulk<PLANS, R, S, SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::kBytes);
    const long long groups = (a.pop + PLANS - 1) / PLANS;
    const int grid = (int)std::min<long long>(numSMs, groups);
    k_plan_reduce_bulk<PLANS, R, S, SUB><<<grid, PLANS * 32, LY::kBytes, st>>>(a);
}
// K6 variant: 0 = plain loads (k_plan_reduce), 1.. = bulk-copy configurations (IPP_REDUCE, for A/B runs; the default is chosen in
// launch_plan_reduce)
int g_reduceVariant = -2;
int reduce_variant() {
    if (g_reduceVariant == -2) {
        const char* e = getenv("IPP_REDUCE");
        g_reduceVariant = e ? atoi(e) : -1;
    }
    return g_reduceVariant;
}
}  // namespace

void set_plan_reduce_variant(int v) { g_reduceVariant = v; }

namespace {
template <int PLANS, int NB, int EU, int MINB>
void launch_rows(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    const size_t smem = 2 * (size_t)kTileTris * sizeof(double) + 2 * sizeof(unsigned long long);
    cudaFuncSetAttribute(k_plan_reduce_rows<PLANS, NB, EU, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_plan_reduce_rows<PLANS, NB, EU, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const long long groups = (a.pop + PLANS - 1) / PLANS;
    const int grid = (int)std::min<long long>((long long)numSMs * MINB, groups);
    k_plan_reduce_rows<PLANS, NB, EU, MINB><<<grid, PLANS * 32, smem, st>>>(a);
}
template <int PLANS, int NB, int MINB>
void launch_rows_g(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    const long long groups = (a.pop + PLANS - 1) / PLANS;
    const int grid = (int)std::min<long long>((long long)numSMs * MINB, groups);
    k_plan_reduce_rows_g<PLANS, NB, MINB><<<grid, PLANS * 32, 0, st>>>(a);
}
}  // namespace

const char* plan_reduce_variant_name(int v) {
    switch (v) {
        case 0: return "ldg tile-major";
        case 1: return "bulk PLANS=8 R=8 S=4 chunk<=512B";
        case 2: return "bulk PLANS=8 R=2 S=4 chunk<=2KB";
        case 3: return "bulk PLANS=8 R=1 S=4 chunk<=4KB";
        case 4: return "bulk PLANS=6 R=1 S=3 chunk<=8KB";
        case 5: return "bulk PLANS=4 R=1 S=5 chunk<=8KB";
        case 6: return "ldg row-major PLANS=8 NB=16 EU=1, 1 CTA/SM";
        case 7: return "ldg row-major PLANS=8 NB=8 EU=1, 2 CTA/SM";
        case 8: return "ldg row-major PLANS=12 NB=9 EU=1, 1 CTA/SM";
        case 9: return "ldg row-major PLANS=8 NB=9 EU=1, 2 CTA/SM";
        case 10: return "ldg row-major PLANS=16 NB=9 EU=1, 1 CTA/SM";
        case 11: return "ldg row-major PLANS=8 NB=4 EU=2, 3 CTA/SM";
        case 12: return "ldg row-major PLANS=12 NB=4 EU=1, 3 CTA/SM";
        case 13: return "ldg row-major, areas from global, PLANS=12 NB=4, 3 CTA/SM";
        case 14: return "ldg row-major, areas from global, PLANS=8 NB=4, 5 CTA/SM";
        case 15: return "ldg row-major, areas from global, PLANS=8 NB=4, 6 CTA/SM";
        case 16: return "ldg row-major, areas from global, PLANS=8 NB=8, 4 CTA/SM";
        case 17: return "ldg row-major, areas from global, PLANS=4 NB=4, 10 CTA/SM";
        case 18: return "ldg row-major, areas from global, PLANS=8 NB=4, 4 CTA/SM";
        case 19: return "ldg row-major, areas from global, PLANS=10 NB=4, 4 CTA/SM";
        case 20: return "ldg row-major, areas from global, PLANS=8 NB=2, 5 CTA/SM";
        case 21: return "ldg row-major, areas from global, PLANS=16 NB=4, 2 CTA/SM";
        case 22: return "ldg row-major, areas from global, PLANS=8 NB=3, 5 CTA/SM";
        default: return "auto";
    }
}

void launch_plan_reduce(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    if (a.pop <= 0) return;
    if (!a.edgeRows && !a.partial && reduce_variant() > 0) {
        switch (reduce_variant()) {
            case 1: launch_bulk<8, 8, 4, 1>(a, numSMs, st); return;
            case 2: launch_bulk<8, 2, 4, 4>(a, numSMs, st); return;
            case 3: launch_bulk<8, 1, 4, 8>(a, numSMs, st); return;
            case 4: launch_bulk<6, 1, 3, 16>(a, numSMs, st); return;
            case 5: launch_bulk<4, 1, 5, 16>(a, numSMs, st); return;
            case 6: launch_rows<8, 16, 1, 1>(a, numSMs, st); return;
            case 7: launch_rows<8, 8, 1, 2>(a, numSMs, st); return;
            case 8: launch_rows<12, 9, 1, 1>(a, numSMs, st); return;
            case 9: launch_rows<8, 9, 1, 2>(a, numSMs, st); return;
            case 10: launch_rows<16, 9, 1, 1>(a, numSMs, st); return;
            case 11: launch_rows<8, 4, 2, 3>(a, numSMs, st); return;
            case 12: launch_rows<12, 4, 1, 3>(a, numSMs, st); return;
            case 13: launch_rows_g<12, 4, 3>(a, numSMs, st); return;
            case 14: launch_rows_g<8, 4, 5>(a, numSMs, st); return;
            case 15: launch_rows_g<8, 4, 6>(a, numSMs, st); return;
            case 16: launch_rows_g<8, 8, 4>(a, numSMs, st); return;
            case 17: launch_rows_g<4, 4, 10>(a, numSMs, st); return;
            case 18: launch_rows_g<8, 4, 4>(a, numSMs, st); return;
            case 19: launch_rows_g<10, 4, 4>(a, numSMs, st); return;
            case 20: launch_rows_g<8, 2, 5>(a, numSMs, st); return;
            case 21: launch_rows_g<16, 4, 2>(a, numSMs, st); return;
            case 22: launch_rows_g<8, 3, 5>(a, numSMs, st); return;
            default: break;
        }
    }
    const long long groups = (a.pop + kPlansPerCta - 1) / kPlansPerCta;
    const int nBlocks = (a.nTiles + kBlockTiles - 1) / kBlockTiles;
    int split = 1;
    if (a.partial && nBlocks > 1 && groups < (long long)numSMs * 2)
        split = (int)std::max<long long>(1, std::min<long long>(nBlocks, ((long long)numSMs * 2 + groups - 1) / groups));
    const bool rows = a.edgeRows != nullptr;
    if (split > 1) {
        if (rows) launch_plan_reduce_t<true, true>(a, numSMs, split, st);
        else launch_plan_reduce_t<false, true>(a, numSMs, split, st);
        k_plan_finish<<<(unsigned)((a.pop + 127) / 128), 128, 0, st>>>(a);
    } else if (rows) {
        launch_plan_reduce_t<true, false>(a, numSMs, 1, st);
    } else if (a.pop >= (long long)numSMs * 3 * kPlansPerCtaWide) {
        // enough plans to fill the machine with 40 warps per SM: the row-major kernel that reads the areas from global memory (no
        // shared-memory area tiles, so no barrier couples the warps of a CTA and five CTAs of 8 plans fit an SM).  Measured at
        // 100 000 plans, L2 flushed (profiles/r2_k6_variants.md): luis4 96 % of the HBM copy peak against 72.7 % for the row-major
        // kernel with staged area tiles and 67 % for the tile-major kernel; mockup_only 68 % against 62.5 %.  Below a full wave
        // (3 000 plans on luis4, 1 000 on mockup_only) the tile-major kernel is 10-20 % faster and keeps those populations.
        // Every plan then reads the whole area array through L1 / L2 (8 T bytes, about as much as its rows): fine up to the 327 680
        // triangles of ico7 (80 tiles: 114 % against 89 %), a loss on the 1.3 M triangles of config 5 (320 tiles: 72 % against
        // 85 %), where the kernel that stages each area tile once per CTA keeps the job.
        if (a.nTiles <= 128) launch_rows_g<8, 4, 5>(a, numSMs, st);
        else launch_rows<12, 4, 1, 3>(a, numSMs, st);
    } else {
        launch_stream<kPlansPerCta>(a, numSMs, st);
    }
}

void launch_union_rows(const uint8_t* d_used, int nKeys, const int* d_rowmap, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_union,
                       cudaStream_t st) {
    k_union_rows<<<(unsigned)((rowWords + 127) / 128), 128, 0, st>>>(d_used, nKeys, d_rowmap, d_rows, (long long)rowWords, d_union);
}
void launch_union_row_list(const int* d_rowList, int n, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_union, cudaStream_t st) {
    k_union_row_list<<<(unsigned)((rowWords + 127) / 128), 128, 0, st>>>(d_rowList, n, d_rows, (long long)rowWords, d_union);
}
void launch_release_rows(const int* d_rowList, int n, int* d_freeStack, int* d_freeTop, cudaStream_t st) {
    if (n > 0) k_release_rows<<<(n + 255) / 256, 256, 0, st>>>(d_rowList, n, d_freeStack, d_freeTop);
}
void launch_fill_u8(uint8_t* p, int64_t n, uint8_t v, cudaStream_t st) {
    if (n > 0) k_fill_u8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, (long long)n, v);
}
void launch_fill_i32(int* p, int64_t n, int v, cudaStream_t st) {
    if (n > 0) k_fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, (long long)n, v);
}
void launch_iota_i32(int* p, int64_t n, cudaStream_t st) {
    if (n > 0) k_iota_i32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, (long long)n);
}
void launch_read_probe(const uint32_t* p, int64_t nWords, int reps, unsigned* d_sink, int numSMs, cudaStream_t st) {
    if (nWords >= 4) k_read_probe<<<numSMs * 8, 512, 0, st>>>((const uint4*)p, (long long)(nWords / 4), reps, d_sink);
}
void launch_flush(uint32_t* p, int64_t nWords, cudaStream_t st) {
    if (nWords > 0) k_flush<<<148 * 8, 256, 0, st>>>(p, (long long)nWords);
}

}  // namespace ipp
