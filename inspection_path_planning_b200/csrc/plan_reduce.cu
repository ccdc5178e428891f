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

        double acc = 0.0;
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
            const double* ar = s_area[tile & 1];
            unsigned nz = __ballot_sync(0xffffffffu, (m.x | m.y | m.z | m.w) != 0u);
            while (nz) {
                int q = __ffs(nz) - 1;
                nz &= nz - 1;
                unsigned w0 = __shfl_sync(0xffffffffu, m.x, q), w1 = __shfl_sync(0xffffffffu, m.y, q),
                         w2 = __shfl_sync(0xffffffffu, m.z, q), w3 = __shfl_sync(0xffffffffu, m.w, q);
                const double* base = ar + q * 128 + lane;
                if ((w0 >> lane) & 1u) acc += base[0];
                if ((w1 >> lane) & 1u) acc += base[32];
                if ((w2 >> lane) & 1u) acc += base[64];
                if ((w3 >> lane) & 1u) acc += base[96];
            }
            __syncthreads();  // everyone is done with this buffer before it is refilled two steps later
        }
        cp_async_wait<0>();
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (act && lane == 0) a.fitness[4 * p] = 1.0 - (acc / a.totalArea);
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
        double acc = 0.0;  // this lane's share of the covered area (kSplit: of the current block of tiles)
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
            const double* ar = s_area[tile & 1];
            unsigned nz = __ballot_sync(0xffffffffu, (m.x | m.y | m.z | m.w) != 0u);
            while (nz) {
                int q = __ffs(nz) - 1;
                nz &= nz - 1;
                unsigned w0 = __shfl_sync(0xffffffffu, m.x, q), w1 = __shfl_sync(0xffffffffu, m.y, q),
                         w2 = __shfl_sync(0xffffffffu, m.z, q), w3 = __shfl_sync(0xffffffffu, m.w, q);
                const double* base = ar + q * 128 + lane;
                if ((w0 >> lane) & 1u) acc += base[0];
                if ((w1 >> lane) & 1u) acc += base[32];
                if ((w2 >> lane) & 1u) acc += base[64];
                if ((w3 >> lane) & 1u) acc += base[96];
            }
            if (kSplit && ((tile + 1) % kBlockTiles == 0 || tile + 1 == a.nTiles)) {
                // covered area of this block of tiles (butterfly: every lane ends with the same sum)
                for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (act && lane == 0) a.partial[(size_t)p * nBlocks + tile / kBlockTiles] = acc;
                acc = 0.0;
            }
            __syncthreads();  // everyone is done with this buffer before it is refilled two steps later
        }
        cp_async_wait<0>();
        if (!kSplit) {  // one reduce over all tiles, the summation order of the streaming kernel
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (act && lane == 0) a.fitness[4 * p] = 1.0 - (acc / a.totalArea);
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

void launch_plan_reduce(const PlanReduceArgs& a, int numSMs, cudaStream_t st) {
    if (a.pop <= 0) return;
    // A population too small to fill the machine (fewer plan groups of 8 than two per SM) on a mesh of several tile blocks
    // spreads the tiles of its plans over several CTAs.  The split sums per block of tiles and adds the blocks up afterwards: the
    // last bit of a coverage score may then differ from what the same plan gets inside a large population.
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
        // enough plans for three CTAs of 12 warps on every SM: 36 warps per SM instead of 24 keep more row reads in flight
        // (+19 % on mockup_only, +9 % on luis4 at 100 000 plans; a smaller population prefers more, smaller CTAs)
        launch_stream<kPlansPerCtaWide>(a, numSMs, st);
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
void launch_flush(uint32_t* p, int64_t nWords, cudaStream_t st) {
    if (nWords > 0) k_flush<<<148 * 8, 256, 0, st>>>(p, (long long)nWords);
}

}  // namespace ipp
