// evaluator.cu -- host orchestration of the CUDA plan evaluator (see evaluator.h).
// Flow of one population (PlanCoverageEstimator::evaluatePlan, PlanCoverageEstimator.cpp:201-282, batched):
//   mark unknown collision flags -> K5 -> energy + gate -> mark unknown edge bitmaps of gated-in plans
//   -> K3 decode -> K4 visibility -> pack -> K6 plan_reduce.
#include "evaluator.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "exact.cuh"
#include "kernels.h"
#include "mesh_io.h"

namespace ipp {

namespace {
template <class T>
T* dalloc(size_t n) {
    T* p = nullptr;
    IPP_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
}  // namespace

Evaluator::Evaluator(const std::string& scenePath, const double specs[8], bool postProcessing, const double* startXYZ,
                     bool planLoopsAround, int device)
    : postProcessing_(postProcessing), loops_(planLoopsAround), device_(device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw std::runtime_error("libipp: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) throw std::invalid_argument("libipp: bad device ordinal");
    bind();
    IPP_CUDA_CHECK(cudaDeviceGetAttribute(&numSMs_, cudaDevAttrMultiProcessorCount, device_));
    IPP_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    IPP_CUDA_CHECK(cudaEventCreate(&evStart_));
    IPP_CUDA_CHECK(cudaEventCreate(&evStop_));
    if (startXYZ) {
        hasStart_ = true;
        for (int k = 0; k < 3; ++k) start_[k] = startXYZ[k];
    }

    // ---- scene (SceneKeeper.cpp:65-71, TriangleData.cpp:41-72) ----
    std::vector<float> tris;
    load_mesh(scenePath, tris);
    const int T = (int)(tris.size() / 9);
    for (int k = 0; k < 3; ++k) { bbmin_[k] = INFINITY; bbmax_[k] = -INFINITY; }
    for (size_t i = 0; i < tris.size(); ++i) {
        int k = (int)(i % 3);
        bbmin_[k] = std::min(bbmin_[k], tris[i]);
        bbmax_[k] = std::max(bbmax_[k], tris[i]);
    }
    // float extents (osg::BoundingBox is float), OsgHelpers.cpp:499-511
    double xl = std::fabs(bbmax_[0] - bbmin_[0]), yl = std::fabs(bbmax_[1] - bbmin_[1]), zl = std::fabs(bbmax_[2] - bbmin_[2]);
    collisionPenalty_ = kCollisionPenaltyModifier * std::max(zl, std::max(xl, yl));
    for (int k = 0; k < 3; ++k) sceneCenter_[k] = (double)((bbmin_[k] + bbmax_[k]) * 0.5f);  // bounding-sphere centre of one geode

    float* d_verts = dalloc<float>(tris.size());
    float* d_canon = dalloc<float>(tris.size());
    float* d_centroid = dalloc<float>((size_t)T * 3);
    uint8_t* d_degenerate = dalloc<uint8_t>(T);
    const int64_t areaPad = round_up(std::max(T, 1), kTileTris);
    nTiles_ = (int)(areaPad / kTileTris);
    scene_.area = dalloc<double>(areaPad);
    IPP_CUDA_CHECK(cudaMemsetAsync(scene_.area, 0, areaPad * sizeof(double), stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_verts, tris.data(), tris.size() * sizeof(float), cudaMemcpyHostToDevice, stream_));
    launch_mesh_prepare(d_verts, T, scene_.area, d_degenerate, d_canon, d_centroid, stream_);
    areaHost_.resize(T);
    IPP_CUDA_CHECK(cudaMemcpyAsync(areaHost_.data(), scene_.area, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    build_bvh(d_verts, d_canon, d_centroid, d_degenerate, T, bbmin_, bbmax_, scene_, stream_);  // synchronises
    nLaunches_ += 6;
    cudaFree(d_verts);
    cudaFree(d_canon);
    cudaFree(d_centroid);
    cudaFree(d_degenerate);
    totalArea_ = 0;
    for (int i = 0; i < T; ++i) totalArea_ += areaHost_[i];  // ascending-ID sum, TriangleData.cpp:65-72

    // ---- sensor (CameraEstimator.cpp:53-63; argument swap at ImageViewerCaptureTool.cpp:22-26) ----
    const double fovY = specs[1], aspect = specs[2] / specs[1];
    const double PI = 3.14159265358979323846;
    cam_.tanY = std::tan((fovY * 0.5) * PI / 180.0);
    cam_.tanX = cam_.tanY * aspect;
    cam_.H = (int)(unsigned)specs[3];
    cam_.W = (int)(unsigned)(cam_.H * aspect);
    cam_.zNear = (float)specs[4];
    cam_.zFar = (float)specs[5];
    cam_.front = specs[6] != 0;
    cam_.down = specs[7] != 0;
    cam_.sampleInterval = specs[0];
    if (cam_.W <= 0 || cam_.H <= 0) throw std::invalid_argument("libipp: image size must be positive");
    if (!(cam_.sampleInterval > 0)) throw std::invalid_argument("libipp: sampling interval must be positive");

    // ---- scalars ----
    d_scalars_ = dalloc<int>(8);
    IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, 8 * sizeof(int), stream_));
    IPP_CUDA_CHECK(cudaHostAlloc((void**)&h_scalars_, 8 * sizeof(int), cudaHostAllocDefault));
    d_visCounters_ = dalloc<unsigned long long>(16);
    IPP_CUDA_CHECK(cudaMemsetAsync(d_visCounters_, 0, 16 * sizeof(unsigned long long), stream_));
    d_angleOne_ = dalloc<float>(1);

    // ---- candidate viewpoints (PlanInterpreterBoxOrder::divideIntoBoxes :193-251) ----
    enumerate_grid_points(bbmin_, bbmax_, grid_);
    {
        const int64_t np = (int64_t)(grid_.points.size() / 3);
        std::vector<uint8_t> close(np), vis(np);
        boxesHitMesh(grid_.points.data(), np, kBoxSafetyMargin, close.data());
        boxesHitMesh(grid_.points.data(), np, kCameraFarPlaneDist, vis.data());
        assign_box_ids(grid_, close.data(), vis.data());
    }
    N_ = (int)(grid_.centres.size() / 3);
    if (N_ == 0) throw std::logic_error("libipp: no candidate viewpoints around this mesh");
    nKeys_ = (N_ + 1) * N_;
    {
        std::vector<double> pos(grid_.centres);
        pos.push_back(start_[0]);
        pos.push_back(start_[1]);
        pos.push_back(start_[2]);
        d_pos_ = dalloc<double>(pos.size());
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_pos_, pos.data(), pos.size() * sizeof(double), cudaMemcpyHostToDevice, stream_));
        sync();
    }

    // ---- memo tables and the edge-bitmap row pool ----
    d_coll_ = dalloc<uint8_t>(round_up(nKeys_, 4));
    d_rowmap_ = dalloc<int>(nKeys_);
    d_used_ = dalloc<uint8_t>(nKeys_);
    d_newKeys_ = dalloc<int>(nKeys_);
    rowWords_ = round_up((scene_.T + 31) / 32, 4);
    {
        size_t freeB = 0, totalB = 0;
        IPP_CUDA_CHECK(cudaMemGetInfo(&freeB, &totalB));
        double frac = 0.70;
        int64_t budget = (int64_t)(freeB * frac);
        if (const char* e = getenv("IPP_CACHE_GB")) budget = std::min<int64_t>(budget, (int64_t)(atof(e) * 1e9));
        // flag buffer for rendering batches comes out of the remaining memory
        int64_t flagBudget = std::min<int64_t>((int64_t)2 << 30, (int64_t)(freeB * 0.05));
        batchCap_ = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(65536, nKeys_), flagBudget / std::max(scene_.Tpad, 32)));
        // worst-case tiles per edge keep the tile prefix sums inside int32
        double diag = 0;
        for (int k = 0; k < 3; ++k) {
            double e = (double)(bbmax_[k] - bbmin_[k]) + 2 * kBoxPadding;
            diag += e * e;
        }
        diag = std::sqrt(diag);
        int64_t worstSnaps = ((int64_t)std::ceil(diag / cam_.sampleInterval) + 3) * 2;
        int64_t tilesPerImage = (int64_t)((cam_.W + 31) / 32) * ((cam_.H + 31) / 32);
        batchCap_ = (int)std::max<int64_t>(1, std::min<int64_t>(batchCap_, ((int64_t)1 << 30) / std::max<int64_t>(1, worstSnaps * tilesPerImage)));
        rowCapacity_ = std::min<int64_t>(nKeys_, budget / (rowWords_ * 4));
        if (rowCapacity_ < 1) throw std::runtime_error("libipp: not enough device memory for the edge-bitmap cache");
    }
    // K4 variant and the per-batch snapshot budget of its leaf tables
    useBins_ = bins_supported(scene_, cam_);
    if (const char* e = getenv("IPP_VISIBILITY")) {
        if (std::string(e) == "bvh") useBins_ = false;
    }
    {
        size_t freeB = 0, totalB = 0;
        IPP_CUDA_CHECK(cudaMemGetInfo(&freeB, &totalB));
        const int64_t tilesPerImage = (int64_t)((cam_.W + 31) / 32) * ((cam_.H + 31) / 32);
        maxSnapsPerBatch_ = std::max<int64_t>(64, ((int64_t)1 << 30) / std::max<int64_t>(1, tilesPerImage));
        bins_.stride = bins_table_stride(scene_.nLeaves);
        bins_.gridBlocks = bins_grid_blocks();
        if (useBins_) {
            const int64_t tableBudget = std::min<int64_t>((int64_t)16 << 30, (int64_t)(freeB * 0.10));
            maxSnapsPerBatch_ = std::max<int64_t>(64, std::min<int64_t>(maxSnapsPerBatch_, tableBudget / ((int64_t)bins_.stride * 8)));
        }
    }
    if (const char* e = getenv("IPP_POSE_DEDUPE")) poseDedupe_ = atoi(e) != 0;
    if (const char* e = getenv("IPP_MAX_POSES_PER_BATCH")) maxSnapsPerBatch_ = std::max<int64_t>(1, std::min<int64_t>(maxSnapsPerBatch_, atoll(e)));
    d_countsAll_ = dalloc<int>(nKeys_ + 1);
    d_rows_ = dalloc<uint32_t>((size_t)(rowCapacity_ + 1) * rowWords_);
    d_freeStack_ = dalloc<int>(rowCapacity_);
    d_flags_ = dalloc<uint8_t>((size_t)batchCap_ * scene_.Tpad);
    IPP_CUDA_CHECK(cudaMemsetAsync(d_flags_, 0, (size_t)batchCap_ * scene_.Tpad, stream_));
    d_rowOfSlot_ = dalloc<int>(batchCap_);
    d_counts_ = dalloc<int>(batchCap_ + 1);
    d_snapOffset_ = dalloc<int>(batchCap_ + 1);
    d_union_ = dalloc<uint32_t>(rowWords_);
    clearMemo();

    // ---- seeds and the energy limit (ProduceSimpleCirclingPlans :40-109, PlanEnergyEvaluator.cpp:40-46) ----
    if (!postProcessing_) {
        levelPlans_ = single_level_circling_plans(grid_);
        completePlans_ = complete_circling_plans(levelPlans_);
        if (completePlans_.empty()) throw std::logic_error("libipp: no circling plans could be generated for this mesh");
        // energy of the longest circling plan, decoded WITHOUT loop closure (decodeAndCalculateEnergyUsage is
        // called directly, not through evaluatePlan)
        const Plan& front = completePlans_.front();
        std::vector<int32_t> ids;
        for (const auto& g : front) ids.push_back((int32_t)(size_t)(g[0] + 0.5));  // -1 -> 0, SURVEY A.11
        std::vector<int64_t> off = {0, (int64_t)ids.size()};
        double fit[4];
        bool savedLoops = loops_;
        loops_ = false;
        maxAllowedEnergy_ = -INFINITY;  // energy only: with this limit every plan fails the gate, no coverage work
        evaluatePopulationHost(ids.data(), nullptr, off.data(), 1, true, false, fit, nullptr);
        loops_ = savedLoops;
        maxAllowedEnergy_ = fit[1] * kMaxEnergyMultiplier;
    } else {
        maxAllowedEnergy_ = -1;
    }
    int64_t c[8];
    counters(c, true);
}

Evaluator::~Evaluator() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    free_bvh(scene_);
    cudaFree(scene_.area);
    cudaFree(d_pos_); cudaFree(d_coll_); cudaFree(d_rowmap_); cudaFree(d_used_); cudaFree(d_evictUsed_); cudaFree(d_newKeys_);
    cudaFree(d_scalars_); cudaFreeHost(h_scalars_); cudaFree(d_rows_); cudaFree(d_freeStack_);
    cudaFree(d_flags_); cudaFree(d_rowOfSlot_); cudaFree(d_counts_); cudaFree(d_snapOffset_);
    cudaFree(d_snaps_); cudaFree(d_ntiles_); cudaFree(d_tileOffset_); cudaFree(d_scanTmp_);
    cudaFree(d_snapsU_); cudaFree(d_ntilesU_); cudaFree(d_slotList_); cudaFree(d_dedupeWork_); cudaFree(d_dedupeIWork_); cudaFree(d_dedupeTmp_);
    cudaFree(d_visCounters_); cudaFree(d_angleOne_);
    cudaFree(bins_.table); cudaFree(bins_.tableCount); cudaFree(bins_.tileMask); cudaFree(bins_.superMask); cudaFree(bins_.nSuper);
    cudaFree(bins_.scratch); cudaFree(bins_.records); cudaFree(d_superOffset_); cudaFree(d_countsAll_);
    cudaFree(d_partial_);
    cudaFree(d_pseudoOffsets_); cudaFree(d_pseudoEdgeOffsets_); cudaFree(d_pseudoActive_); cudaFree(d_pseudoFitness_);
    cudaFree(d_angleBuf_); cudaFree(d_edgeRows_); cudaFree(d_edgeOffsets_); cudaFree(d_rowList_);
    cudaFree(d_sortedKeys_); cudaFree(d_sortTmp_); cudaFree(d_exchange_); cudaFree(d_activeShared_); cudaFree(d_fitShared_);
    cudaFree(d_ids_); cudaFree(d_offsets_); cudaFree(d_fitness_); cudaFree(d_active_); cudaFree(d_union_);
    cudaFree(d_flushBuf_);
    for (auto& e : pendingEv_) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : freeEv_) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaEventDestroy(evStart_);
    cudaEventDestroy(evStop_);
    cudaStreamDestroy(stream_);
}

// ---------------------------------------------------------------------------------------------
void Evaluator::readScalars() {
    IPP_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, d_scalars_, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    sync();
}
int Evaluator::readInt(const int* d) {
    IPP_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, d, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    sync();
    return h_scalars_[0];
}

void Evaluator::groupBegin(int g) {
    if (!profiling_) return;
    EvPair e;
    if (!freeEv_.empty()) {
        e = freeEv_.back();
        freeEv_.pop_back();
    } else {
        IPP_CUDA_CHECK(cudaEventCreate(&e.a));
        IPP_CUDA_CHECK(cudaEventCreate(&e.b));
    }
    e.group = g;
    IPP_CUDA_CHECK(cudaEventRecord(e.a, stream_));
    pendingEv_.push_back(e);
}
void Evaluator::groupEnd(int g) {
    if (!profiling_) return;
    IPP_CUDA_CHECK(cudaEventRecord(pendingEv_.back().b, stream_));
    groupLaunches_[g]++;
    if (pendingEv_.size() > 4096) harvestTimings(false);
}
void Evaluator::harvestTimings(bool wait) {
    size_t keep = 0;
    for (size_t i = 0; i < pendingEv_.size(); ++i) {
        EvPair& e = pendingEv_[i];
        cudaError_t q = wait ? cudaEventSynchronize(e.b) : cudaEventQuery(e.b);
        if (q == cudaSuccess) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e.a, e.b);
            groupMs_[e.group] += ms;
            freeEv_.push_back(e);
        } else {
            pendingEv_[keep++] = e;
        }
    }
    pendingEv_.resize(keep);
    cudaGetLastError();
}
void Evaluator::kernelTime(int which, double* ms, int64_t* launches, bool reset) {
    if (which < 0 || which >= KG_COUNT) throw std::invalid_argument("libipp: bad kernel group");
    bind();
    harvestTimings(true);
    *ms = groupMs_[which];
    *launches = groupLaunches_[which];
    if (reset) {
        groupMs_[which] = 0;
        groupLaunches_[which] = 0;
    }
}
void Evaluator::counters(int64_t out[8], bool reset) {
    bind();
    unsigned long long rays = 0;
    IPP_CUDA_CHECK(cudaMemcpyAsync(&rays, d_visCounters_, sizeof(rays), cudaMemcpyDeviceToHost, stream_));
    sync();
    out[0] = nLaunches_; out[1] = (int64_t)rays; out[2] = nSnapshots_; out[3] = nEdgesRendered_;
    out[4] = nCollisionQueries_; out[5] = nReduceLaunches_; out[6] = nPlansReduced_; out[7] = nSnapshotsRendered_;
    if (reset) {
        nLaunches_ = nSnapshots_ = nEdgesRendered_ = nCollisionQueries_ = nReduceLaunches_ = nPlansReduced_ = nRowsRead_ = nSnapshotsRendered_ = 0;
        IPP_CUDA_CHECK(cudaMemsetAsync(d_visCounters_, 0, 16 * sizeof(unsigned long long), stream_));
    }
}
void Evaluator::visibilityStats(int64_t out[8], bool reset) {
    bind();
    unsigned long long c[16];
    IPP_CUDA_CHECK(cudaMemcpyAsync(c, d_visCounters_, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    sync();
    // out[0]: 64-byte units of acceleration-structure data fetched = BVH nodes (visibility.cu) + 8-byte bin entries / 8 (visibility_bins.cu)
    out[0] = (int64_t)c[2] + (int64_t)(c[5] / 8); out[1] = (int64_t)c[3]; out[2] = (int64_t)c[4]; out[3] = (int64_t)c[0];
    out[4] = (int64_t)c[6]; out[5] = (int64_t)c[7];  // binned variant only
    out[6] = (int64_t)c[8]; out[7] = (int64_t)c[9];  // supertiles rasterised from set-up records / by the leaf walk (overflow)
    if (reset) IPP_CUDA_CHECK(cudaMemsetAsync(d_visCounters_, 0, 16 * sizeof(unsigned long long), stream_));
}
void Evaluator::timerStart() {
    bind();
    IPP_CUDA_CHECK(cudaEventRecord(evStart_, stream_));
}
double Evaluator::timerStop() {
    bind();
    IPP_CUDA_CHECK(cudaEventRecord(evStop_, stream_));
    IPP_CUDA_CHECK(cudaEventSynchronize(evStop_));
    float ms = 0;
    IPP_CUDA_CHECK(cudaEventElapsedTime(&ms, evStart_, evStop_));
    return ms;
}
void Evaluator::flushL2() {
    bind();
    if (!d_flushBuf_) {
        flushWords_ = (int64_t)(256 << 20) / 4;  // 256 MiB > 126 MB L2
        d_flushBuf_ = dalloc<uint32_t>(flushWords_);
    }
    launch_flush(d_flushBuf_, flushWords_, stream_);
    nLaunches_++;
}

// Measured read bandwidth of a buffer of `bytes` bytes read `reps` times by every SM (GB/s): a buffer well inside L2 (e.g. 32 MiB)
// gives the L2 read peak this box reaches with plain 16-byte loads, one larger than L2 the HBM read peak.
double Evaluator::readBandwidth(int64_t bytes, int reps) {
    bind();
    if (!d_flushBuf_) flushL2();
    const int64_t words = std::min<int64_t>(flushWords_, std::max<int64_t>(bytes / 4, 1024));
    unsigned* d_sink = reinterpret_cast<unsigned*>(d_scalars_ + 7);
    launch_read_probe(d_flushBuf_, words, 2, d_sink, numSMs_, stream_);  // warm: the buffer is in L2 if it fits
    cudaEvent_t a, b;
    IPP_CUDA_CHECK(cudaEventCreate(&a));
    IPP_CUDA_CHECK(cudaEventCreate(&b));
    IPP_CUDA_CHECK(cudaEventRecord(a, stream_));
    launch_read_probe(d_flushBuf_, words, reps, d_sink, numSMs_, stream_);
    IPP_CUDA_CHECK(cudaEventRecord(b, stream_));
    IPP_CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0;
    IPP_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    nLaunches_ += 2;
    return ms > 0 ? (double)words * 4.0 * reps / (ms * 1e-3) / 1e9 : 0.0;
}

void Evaluator::exclusiveScan(const int* d_in, int* d_out, int n) {
    size_t need = scan_tmp_bytes(n);
    if (need > scanTmpBytes_) {
        cudaFree(d_scanTmp_);
        scanTmpBytes_ = need * 2;
        IPP_CUDA_CHECK(cudaMalloc(&d_scanTmp_, scanTmpBytes_));
    }
    exclusive_scan_i32(d_in, d_out, n, d_scanTmp_, scanTmpBytes_, stream_);
    nLaunches_++;
}

void Evaluator::ensurePlanBuffers(int64_t pop, int64_t nGenes) {
    if (pop > popCap_) {
        cudaFree(d_offsets_);
        cudaFree(d_fitness_);
        popCap_ = pop + pop / 4 + 16;
        d_offsets_ = dalloc<int64_t>(popCap_ + 1);
        d_fitness_ = dalloc<double>(popCap_ * 4);
    }
    if (nGenes > geneCap_) {
        cudaFree(d_ids_);
        geneCap_ = nGenes + nGenes / 4 + 16;
        d_ids_ = dalloc<int32_t>(geneCap_);
    }
}
void Evaluator::ensureSnapBuffers(int64_t nSnaps) {
    if (nSnaps + 1 > snapCap_) {
        cudaFree(d_snaps_);
        cudaFree(d_ntiles_);
        cudaFree(d_tileOffset_);
        snapCap_ = std::min<int64_t>(nSnaps + nSnaps / 4 + 64, std::max<int64_t>(maxSnapsPerBatch_, nSnaps) + 1);
        d_snaps_ = dalloc<Snapshot>(snapCap_);
        d_ntiles_ = dalloc<int>(snapCap_);
        d_tileOffset_ = dalloc<int>(snapCap_);
        cudaFree(d_snapsU_); cudaFree(d_ntilesU_); cudaFree(d_slotList_); cudaFree(d_dedupeWork_); cudaFree(d_dedupeIWork_);
        d_snapsU_ = dalloc<Snapshot>(snapCap_);
        d_ntilesU_ = dalloc<int>(snapCap_);
        d_slotList_ = dalloc<int>(snapCap_);
        d_dedupeWork_ = dalloc<unsigned long long>((size_t)snapCap_ * 2);
        d_dedupeIWork_ = dalloc<int>((size_t)snapCap_ * 7);
        if (useBins_) {
            cudaFree(bins_.table); cudaFree(bins_.tableCount); cudaFree(bins_.tileMask); cudaFree(bins_.superMask); cudaFree(bins_.nSuper);
            cudaFree(d_superOffset_);
            bins_.table = dalloc<unsigned long long>((size_t)snapCap_ * bins_.stride);
            bins_.tableCount = dalloc<int>(snapCap_);
            bins_.tileMask = dalloc<uint32_t>((size_t)snapCap_ * 32);
            bins_.superMask = dalloc<unsigned long long>(snapCap_);
            bins_.nSuper = dalloc<int>(snapCap_);
            d_superOffset_ = dalloc<int>(snapCap_);
        }
    }
    if (useBins_ && !bins_.scratch)
        bins_.scratch = dalloc<unsigned long long>((size_t)bins_.gridBlocks * bins_warps_per_block() * bins_.stride);
    if (useBins_ && !bins_.records)
        bins_.records = reinterpret_cast<float4*>(dalloc<float>((size_t)bins_.gridBlocks * bins_warps_per_block() * bins_record_floats_per_warp()));
}

void Evaluator::setUseBins(bool on) {
    bind();
    sync();
    on = on && bins_supported(scene_, cam_);
    if (on == useBins_) return;
    useBins_ = on;
    snapCap_ = 0;  // buffers are re-made for the variant on the next batch
}

// K4 for the nSnaps camera poses in d_snaps_ (either variant); d_idImage: test hook.  Poses that several edges of the batch share
// are rendered once (pose_dedupe.cu): the kernels see the unique poses in d_snapsU_.
void Evaluator::renderSnapshots(int nSnaps, int32_t* d_idImage) {
    if (nSnaps <= 0) return;
    const size_t need = pose_dedupe_tmp_bytes(nSnaps);
    if (need > dedupeTmpBytes_) {
        cudaFree(d_dedupeTmp_);
        dedupeTmpBytes_ = need * 2;
        IPP_CUDA_CHECK(cudaMalloc(&d_dedupeTmp_, dedupeTmpBytes_));
    }
    groupBegin(KG_PLAN);
    launch_pose_dedupe(d_snaps_, d_ntiles_, nSnaps, poseDedupe_ && nSnaps > 1, d_snapsU_, d_ntilesU_, d_slotList_, d_scalars_ + 5, d_dedupeWork_,
                       d_dedupeIWork_, d_dedupeTmp_, dedupeTmpBytes_, stream_);
    groupEnd(KG_PLAN);
    nLaunches_ += (poseDedupe_ && nSnaps > 1) ? 7 : 1;
    const int nU = readInt(d_scalars_ + 5);
    nSnapshotsRendered_ += nU;
    if (useBins_) {
        groupBegin(KG_TABLE);
        launch_leaf_table(scene_, cam_, d_snapsU_, nU, bins_, stream_);
        groupEnd(KG_TABLE);
        if (d_idImage) {
            // test hook: every supertile of the image is scheduled so that background pixels are reported too
            const int nsx = (cam_.W + 127) / 128, nsy = (cam_.H + 127) / 128;
            unsigned long long m = 0;
            for (int y = 0; y < nsy; ++y)
                for (int x = 0; x < nsx; ++x) m |= 1ull << (y * 8 + x);
            int cnt = nsx * nsy;
            IPP_CUDA_CHECK(cudaMemcpyAsync(bins_.superMask, &m, sizeof(m), cudaMemcpyHostToDevice, stream_));
            IPP_CUDA_CHECK(cudaMemcpyAsync(bins_.nSuper, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, stream_));
            sync();
        }
        IPP_CUDA_CHECK(cudaMemsetAsync(bins_.nSuper + nU, 0, sizeof(int), stream_));
        exclusiveScan(bins_.nSuper, d_superOffset_, nU + 1);
        int totalSuper = readInt(d_superOffset_ + nU);
        groupBegin(KG_VISIBILITY);
        launch_visibility_bins(scene_, cam_, d_snapsU_, d_superOffset_, nU, totalSuper, bins_, d_flags_, d_slotList_, d_idImage, d_visCounters_,
                               stream_);
        groupEnd(KG_VISIBILITY);
        nLaunches_ += 3;
    } else {
        IPP_CUDA_CHECK(cudaMemsetAsync(d_ntilesU_ + nU, 0, sizeof(int), stream_));
        exclusiveScan(d_ntilesU_, d_tileOffset_, nU + 1);
        int totalTiles = readInt(d_tileOffset_ + nU);
        groupBegin(KG_VISIBILITY);
        launch_visibility(scene_, cam_, d_snapsU_, d_tileOffset_, nU, totalTiles, d_flags_, d_slotList_, d_idImage, d_visCounters_, stream_);
        groupEnd(KG_VISIBILITY);
        nLaunches_ += 2;
    }
}

void Evaluator::clearMemo() {
    bind();
    launch_fill_u8(d_coll_, round_up(nKeys_, 4), 0xFF, stream_);
    launch_fill_i32(d_rowmap_, nKeys_, -1, stream_);
    launch_iota_i32(d_freeStack_, rowCapacity_, stream_);
    int top = (int)rowCapacity_;
    h_scalars_[1] = top;
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_scalars_ + 1, h_scalars_ + 1, sizeof(int), cudaMemcpyHostToDevice, stream_));
    sync();
    nLaunches_ += 3;
    liveRows_ = 0;
    angleMemo_.clear();
    headingReversed_.clear();
}

// ---------------------------------------------------------------------------------------------
// K3 + K4 + pack for n freshly claimed keys (batched by the flag buffer capacity)
// ---------------------------------------------------------------------------------------------
void Evaluator::renderBatch(const int* d_keys, const float* d_angles, int n, const int* d_rowOfSlot) {
    // samples per edge -> snapshot offsets
    launch_count_snapshots(d_keys, n, d_pos_, N_, cam_, d_counts_, stream_);
    IPP_CUDA_CHECK(cudaMemsetAsync(d_counts_ + n, 0, sizeof(int), stream_));
    exclusiveScan(d_counts_, d_snapOffset_, n + 1);
    int nSnaps = readInt(d_snapOffset_ + n);
    ensureSnapBuffers(nSnaps);
    groupBegin(KG_PLAN);
    launch_fill_snapshots(d_keys, d_angles, n, d_pos_, N_, sceneCenter_, cam_, scene_, d_snapOffset_, d_snaps_, d_ntiles_, stream_);
    groupEnd(KG_PLAN);
    renderSnapshots(nSnaps, nullptr);
    groupBegin(KG_PACK);
    launch_pack_rows(d_flags_, n, scene_.Tpad, d_rowOfSlot, d_rows_, rowWords_, stream_);
    groupEnd(KG_PACK);
    nLaunches_ += 3;
    nSnapshots_ += nSnaps;
    nEdgesRendered_ += n;
}

void Evaluator::renderNewEdges(int nNew, const float* d_angles) { renderKeys(d_newKeys_, nNew, d_angles); }

// K3 + K4 + pack for n claimed keys, in batches that respect the flag rows and the pose budget.  publish: the rows become the
// memo entries of the keys (rowmap); otherwise they are only handed back through rowsOut ([id, angle] genes keep their own memo).
void Evaluator::renderKeys(const int* d_keys, int nNew, const float* d_angles, bool publish, std::vector<int>* rowsOut) {
    if (nNew <= 0) return;
    // camera poses per edge, so that a batch respects both the flag rows and the pose budget (leaf tables, int32 work offsets)
    std::vector<int> perEdge(nNew);
    launch_count_snapshots(d_keys, nNew, d_pos_, N_, cam_, d_countsAll_, stream_);
    IPP_CUDA_CHECK(cudaMemcpyAsync(perEdge.data(), d_countsAll_, (size_t)nNew * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    sync();
    nLaunches_++;
    if (rowsOut) rowsOut->resize(nNew);
    for (int b0 = 0, n = 0; b0 < nNew; b0 += n) {
        int64_t snaps = 0;
        for (n = 0; b0 + n < nNew && n < batchCap_ && (n == 0 || snaps + perEdge[b0 + n] <= maxSnapsPerBatch_); ++n) snaps += perEdge[b0 + n];
        launch_assign_rows(d_keys + b0, n, d_rowmap_, d_rowOfSlot_, d_freeStack_, d_scalars_ + 1, (int)rowCapacity_, d_scalars_ + 4, stream_);
        renderBatch(d_keys + b0, d_angles ? d_angles + b0 : nullptr, n, d_rowOfSlot_);
        if (publish) launch_publish_rows(d_keys + b0, d_rowOfSlot_, n, d_rowmap_, stream_);
        if (rowsOut) {
            IPP_CUDA_CHECK(cudaMemcpyAsync(rowsOut->data() + b0, d_rowOfSlot_, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream_));
            sync();
        }
        nLaunches_ += 2;
    }
    if (publish) liveRows_ += nNew;
    readScalars();
    if (h_scalars_[4]) {
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 4, 0, sizeof(int), stream_));
        throw std::logic_error("libipp: internal error: the edge-bitmap row pool ran dry while rows were being assigned");
    }
}

int64_t Evaluator::countUnseenEdges(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit) {
    bind();
    if (pop <= 0) return 0;
    if (pop > sharedCap_) {
        cudaFree(d_activeShared_);
        cudaFree(d_fitShared_);
        sharedCap_ = pop + pop / 4 + 16;
        d_activeShared_ = dalloc<uint8_t>(sharedCap_);
        d_fitShared_ = dalloc<double>(sharedCap_ * 4);
    }
    energyAndGate(d_ids, d_offsets, pop, noLimit, d_fitShared_);  // (active flags land in d_active_)
    PlanMarkArgs mk;
    mk.ids = d_ids; mk.offsets = d_offsets; mk.pop = pop; mk.N = N_; mk.hasStart = hasStart_; mk.loops = loops_;
    mk.active = d_active_; mk.coll = nullptr; mk.rowmap = d_rowmap_; mk.used = nullptr; mk.newKeys = d_newKeys_; mk.newCount = d_scalars_;
    mk.errFlag = d_scalars_ + 3;
    IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, sizeof(int), stream_));
    launch_probe_rows(mk, stream_);
    nLaunches_++;
    return readInt(d_scalars_);
}

int64_t Evaluator::renderEdgesShared(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit, Collectives& cx) {
    bind();
    if (pop <= 0) return 0;
    requireQuickHeading();
    if (pop > sharedCap_) {
        cudaFree(d_activeShared_);
        cudaFree(d_fitShared_);
        sharedCap_ = pop + pop / 4 + 16;
        d_activeShared_ = dalloc<uint8_t>(sharedCap_);
        d_fitShared_ = dalloc<double>(sharedCap_ * 4);
    }
    // 1. collision flags of the whole population (replicated: a millisecond per 20 k edges), energy gate, then the edges
    //    whose bitmap nobody has yet
    // 2. every rank must have found the same set (same population, same memo contents); every decision below is taken on
    //    all-reduced values, so the ranks take the same branch and no rank leaves the others waiting in a collective:
    //      * the ranks disagree (their memos diverged, e.g. one of them had to evict rows while it evaluated its shard): every
    //        rank clears its memo and marks again; a second disagreement means different populations and is an error;
    //      * the new rows do not fit: the rows of edges the population does not use are evicted everywhere, and if that is
    //        not enough the shared rendering is skipped (-1): every rank then renders what its own shard needs, in pieces
    //        (coverageRange), and the next shared call starts from cleared memos.
    int nNew = 0;
    bool evicted = false;
    for (int round = 0;; ++round) {
        PlanMarkArgs mk;
        mk.ids = d_ids; mk.offsets = d_offsets; mk.pop = pop; mk.N = N_; mk.hasStart = hasStart_; mk.loops = loops_;
        mk.active = nullptr; mk.coll = d_coll_; mk.rowmap = nullptr; mk.used = nullptr; mk.newKeys = d_newKeys_; mk.newCount = d_scalars_;
        mk.errFlag = d_scalars_ + 3;
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, sizeof(int), stream_));
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 3, 0, sizeof(int), stream_));
        launch_mark_collisions(mk, stream_);
        nLaunches_++;
        readScalars();
        const bool badId = h_scalars_[3] != 0;
        if (h_scalars_[0] > 0) {
            launch_collide_keys(scene_, d_newKeys_, h_scalars_[0], d_pos_, N_, d_coll_, stream_);
            nLaunches_++;
            nCollisionQueries_ += h_scalars_[0];
        }
        if (badId) throw std::out_of_range("libipp: box id out of range");  // (the same population on every rank: all of them throw)
        PlanEnergyArgs en;
        en.ids = d_ids; en.offsets = d_offsets; en.pop = pop; en.N = N_; en.hasStart = hasStart_; en.loops = loops_;
        en.pos = d_pos_; en.coll = d_coll_; en.collisionPenalty = collisionPenalty_; en.maxEnergy = maxAllowedEnergy_;
        en.disableLimit = noLimit; en.fitness = d_fitShared_; en.active = d_activeShared_;
        launch_energy(en, stream_);
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, sizeof(int), stream_));
        mk.active = d_activeShared_; mk.coll = nullptr; mk.rowmap = d_rowmap_;
        launch_mark_rows(mk, stream_);
        nLaunches_ += 2;
        readScalars();
        nNew = h_scalars_[0];
        const int freeRows = h_scalars_[1];
        double agree[3] = {(double)nNew, -(double)nNew, nNew > freeRows ? 1.0 : 0.0};
        cx.allReduceMax(agree, 3, stream_);
        if (agree[0] != (double)nNew || agree[1] != -(double)nNew || memoDiverged_) {
            launch_unclaim_rows(d_newKeys_, nNew, d_rowmap_, stream_);
            nLaunches_++;
            if (round > 0 && !memoDiverged_)
                throw std::runtime_error("libipp: the ranks disagree on the new edges of the population (different populations)");
            memoDiverged_ = false;
            clearMemo();
            continue;
        }
        if (nNew == 0) return 0;
        if (agree[2] == 0.0) break;
        launch_unclaim_rows(d_newKeys_, nNew, d_rowmap_, stream_);
        nLaunches_++;
        if (evicted) {
            memoDiverged_ = true;  // (every rank takes this branch)
            return -1;
        }
        PlanMarkArgs keep = mk;
        keep.rowmap = nullptr; keep.used = d_evictUsed();
        launch_fill_u8(keep.used, nKeys_, 0, stream_);
        launch_mark_rows(keep, stream_);
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 2, 0, sizeof(int), stream_));
        launch_gc_rows(d_rowmap_, keep.used, nKeys_, d_freeStack_, d_scalars_ + 1, d_scalars_ + 2, stream_);
        nLaunches_ += 3;
        liveRows_ = readInt(d_scalars_ + 2);
        nEvictions_++;
        evicted = true;
    }
    // 3. the same order everywhere, and a random one: neighbouring keys (edges leaving the same viewpoint) cost alike, contiguous
    //    chunks of the ascending order gave one rank of two 1.65x the mean work
    if (!d_sortedKeys_) d_sortedKeys_ = dalloc<int>(nKeys_);
    size_t need = sort_tmp_bytes(nNew);
    if (need > sortTmpBytes_) {
        cudaFree(d_sortTmp_);
        sortTmpBytes_ = need * 2;
        IPP_CUDA_CHECK(cudaMalloc(&d_sortTmp_, sortTmpBytes_));
    }
    scrambled_order_i32(d_newKeys_, d_sortedKeys_, nNew, d_sortTmp_, sortTmpBytes_, stream_);
    nLaunches_++;
    // 4. this rank renders its contiguous chunk
    const int G = cx.world, r = cx.rank;
    const int chunk = (nNew + G - 1) / G;
    const int lo = std::min(nNew, r * chunk), hi = std::min(nNew, lo + chunk);
    renderKeys(d_sortedKeys_ + lo, hi - lo, nullptr);
    // 5. exchange the fresh rows slab by slab: pack -> in-place all-gather -> rows for the other ranks' keys -> unpack
    const int64_t rowBytes = rowWords_ * 4;
    const int64_t slabBudget = (int64_t)1 << 30;
    const int B = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(chunk, batchCap_), slabBudget / (rowBytes * G)));
    const size_t needWords = (size_t)G * B * rowWords_;
    if (needWords > exchangeWords_) {
        cudaFree(d_exchange_);
        exchangeWords_ = needWords;
        d_exchange_ = dalloc<uint32_t>(exchangeWords_);
    }
    for (int s0 = 0; s0 < chunk; s0 += B) {
        const int myN = std::max(0, std::min(B, hi - (lo + s0)));
        launch_rows_gather(d_sortedKeys_ + lo + s0, myN, d_rowmap_, d_rows_, rowWords_, d_exchange_ + (size_t)r * B * rowWords_, stream_);
        cx.allGatherInPlace(d_exchange_, (size_t)B * rowBytes, stream_);
        nLaunches_ += 2;
        for (int q = 0; q < G; ++q) {
            if (q == r) continue;
            const int qlo = std::min(nNew, q * chunk), qhi = std::min(nNew, qlo + chunk);
            const int qN = std::max(0, std::min(B, qhi - (qlo + s0)));
            if (qN == 0) continue;
            const int* keys = d_sortedKeys_ + qlo + s0;
            launch_assign_rows(keys, qN, d_rowmap_, d_rowOfSlot_, d_freeStack_, d_scalars_ + 1, (int)rowCapacity_, d_scalars_ + 4, stream_);
            launch_rows_scatter(d_exchange_ + (size_t)q * B * rowWords_, d_rowOfSlot_, qN, d_rows_, rowWords_, stream_);
            launch_publish_rows(keys, d_rowOfSlot_, qN, d_rowmap_, stream_);
            nLaunches_ += 3;
            liveRows_ += qN;
        }
    }
    sync();
    return hi - lo;
}

int Evaluator::headingChunk() const { return std::max(1, std::min(nKeys_ / 2, batchCap_)); }
uint8_t* Evaluator::d_evictUsed() {
    if (!d_evictUsed_) d_evictUsed_ = dalloc<uint8_t>(nKeys_);
    return d_evictUsed_;
}

double* Evaluator::reducePartials(int64_t pop) {
    if (nTiles_ <= 8 || pop > (int64_t)numSMs_ * 8 * 2) return nullptr;  // (launch_plan_reduce splits below two plan groups per SM)
    const int64_t need = pop * nTiles_;
    if (need > partialCap_) {
        cudaFree(d_partial_);
        partialCap_ = need + need / 4 + 16;
        d_partial_ = dalloc<double>(partialCap_);
    }
    return d_partial_;
}

// Steps 1-2 of evaluatePlan for a population: collision flags never seen before -> K5, then energy, path length, colliding-edge
// count and the energy gate (fitness rows, active flags).  Independent of heading offsets.
void Evaluator::energyAndGate(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit, double* d_fitness) {
    if (pop > activeCap_) {
        cudaFree(d_active_);
        activeCap_ = pop + pop / 4 + 16;
        d_active_ = dalloc<uint8_t>(activeCap_);
    }
    PlanMarkArgs mk;
    mk.ids = d_ids; mk.offsets = d_offsets; mk.pop = pop; mk.N = N_; mk.hasStart = hasStart_; mk.loops = loops_;
    mk.active = nullptr; mk.coll = d_coll_; mk.rowmap = nullptr; mk.used = nullptr; mk.newKeys = d_newKeys_; mk.newCount = d_scalars_;
    mk.errFlag = d_scalars_ + 3;

    // 1. collision flags never seen before -> K5 (also validates the box ids: assert at PlanInterpreterBoxOrder.cpp:282)
    IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, sizeof(int), stream_));
    IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 3, 0, sizeof(int), stream_));
    groupBegin(KG_PLAN);
    launch_mark_collisions(mk, stream_);
    groupEnd(KG_PLAN);
    nLaunches_++;
    readScalars();
    const bool badId = h_scalars_[3] != 0;
    int nColl = h_scalars_[0];
    if (nColl > 0) {
        // (also when a plan carried a bad id: the flags the other plans claimed are resolved, none is left pending)
        groupBegin(KG_COLLISION);
        launch_collide_keys(scene_, d_newKeys_, nColl, d_pos_, N_, d_coll_, stream_);
        groupEnd(KG_COLLISION);
        nLaunches_++;
        nCollisionQueries_ += nColl;
    }
    if (badId) throw std::out_of_range("libipp: box id out of range");

    // 2. energy + gate
    PlanEnergyArgs en;
    en.ids = d_ids; en.offsets = d_offsets; en.pop = pop; en.N = N_; en.hasStart = hasStart_; en.loops = loops_;
    en.pos = d_pos_; en.coll = d_coll_; en.collisionPenalty = collisionPenalty_; en.maxEnergy = maxAllowedEnergy_;
    en.disableLimit = noLimit; en.fitness = d_fitness; en.active = d_active_;
    groupBegin(KG_PLAN);
    launch_energy(en, stream_);
    nLaunches_++;
    groupEnd(KG_PLAN);
}

// ---------------------------------------------------------------------------------------------
void Evaluator::evaluatePopulationDevice(const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t pop,
                                         int64_t nGenes, bool memo, bool noLimit, double* d_fitness, uint32_t* unionBitmapHost) {
    bind();
    if (pop <= 0) return;
    if (d_angles) {
        // [id, angleOffset] genes: the memo of their edges is keyed on the host (edge key, offset bits), so the genes come back once
        // (12 bytes per gene); everything else stays where it is
        std::vector<int32_t> ids((size_t)nGenes);
        std::vector<float> angles((size_t)nGenes);
        std::vector<int64_t> offsets((size_t)pop + 1);
        IPP_CUDA_CHECK(cudaMemcpyAsync(ids.data(), d_ids, (size_t)nGenes * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(angles.data(), d_angles, (size_t)nGenes * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(offsets.data(), d_offsets, (size_t)(pop + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, stream_));
        sync();
        if (offsets[0] != 0 || offsets[pop] != nGenes) throw std::invalid_argument("libipp: offsets must start at 0 and end at n_genes");
        evaluatePopulationAngles(ids.data(), angles.data(), offsets.data(), pop, memo, noLimit, d_ids, d_offsets, d_fitness, unionBitmapHost);
        return;
    }
    energyAndGate(d_ids, d_offsets, pop, noLimit, d_fitness);
    const bool wantUnion = unionBitmapHost != nullptr;
    if (wantUnion) IPP_CUDA_CHECK(cudaMemsetAsync(d_union_, 0, (size_t)rowWords_ * 4, stream_));
    coverageRange(d_ids, d_offsets, 0, pop, memo, d_fitness, wantUnion);
    if (wantUnion) {
        IPP_CUDA_CHECK(cudaMemcpyAsync(unionBitmapHost, d_union_, (size_t)bitmapWords() * 4, cudaMemcpyDeviceToHost, stream_));
        sync();
    }
}

// Steps 3-4 of evaluatePlan for the plans [lo, hi) of a population whose energy gate has run (d_active_): edge bitmaps the memo
// does not hold -> K3, K4, pack; then K6.  The reference's memo is a std::map that never refuses an entry
// (PlanCoverageEstimator.cpp:70,102-110); the row pool here is finite, so when the new edges of the range do not fit, the rows of
// edges the range does not use are evicted first (they would be re-rendered on their next use, exactly as after
// updateMemoisedEdges), and if the range still needs more rows than the pool has it is evaluated in two halves, recursively.
// Results do not depend on what was cached.
void Evaluator::coverageRange(const int32_t* d_ids, const int64_t* d_offsets, int64_t lo, int64_t hi, bool memo, double* d_fitness,
                              bool wantUnion) {
    const int64_t pop = hi - lo;
    if (pop <= 0) return;
    PlanMarkArgs mk;
    mk.ids = d_ids; mk.offsets = d_offsets + lo; mk.pop = pop; mk.N = N_; mk.hasStart = hasStart_; mk.loops = loops_;
    mk.newKeys = d_newKeys_; mk.newCount = d_scalars_; mk.errFlag = d_scalars_ + 3; mk.used = nullptr;
    mk.active = d_active_ + lo; mk.coll = nullptr; mk.rowmap = d_rowmap_;

    int nNew = 0;
    for (int attempt = 0;; ++attempt) {
        groupBegin(KG_PLAN);
        IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_, 0, sizeof(int), stream_));
        if (wantUnion) {
            launch_fill_u8(d_used_, nKeys_, 0, stream_);
            mk.used = d_used_;
            nLaunches_++;
        }
        launch_mark_rows(mk, stream_);
        groupEnd(KG_PLAN);
        nLaunches_++;
        readScalars();
        nNew = h_scalars_[0];
        const int64_t freeRows = h_scalars_[1];
        // the heading search renders both candidate headings of a chunk of edges before it hands the losers' rows back
        const int64_t need = postProcessing_ ? (int64_t)nNew + std::min<int64_t>(nNew, headingChunk()) : nNew;
        if (need <= freeRows) break;
        launch_unclaim_rows(d_newKeys_, nNew, d_rowmap_, stream_);
        nLaunches_++;
        if (attempt == 0) {
            // evict what this range does not use
            PlanMarkArgs keep = mk;
            keep.rowmap = nullptr; keep.used = d_evictUsed();
            launch_fill_u8(keep.used, nKeys_, 0, stream_);
            launch_mark_rows(keep, stream_);
            IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 2, 0, sizeof(int), stream_));
            launch_gc_rows(d_rowmap_, keep.used, nKeys_, d_freeStack_, d_scalars_ + 1, d_scalars_ + 2, stream_);
            nLaunches_ += 3;
            liveRows_ = readInt(d_scalars_ + 2);
            nEvictions_++;
            continue;
        }
        if (pop == 1)
            throw std::runtime_error("libipp: the edge-bitmap cache (" + std::to_string(rowCapacity_) + " rows) cannot hold the " +
                                     std::to_string(need) + " edges of a single plan");
        const int64_t mid = lo + pop / 2;
        coverageRange(d_ids, d_offsets, lo, mid, memo, d_fitness, wantUnion);
        coverageRange(d_ids, d_offsets, mid, hi, memo, d_fitness, wantUnion);
        return;
    }
    if (nNew > 0) {
        if (postProcessing_) renderKeysHeadingSearch(d_newKeys_, nNew);
        else renderNewEdges(nNew, nullptr);
    }

    // 4. OR + area-weighted popcount
    PlanReduceArgs rd;
    rd.ids = d_ids; rd.offsets = d_offsets + lo; rd.pop = pop; rd.N = N_; rd.hasStart = hasStart_; rd.loops = loops_;
    rd.active = d_active_ + lo; rd.rowmap = d_rowmap_; rd.rows = d_rows_; rd.rowWords = rowWords_; rd.area = scene_.area;
    rd.nTiles = nTiles_; rd.totalArea = totalArea_; rd.fitness = d_fitness + 4 * lo; rd.partial = reducePartials(pop);
    groupBegin(KG_REDUCE);
    launch_plan_reduce(rd, numSMs_, stream_);
    groupEnd(KG_REDUCE);
    nLaunches_++;
    nReduceLaunches_++;
    nPlansReduced_ += pop;

    if (wantUnion) {
        launch_union_rows(d_used_, nKeys_, d_rowmap_, d_rows_, rowWords_, d_union_, stream_);
        nLaunches_++;
    }
    if (!memo && nNew > 0) {
        // memoization == false: edges rendered for this call are not remembered (evaluatePlanInternal,
        // PlanCoverageEstimator.cpp:128-151); their rows go straight back
        launch_unpublish_rows(d_newKeys_, nNew, d_rowmap_, d_freeStack_, d_scalars_ + 1, stream_);
        nLaunches_++;
        liveRows_ -= nNew;
    }
}

// ---------------------------------------------------------------------------------------------
// [id, angleOffset] genes (SIMPLIFIED_VIEW_ANGLE = False, PlanInterpreterBoxOrder.cpp:278-299: the heading of the edge INTO a
// viewpoint is turned by that viewpoint's offset, narrowed to float).  Energy, path length and collisions do not depend on the
// offsets and take the common path; the edge bitmaps do, so their memo is keyed by (edge key, offset bits) -- the reference's
// string key "prevID-currID,offset" (PlanCoverageEstimator.cpp:58-65) -- and lives on the host next to the integer-keyed one.
// Host-pointer entry points only.
// ---------------------------------------------------------------------------------------------
namespace {
inline uint64_t angle_key(int key, float angle) {
    uint32_t bits;
    memcpy(&bits, &angle, 4);
    return ((uint64_t)(uint32_t)key << 32) | bits;
}
}  // namespace

void Evaluator::evaluatePopulationAngles(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, bool memo, bool noLimit,
                                         const int32_t* d_ids, const int64_t* d_offsets, double* d_fitness, uint32_t* unionBitmap) {
    requireQuickHeading();
    const int64_t nGenes = offsets[pop];
    for (int64_t g = 0; g < nGenes; ++g)
        if (ids[g] < 0 || ids[g] >= N_) throw std::out_of_range("libipp: box id out of range");
    energyAndGate(d_ids, d_offsets, pop, noLimit, d_fitness);
    std::vector<uint8_t> active(pop);
    IPP_CUDA_CHECK(cudaMemcpyAsync(active.data(), d_active_, (size_t)pop, cudaMemcpyDeviceToHost, stream_));
    sync();
    if (unionBitmap) IPP_CUDA_CHECK(cudaMemsetAsync(d_union_, 0, (size_t)rowWords_ * 4, stream_));
    anglesRange(ids, angles, offsets, active.data(), 0, pop, memo, d_ids, d_offsets, d_fitness, unionBitmap != nullptr);
    if (unionBitmap) IPP_CUDA_CHECK(cudaMemcpyAsync(unionBitmap, d_union_, (size_t)bitmapWords() * 4, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void Evaluator::rowListToDevice(const std::vector<int>& list) {
    if ((int64_t)list.size() > rowListCap_) {
        cudaFree(d_rowList_);
        rowListCap_ = (int64_t)list.size() + (int64_t)list.size() / 4 + 16;
        d_rowList_ = dalloc<int>(rowListCap_);
    }
    if (!list.empty()) IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowList_, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
}

// coverage of the plans [lo, hi) of a population of [id, offset] genes (the gate has run: active[]).  A row pool too small for the
// new (edge, offset) pairs of the range first drops every memo entry the range does not use, then halves the range (see coverageRange).
void Evaluator::anglesRange(const int32_t* ids, const float* angles, const int64_t* offsets, const uint8_t* active, int64_t lo, int64_t hi,
                            bool memo, const int32_t* d_ids, const int64_t* d_offsets, double* d_fitness, bool wantUnion) {
    const int64_t pop = hi - lo;
    if (pop <= 0) return;
    std::vector<int64_t> edgeOffsets;
    std::vector<int> edgeRows;  // >= 0 row, < 0: -(slot + 1)
    std::vector<int> newKeys;
    std::vector<float> newAngles;
    std::unordered_map<uint64_t, int> slotOf;
    for (int attempt = 0;; ++attempt) {
        edgeOffsets.assign(pop + 1, 0);
        edgeRows.clear(); newKeys.clear(); newAngles.clear(); slotOf.clear();
        // the edges of the plans that passed the gate: memoised row, or a slot of this call's new (key, offset) pairs
        for (int64_t p = lo; p < hi; ++p) {
            const int64_t g0 = offsets[p], n = offsets[p + 1] - g0;
            if (active[p] && n > 0) {
                // node sequence [start] g_0 .. g_{n-1} [g_0 if the plan loops around]; the offset of an edge is its end node's
                int prev = hasStart_ ? N_ : ids[g0];
                const int64_t firstEnd = hasStart_ ? 0 : 1, ends = n + (loops_ ? 1 : 0);
                for (int64_t k = firstEnd; k < ends; ++k) {
                    const int64_t gk = k < n ? k : 0;
                    const int cur = ids[g0 + gk];
                    const float ang = angles[g0 + gk];
                    const int key = prev * N_ + cur;
                    const uint64_t mk = angle_key(key, ang);
                    auto it = angleMemo_.find(mk);
                    if (it != angleMemo_.end()) {
                        edgeRows.push_back(it->second);
                    } else {
                        auto ins = slotOf.emplace(mk, (int)newKeys.size());
                        if (ins.second) {
                            newKeys.push_back(key);
                            newAngles.push_back(ang);
                        }
                        edgeRows.push_back(-(ins.first->second + 1));
                    }
                    prev = cur;
                }
            }
            edgeOffsets[p - lo + 1] = (int64_t)edgeRows.size();
        }
        readScalars();
        const int64_t freeRows = h_scalars_[1];
        if ((int64_t)newKeys.size() <= freeRows) break;
        if (attempt == 0) {
            // drop what the range does not use: every integer-keyed row (a population of [id, offset] genes uses none) and the
            // (edge, offset) entries no plan of the range refers to
            {
                std::unordered_map<uint64_t, int> kept;
                for (int64_t p = lo; p < hi; ++p) {
                    const int64_t g0 = offsets[p], n = offsets[p + 1] - g0;
                    if (!active[p] || n <= 0) continue;
                    int prev = hasStart_ ? N_ : ids[g0];
                    const int64_t firstEnd = hasStart_ ? 0 : 1, ends = n + (loops_ ? 1 : 0);
                    for (int64_t k = firstEnd; k < ends; ++k) {
                        const int64_t gk = k < n ? k : 0;
                        const uint64_t mk = angle_key(prev * N_ + ids[g0 + gk], angles[g0 + gk]);
                        auto it = angleMemo_.find(mk);
                        if (it != angleMemo_.end()) kept.emplace(mk, it->second);
                        prev = ids[g0 + gk];
                    }
                }
                std::vector<int> freed;
                for (const auto& kv : angleMemo_)
                    if (!kept.count(kv.first)) freed.push_back(kv.second);
                angleMemo_.swap(kept);
                if (!freed.empty()) {
                    rowListToDevice(freed);
                    launch_release_rows(d_rowList_, (int)freed.size(), d_freeStack_, d_scalars_ + 1, stream_);
                    nLaunches_++;
                }
            }
            launch_fill_u8(d_evictUsed(), nKeys_, 0, stream_);
            IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 2, 0, sizeof(int), stream_));
            launch_gc_rows(d_rowmap_, d_evictUsed_, nKeys_, d_freeStack_, d_scalars_ + 1, d_scalars_ + 2, stream_);
            nLaunches_ += 2;
            liveRows_ = readInt(d_scalars_ + 2);
            headingReversed_.clear();
            nEvictions_++;
            continue;
        }
        if (pop == 1)
            throw std::runtime_error("libipp: the edge-bitmap cache (" + std::to_string(rowCapacity_) + " rows) cannot hold the " +
                                     std::to_string(newKeys.size()) + " edges of a single plan");
        const int64_t mid = lo + pop / 2;
        anglesRange(ids, angles, offsets, active, lo, mid, memo, d_ids, d_offsets, d_fitness, wantUnion);
        anglesRange(ids, angles, offsets, active, mid, hi, memo, d_ids, d_offsets, d_fitness, wantUnion);
        return;
    }
    const int nNew = (int)newKeys.size();
    // render the new pairs, at most nKeys_ at a time (the size of the key and offset staging buffers)
    std::vector<int> newRows(nNew);
    if (nNew > 0 && !d_angleBuf_) d_angleBuf_ = dalloc<float>(nKeys_);
    for (int c0 = 0; c0 < nNew; c0 += nKeys_) {
        const int n = std::min(nKeys_, nNew - c0);
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_newKeys_, newKeys.data() + c0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_angleBuf_, newAngles.data() + c0, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, stream_));
        std::vector<int> rows;
        renderKeys(d_newKeys_, n, d_angleBuf_, /*publish=*/false, &rows);
        std::copy(rows.begin(), rows.end(), newRows.begin() + c0);
    }
    for (int& r : edgeRows)
        if (r < 0) r = newRows[-r - 1];

    // OR + area-weighted popcount through the explicit row list
    const int64_t nEdges = (int64_t)edgeRows.size();
    if (nEdges + 1 > edgeRowsCap_) {
        cudaFree(d_edgeRows_);
        edgeRowsCap_ = nEdges + nEdges / 4 + 16;
        d_edgeRows_ = dalloc<int>(edgeRowsCap_);
    }
    if (pop + 2 > edgeOffCap_) {
        cudaFree(d_edgeOffsets_);
        edgeOffCap_ = pop + pop / 4 + 16;
        d_edgeOffsets_ = dalloc<int64_t>(edgeOffCap_);
    }
    if (nEdges > 0) IPP_CUDA_CHECK(cudaMemcpyAsync(d_edgeRows_, edgeRows.data(), (size_t)nEdges * sizeof(int), cudaMemcpyHostToDevice, stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_edgeOffsets_, edgeOffsets.data(), (size_t)(pop + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
    PlanReduceArgs rd;
    rd.ids = d_ids; rd.offsets = d_offsets + lo; rd.pop = pop; rd.N = N_; rd.hasStart = hasStart_; rd.loops = loops_;
    rd.active = d_active_ + lo; rd.rowmap = d_rowmap_; rd.edgeRows = d_edgeRows_; rd.edgeOffsets = d_edgeOffsets_;
    rd.rows = d_rows_; rd.rowWords = rowWords_; rd.area = scene_.area;
    rd.nTiles = nTiles_; rd.totalArea = totalArea_; rd.fitness = d_fitness + 4 * lo; rd.partial = reducePartials(pop);
    groupBegin(KG_REDUCE);
    launch_plan_reduce(rd, numSMs_, stream_);
    groupEnd(KG_REDUCE);
    nLaunches_++;
    nReduceLaunches_++;
    nPlansReduced_ += pop;

    if (wantUnion) {
        std::vector<int> uniq(edgeRows);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        rowListToDevice(uniq);
        launch_union_row_list(d_rowList_, (int)uniq.size(), d_rows_, rowWords_, d_union_, stream_);
        nLaunches_++;
    }
    sync();  // (the host vectors above feed asynchronous copies)
    if (memo) {
        for (const auto& kv : slotOf) angleMemo_[kv.first] = newRows[kv.second];
    } else if (nNew > 0) {
        // memoization == false: nothing rendered for this call is remembered (evaluatePlanInternal, PlanCoverageEstimator.cpp:128-151)
        rowListToDevice(newRows);
        launch_release_rows(d_rowList_, nNew, d_freeStack_, d_scalars_ + 1, stream_);
        nLaunches_++;
        sync();
    }
}

int64_t Evaluator::updateMemoisedEdgesAngles(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop) {
    // keys of the raw population: "start-first" only exists with a start location and the loop-closing edge is never part of the
    // set (PlanCoverageEstimator.cpp:284-310)
    std::unordered_map<uint64_t, int> keep;
    for (int64_t p = 0; p < pop; ++p) {
        const int64_t lo = offsets[p], n = offsets[p + 1] - lo;
        if (n <= 0) continue;
        int prev = hasStart_ ? N_ : ids[lo];
        for (int64_t k = hasStart_ ? 0 : 1; k < n; ++k) {
            const int cur = ids[lo + k];
            const uint64_t mk = angle_key(prev * N_ + cur, angles[lo + k]);
            auto it = angleMemo_.find(mk);
            if (it != angleMemo_.end()) keep.emplace(mk, it->second);
            prev = cur;
        }
    }
    std::vector<int> freed;
    for (const auto& kv : angleMemo_)
        if (!keep.count(kv.first)) freed.push_back(kv.second);
    angleMemo_.swap(keep);
    if (!freed.empty()) {
        if ((int64_t)freed.size() > rowListCap_) {
            cudaFree(d_rowList_);
            rowListCap_ = (int64_t)freed.size() + (int64_t)freed.size() / 4 + 16;
            d_rowList_ = dalloc<int>(rowListCap_);
        }
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowList_, freed.data(), freed.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
        launch_release_rows(d_rowList_, (int)freed.size(), d_freeStack_, d_scalars_ + 1, stream_);
        nLaunches_++;
        sync();
    }
    return (int64_t)angleMemo_.size();
}

void Evaluator::evaluatePopulationHost(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, bool memo,
                                       bool noLimit, double* fitness, uint32_t* unionBitmap) {
    bind();
    if (pop <= 0) return;
    const int64_t nGenes = offsets[pop] - offsets[0];
    if (offsets[0] != 0) throw std::invalid_argument("libipp: offsets[0] must be 0");
    ensurePlanBuffers(pop, nGenes);
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_ids_, ids, (size_t)nGenes * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_offsets_, offsets, (size_t)(pop + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
    if (angles) {
        evaluatePopulationAngles(ids, angles, offsets, pop, memo, noLimit, d_ids_, d_offsets_, d_fitness_, unionBitmap);
        IPP_CUDA_CHECK(cudaMemcpyAsync(fitness, d_fitness_, (size_t)pop * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        sync();
        return;
    }
    evaluatePopulationDevice(d_ids_, nullptr, d_offsets_, pop, nGenes, memo, noLimit, d_fitness_, unionBitmap);
    IPP_CUDA_CHECK(cudaMemcpyAsync(fitness, d_fitness_, (size_t)pop * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    sync();
}

int64_t Evaluator::updateMemoisedEdges(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop) {
    bind();
    const int64_t nGenes = pop > 0 ? offsets[pop] : 0;
    for (int64_t g = 0; g < nGenes; ++g)
        if (ids[g] < 0 || ids[g] >= N_) throw std::out_of_range("libipp: box id out of range");
    // One memo in the reference, two here: a population of [id, offset] genes keeps none of the integer-keyed edges and the
    // other way round.
    const int64_t zero = 0;
    const int64_t keptAngles = angles ? updateMemoisedEdgesAngles(ids, angles, offsets, pop) : updateMemoisedEdgesAngles(nullptr, nullptr, &zero, 0);
    const int64_t popInt = angles ? 0 : pop;
    ensurePlanBuffers(std::max<int64_t>(popInt, 1), std::max<int64_t>(nGenes, 1));
    launch_fill_u8(d_used_, nKeys_, 0, stream_);
    if (popInt > 0) {
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_ids_, ids, (size_t)nGenes * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_offsets_, offsets, (size_t)(popInt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
        // keys of the raw population: "start-first" only exists with a start location and the loop-closing edge is
        // never part of the set (PlanCoverageEstimator.cpp:284-310)
        PlanMarkArgs mk;
        mk.ids = d_ids_; mk.offsets = d_offsets_; mk.pop = popInt; mk.N = N_; mk.hasStart = hasStart_; mk.loops = false;
        mk.active = nullptr; mk.coll = nullptr; mk.rowmap = nullptr; mk.used = d_used_; mk.newKeys = d_newKeys_; mk.newCount = d_scalars_;
        mk.errFlag = d_scalars_ + 3;
        launch_mark_rows(mk, stream_);
        nLaunches_++;
    }
    IPP_CUDA_CHECK(cudaMemsetAsync(d_scalars_ + 2, 0, sizeof(int), stream_));
    launch_gc_rows(d_rowmap_, d_used_, nKeys_, d_freeStack_, d_scalars_ + 1, d_scalars_ + 2, stream_);
    nLaunches_ += 2;
    liveRows_ = readInt(d_scalars_ + 2);
    return liveRows_ + keptAngles;
}

// ---------------------------------------------------------------------------------------------
// postProcessing = true: calculateViewingDirectionTowardsPrimitives (OsgHelpers.cpp:451-489).  The heading of an edge keeps the
// axis of the quick rule but its side is chosen by looking: both candidates are rendered along the edge (K4, the second one
// through the +infinity sentinel of sensor_direction), K6 turns each bitmap into its coverage score (pseudo plans of one edge),
// the lower score wins and a tie keeps the quick heading.  The winner's bitmap IS the bitmap of the edge (the reference renders
// it once more with the same heading); the loser's row goes back to the free stack.
// ---------------------------------------------------------------------------------------------
void Evaluator::renderKeysHeadingSearch(const int* d_keys, int nNew) {
    if (nNew <= 0) return;
    if (!d_sortedKeys_) d_sortedKeys_ = dalloc<int>(nKeys_);
    if (!d_angleBuf_) d_angleBuf_ = dalloc<float>(nKeys_);
    const int chunkMax = headingChunk();
    std::vector<int> keys(nNew);
    IPP_CUDA_CHECK(cudaMemcpyAsync(keys.data(), d_keys, (size_t)nNew * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    sync();
    for (int c0 = 0; c0 < nNew; c0 += chunkMax) {
        const int C = std::min(chunkMax, nNew - c0), C2 = 2 * C;
        std::vector<int> keys2(C2);
        std::vector<float> ang2(C2);
        for (int i = 0; i < C; ++i) {
            keys2[2 * i] = keys2[2 * i + 1] = keys[c0 + i];
            ang2[2 * i] = 0.f;
            ang2[2 * i + 1] = INFINITY;
        }
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_sortedKeys_, keys2.data(), (size_t)C2 * sizeof(int), cudaMemcpyHostToDevice, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_angleBuf_, ang2.data(), (size_t)C2 * sizeof(float), cudaMemcpyHostToDevice, stream_));
        std::vector<int> rows;
        renderKeys(d_sortedKeys_, C2, d_angleBuf_, /*publish=*/false, &rows);
        // coverage score of every candidate bitmap: C2 pseudo plans of one edge each through K6
        if (C2 > pseudoCap_) {
            cudaFree(d_pseudoOffsets_); cudaFree(d_pseudoEdgeOffsets_); cudaFree(d_pseudoActive_); cudaFree(d_pseudoFitness_);
            pseudoCap_ = C2 + C2 / 4 + 16;
            d_pseudoOffsets_ = dalloc<int64_t>(pseudoCap_ + 1);
            d_pseudoEdgeOffsets_ = dalloc<int64_t>(pseudoCap_ + 1);
            d_pseudoActive_ = dalloc<uint8_t>(pseudoCap_);
            d_pseudoFitness_ = dalloc<double>((size_t)pseudoCap_ * 4);
            std::vector<int64_t> off(pseudoCap_ + 1), eoff(pseudoCap_ + 1);
            for (int i = 0; i <= pseudoCap_; ++i) { off[i] = 2 * (int64_t)i; eoff[i] = i; }
            IPP_CUDA_CHECK(cudaMemcpyAsync(d_pseudoOffsets_, off.data(), off.size() * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
            IPP_CUDA_CHECK(cudaMemcpyAsync(d_pseudoEdgeOffsets_, eoff.data(), eoff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, stream_));
            launch_fill_u8(d_pseudoActive_, pseudoCap_, 1, stream_);
            sync();
        }
        if (C2 + 1 > edgeRowsCap_) {
            cudaFree(d_edgeRows_);
            edgeRowsCap_ = C2 + C2 / 4 + 16;
            d_edgeRows_ = dalloc<int>(edgeRowsCap_);
        }
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_edgeRows_, rows.data(), (size_t)C2 * sizeof(int), cudaMemcpyHostToDevice, stream_));
        PlanReduceArgs rd;
        rd.ids = d_sortedKeys_; rd.offsets = d_pseudoOffsets_; rd.pop = C2; rd.N = N_; rd.hasStart = false; rd.loops = false;
        rd.active = d_pseudoActive_; rd.rowmap = d_rowmap_; rd.edgeRows = d_edgeRows_; rd.edgeOffsets = d_pseudoEdgeOffsets_;
        rd.rows = d_rows_; rd.rowWords = rowWords_; rd.area = scene_.area;
        rd.nTiles = nTiles_; rd.totalArea = totalArea_; rd.fitness = d_pseudoFitness_; rd.partial = reducePartials(C2);
        launch_plan_reduce(rd, numSMs_, stream_);
        nLaunches_++;
        std::vector<double> fit((size_t)C2 * 4);
        IPP_CUDA_CHECK(cudaMemcpyAsync(fit.data(), d_pseudoFitness_, fit.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        sync();
        std::vector<int> winners(C), losers(C);
        for (int i = 0; i < C; ++i) {
            const double quick = fit[(size_t)(2 * i) * 4], reversed = fit[(size_t)(2 * i + 1) * 4];
            const bool rev = reversed < quick;  // lower score = more surface seen; a tie keeps the quick heading
            winners[i] = rows[2 * i + (rev ? 1 : 0)];
            losers[i] = rows[2 * i + (rev ? 0 : 1)];
            headingReversed_[keys[c0 + i]] = rev;
        }
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowOfSlot_, winners.data(), (size_t)C * sizeof(int), cudaMemcpyHostToDevice, stream_));
        launch_publish_rows(d_keys + c0, d_rowOfSlot_, C, d_rowmap_, stream_);
        if ((int64_t)C > rowListCap_) {
            cudaFree(d_rowList_);
            rowListCap_ = C + C / 4 + 16;
            d_rowList_ = dalloc<int>(rowListCap_);
        }
        IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowList_, losers.data(), (size_t)C * sizeof(int), cudaMemcpyHostToDevice, stream_));
        launch_release_rows(d_rowList_, C, d_freeStack_, d_scalars_ + 1, stream_);
        nLaunches_ += 2;
        sync();
        liveRows_ += C;
    }
}

// headings of keys the memo does not hold yet (interpretAndExportPlan of a post-processing evaluator)
void Evaluator::decideHeadings(const std::vector<int>& keys) {
    std::vector<int> todo;
    for (int k : keys)
        if (!headingReversed_.count(k) && std::find(todo.begin(), todo.end(), k) == todo.end()) todo.push_back(k);
    if (todo.empty()) return;
    bind();
    readScalars();
    if ((int)todo.size() * 2 > h_scalars_[1]) throw std::runtime_error("libipp: edge-bitmap cache full; call updateMemoisedEdges");
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_newKeys_, todo.data(), todo.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
    renderKeysHeadingSearch(d_newKeys_, (int)todo.size());
}

// InterpretPlan, PlanInterpreterBoxOrder.cpp:301-324 (host twin of the device decode; used for export only)
void Evaluator::interpretPlan(const double* genes, int n, int width, double* pos, double* dirs, int* nPos, int* nDirs) {
    *nPos = 0;
    *nDirs = 0;
    if (n == 0) return;
    D3 center = d3(sceneCenter_[0], sceneCenter_[1], sceneCenter_[2]);
    std::vector<int> idx;  // position sequence as box indices (N_ = start location)
    if (hasStart_) idx.push_back(N_);
    for (int i = 0; i < n; ++i) {
        size_t id = (size_t)(genes[(size_t)i * width] + 0.5);
        if (id >= (size_t)N_) throw std::out_of_range("libipp: box id out of range");
        idx.push_back((int)id);
    }
    if (postProcessing_) {
        if (width > 1) throw std::invalid_argument("libipp: [id, angle] genes are not supported together with postProcessing = true");
        std::vector<int> keys;
        for (size_t k = 1; k < idx.size(); ++k) keys.push_back(idx[k - 1] * N_ + idx[k]);
        decideHeadings(keys);
    }
    std::vector<D3> p;
    if (hasStart_) p.push_back(d3(start_[0], start_[1], start_[2]));
    int nd = 0;
    for (int i = 0; i < n; ++i) {
        const int id = idx[i + (hasStart_ ? 1 : 0)];
        float off = width > 1 ? (float)genes[(size_t)i * width + 1] : 0.f;
        D3 cur = d3(grid_.centres[3 * (size_t)id], grid_.centres[3 * (size_t)id + 1], grid_.centres[3 * (size_t)id + 2]);
        if (!p.empty()) {
            if (postProcessing_) {
                const int prevIdx = idx[i + (hasStart_ ? 1 : 0) - 1];
                off = headingReversed_.at(prevIdx * N_ + id) ? INFINITY : 0.f;
            }
            D3 h = sensor_direction(p.back(), cur, center, off);
            dirs[3 * nd] = h.x; dirs[3 * nd + 1] = h.y; dirs[3 * nd + 2] = h.z;
            nd++;
        }
        p.push_back(cur);
    }
    for (size_t i = 0; i < p.size(); ++i) { pos[3 * i] = p[i].x; pos[3 * i + 1] = p[i].y; pos[3 * i + 2] = p[i].z; }
    *nPos = (int)p.size();
    *nDirs = nd;
}

// ---------------------------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------------------------
// What a post-processing evaluator does not do: [id, angle] genes, the per-edge test hook and the multi-GPU edge sharing keep the
// quick heading rule and refuse instead of answering with it.
void Evaluator::requireQuickHeading() const {
    if (postProcessing_)
        throw std::logic_error("libipp: not available with postProcessing = true (heading search towards the primitives): [id, angle] "
                               "genes, ipp_edge_bitmap and the multi-GPU edge sharing use the quick heading rule");
}

void Evaluator::edgeBitmap(int prev, int curr, double angle, uint32_t* out) {
    bind();
    requireQuickHeading();
    if (prev == -1) {
        if (!hasStart_) throw std::invalid_argument("libipp: no start location");
        prev = N_;
    }
    if (prev < 0 || prev > N_ || curr < 0 || curr >= N_) throw std::out_of_range("libipp: box id out of range");
    int key = prev * N_ + curr;
    float a = (float)angle;
    // rendered into the scratch row (index rowCapacity_), the memo is left untouched
    int scratch = (int)rowCapacity_;
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_newKeys_, &key, sizeof(int), cudaMemcpyHostToDevice, stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowOfSlot_, &scratch, sizeof(int), cudaMemcpyHostToDevice, stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_angleOne_, &a, sizeof(float), cudaMemcpyHostToDevice, stream_));
    renderBatch(d_newKeys_, d_angleOne_, 1, d_rowOfSlot_);
    IPP_CUDA_CHECK(cudaMemcpyAsync(out, d_rows_ + (size_t)scratch * rowWords_, (size_t)bitmapWords() * 4, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void Evaluator::edgesCollide(const double* edges, int64_t n, uint8_t* out) {
    bind();
    if (n <= 0) return;
    double* d_e = dalloc<double>((size_t)n * 6);
    uint8_t* d_o = dalloc<uint8_t>(n);
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_e, edges, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, stream_));
    launch_collide_xyz(scene_, d_e, n, d_o, stream_);
    nLaunches_++;
    IPP_CUDA_CHECK(cudaMemcpyAsync(out, d_o, n, cudaMemcpyDeviceToHost, stream_));
    sync();
    cudaFree(d_e);
    cudaFree(d_o);
}

void Evaluator::boxesHitMesh(const double* centres, int64_t n, double half, uint8_t* out) {
    bind();
    if (n <= 0) return;
    double* d_c = dalloc<double>((size_t)n * 3);
    uint8_t* d_o = dalloc<uint8_t>(n);
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_c, centres, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
    launch_boxes_hit(scene_, d_c, n, half, d_o, stream_);
    nLaunches_++;
    IPP_CUDA_CHECK(cudaMemcpyAsync(out, d_o, n, cudaMemcpyDeviceToHost, stream_));
    sync();
    cudaFree(d_c);
    cudaFree(d_o);
}

void Evaluator::renderIds(const double* eye, const double* center, const double* up, int32_t* out) {
    bind();
    ensureSnapBuffers(2);
    int32_t* d_img = dalloc<int32_t>((size_t)cam_.W * cam_.H);
    launch_single_snapshot(d_snaps_, eye, center, up, cam_, scene_, d_ntiles_, stream_);
    renderSnapshots(1, d_img);
    int scratch = (int)rowCapacity_;
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_rowOfSlot_, &scratch, sizeof(int), cudaMemcpyHostToDevice, stream_));
    launch_pack_rows(d_flags_, 1, scene_.Tpad, d_rowOfSlot_, d_rows_, rowWords_, stream_);  // clears the flags
    nLaunches_ += 2;
    IPP_CUDA_CHECK(cudaMemcpyAsync(out, d_img, (size_t)cam_.W * cam_.H * 4, cudaMemcpyDeviceToHost, stream_));
    sync();
    cudaFree(d_img);
}

void Evaluator::leafTable(const double* eye, const double* center, const double* up, uint64_t* out, int64_t cap, int64_t* count,
                          uint32_t* mask32, int32_t* outFirst, int32_t* outCount) {
    bind();
    if (!useBins_) throw std::logic_error("libipp: the leaf-bin variant is not in force for this evaluator");
    ensureSnapBuffers(2);
    launch_single_snapshot(d_snaps_, eye, center, up, cam_, scene_, d_ntiles_, stream_);
    launch_leaf_table(scene_, cam_, d_snaps_, 1, bins_, stream_);
    nLaunches_ += 2;
    int n = readInt(bins_.tableCount);
    *count = n;
    int64_t m = std::min<int64_t>(n, cap);
    if (m > 0) IPP_CUDA_CHECK(cudaMemcpyAsync(out, bins_.table, (size_t)m * 8, cudaMemcpyDeviceToHost, stream_));
    if (mask32) IPP_CUDA_CHECK(cudaMemcpyAsync(mask32, bins_.tileMask, 32 * 4, cudaMemcpyDeviceToHost, stream_));
    sync();
    if (outFirst || outCount) {  // the triangles of every entry's cluster (the entry carries the cluster index)
        std::vector<float4> lo(scene_.nLeaves), hi(scene_.nLeaves);
        IPP_CUDA_CHECK(cudaMemcpyAsync(lo.data(), scene_.leafLo, lo.size() * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
        IPP_CUDA_CHECK(cudaMemcpyAsync(hi.data(), scene_.leafHi, hi.size() * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
        sync();
        for (int64_t k = 0; k < m; ++k) {
            const uint32_t ref = (uint32_t)(out[k] & 0x0fffffffu);
            if (outFirst) memcpy(outFirst + k, &lo[ref].w, 4);
            if (outCount) memcpy(outCount + k, &hi[ref].w, 4);
        }
    }
}

void Evaluator::bvhOrder(int32_t* out) {
    bind();
    std::vector<float4> a(scene_.T);
    if (scene_.T > 0) IPP_CUDA_CHECK(cudaMemcpyAsync(a.data(), scene_.ta, (size_t)scene_.T * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
    sync();
    for (int i = 0; i < scene_.T; ++i) memcpy(out + i, &a[i].w, 4);
}

void Evaluator::bvhSelftest(const float* rays, int64_t n, int32_t* outIds, int64_t* mismatches, int32_t* outBrute, float* outT) {
    bind();
    *mismatches = 0;
    if (n <= 0) return;
    float* d_r = dalloc<float>((size_t)n * 6);
    int32_t* d_a = dalloc<int32_t>(n);
    int32_t* d_b = dalloc<int32_t>(n);
    float* d_t = outT ? dalloc<float>((size_t)n * 2) : nullptr;
    IPP_CUDA_CHECK(cudaMemcpyAsync(d_r, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, stream_));
    launch_bvh_selftest(scene_, d_r, n, d_a, d_b, d_t, stream_);
    nLaunches_++;
    std::vector<int32_t> b(n);
    IPP_CUDA_CHECK(cudaMemcpyAsync(outIds, d_a, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    IPP_CUDA_CHECK(cudaMemcpyAsync(b.data(), d_b, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    if (outT) IPP_CUDA_CHECK(cudaMemcpyAsync(outT, d_t, (size_t)n * 8, cudaMemcpyDeviceToHost, stream_));
    sync();
    for (int64_t i = 0; i < n; ++i)
        if (outIds[i] != b[i]) (*mismatches)++;
    if (outBrute) memcpy(outBrute, b.data(), (size_t)n * 4);
    cudaFree(d_r);
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_t);
}

}  // namespace ipp
