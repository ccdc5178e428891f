// evaluator.h -- host-side owner of one B200-resident plan evaluator (the object behind ipp_t).
// Mirrors PlanCoverageEstimator (plan_evaluator/Evaluator/src/PlanCoverageEstimator.h:51-165).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "planner.h"

namespace ipp {

struct Comm;  // comm.cpp

// What renderEdgesShared needs from the communicator (implemented on NCCL in comm.cu).
struct Collectives {
    int rank = 0, world = 1;
    virtual ~Collectives() {}
    // in place: every rank contributes bytesPerRank bytes at d_buf + rank * bytesPerRank and receives all slots
    virtual void allGatherInPlace(void* d_buf, size_t bytesPerRank, cudaStream_t st) = 0;
    virtual void allReduceMax(double* inout, int n, cudaStream_t st) = 0;
};

enum KernelGroup { KG_REDUCE = 0, KG_VISIBILITY = 1, KG_COLLISION = 2, KG_PACK = 3, KG_PLAN = 4, KG_TABLE = 5, KG_COUNT = 6 };

class Evaluator {
   public:
    Evaluator(const std::string& scenePath, const double sensorSpecs[8], bool postProcessing, const double* startXYZ, bool planLoopsAround,
              int device);
    ~Evaluator();

    // ---- reference API ----
    int numBoxes() const { return N_; }
    int64_t numTriangles() const { return scene_.T; }
    int64_t bitmapWords() const { return (scene_.T + 31) / 32; }
    double maxAllowedEnergy() const { return maxAllowedEnergy_; }
    const std::vector<Plan>& simplePlans() const { return completePlans_; }
    void evaluatePopulationHost(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, bool memo, bool noLimit,
                                double* fitness, uint32_t* unionBitmap);
    void evaluatePopulationDevice(const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t pop, int64_t nGenes,
                                  bool memo, bool noLimit, double* d_fitness, uint32_t* unionBitmapHost);
    // Multi-GPU: the edges the WHOLE population needs are rendered once across the ranks (contiguous chunks of the sorted new
    // keys), the fresh bitmap rows are all-gathered, and every rank's memo ends up complete.  Every rank must pass the same
    // population and hold the same memo contents (checked).  Returns the number of edges this rank rendered itself.
    // number of edge uses of this (device-resident) population whose bitmap the memo does not hold yet (energy gate applied)
    int64_t countUnseenEdges(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit);
    int64_t renderEdgesShared(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit, Collectives& cx);
    int64_t updateMemoisedEdges(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop);
    int64_t memoSize() const { return liveRows_ + (int64_t)angleMemo_.size(); }
    int64_t evictions() const { return nEvictions_; }  // times the row pool ran full and rows of unused edges were dropped
    void clearMemo();
    void interpretPlan(const double* genes, int n, int width, double* pos, double* dirs, int* nPos, int* nDirs);

    // ---- test hooks ----
    void edgeBitmap(int prev, int curr, double angle, uint32_t* out);
    void edgesCollide(const double* edges, int64_t n, uint8_t* out);
    void boxesHitMesh(const double* centres, int64_t n, double half, uint8_t* out);
    void renderIds(const double* eye, const double* center, const double* up, int32_t* out);
    void leafTable(const double* eye, const double* center, const double* up, uint64_t* out, int64_t cap, int64_t* count, uint32_t* mask32,
                   int32_t* outFirst = nullptr, int32_t* outCount = nullptr);
    void bvhOrder(int32_t* out);
    void bvhSelftest(const float* rays, int64_t n, int32_t* outIds, int64_t* mismatches, int32_t* outBrute = nullptr, float* outT = nullptr);

    // ---- plumbing ----
    void sync() { IPP_CUDA_CHECK(cudaStreamSynchronize(stream_)); }
    void timerStart();
    double timerStop();
    void flushL2();
    double readBandwidth(int64_t bytes, int reps);
    void counters(int64_t out[8], bool reset);
    void visibilityStats(int64_t out[8], bool reset);
    void kernelTime(int which, double* ms, int64_t* launches, bool reset);
    void setProfiling(bool on) { profiling_ = on; }
    bool usesBins() const { return useBins_; }
    void setUseBins(bool on);  // K4 variant: leaf bins (default when supported) or BVH walk per tile
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    void bind() const { IPP_CUDA_CHECK(cudaSetDevice(device_)); }

    // scene facts
    double totalArea_ = 0, collisionPenalty_ = 0;
    double sceneCenter_[3] = {0, 0, 0};
    float bbmin_[3], bbmax_[3];
    std::vector<double> areaHost_;
    ViewpointGrid grid_;
    CameraParams cam_;
    bool postProcessing_ = false, hasStart_ = false, loops_ = false;
    double start_[3] = {0, 0, 0};
    Comm* comm_ = nullptr;

   private:
    void renderNewEdges(int nNew, const float* d_angles);
    void renderKeys(const int* d_keys, int n, const float* d_angles, bool publish = true, std::vector<int>* rowsOut = nullptr);
    void energyAndGate(const int32_t* d_ids, const int64_t* d_offsets, int64_t pop, bool noLimit, double* d_fitness);
    void coverageRange(const int32_t* d_ids, const int64_t* d_offsets, int64_t lo, int64_t hi, bool memo, double* d_fitness, bool wantUnion);
    int headingChunk() const;
    uint8_t* d_evictUsed();
    void evaluatePopulationAngles(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, bool memo, bool noLimit,
                                  const int32_t* d_ids, const int64_t* d_offsets, double* d_fitness, uint32_t* unionBitmap);
    void anglesRange(const int32_t* ids, const float* angles, const int64_t* offsets, const uint8_t* active, int64_t lo, int64_t hi, bool memo,
                     const int32_t* d_ids, const int64_t* d_offsets, double* d_fitness, bool wantUnion);
    void rowListToDevice(const std::vector<int>& list);
    int64_t updateMemoisedEdgesAngles(const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop);
    void renderBatch(const int* d_keys, const float* d_angles, int n, const int* d_rowOfSlot);
    void renderSnapshots(int nSnaps, int32_t* d_idImage);
    void requireQuickHeading() const;
    double* reducePartials(int64_t pop);  // K6 scratch for populations too small to fill the GPU (nullptr otherwise)
    // postProcessing = true: both candidate headings of every key are rendered, the better one becomes the memo row
    void renderKeysHeadingSearch(const int* d_keys, int n);
    void decideHeadings(const std::vector<int>& keys);
    void ensurePlanBuffers(int64_t pop, int64_t nGenes);
    void ensureSnapBuffers(int64_t nSnaps);
    int readInt(const int* d);
    void readScalars();
    void groupBegin(int g);
    void groupEnd(int g);
    void harvestTimings(bool wait);
    void exclusiveScan(const int* d_in, int* d_out, int n);

    int device_ = 0, numSMs_ = 148;
    cudaStream_t stream_ = nullptr;
    DeviceScene scene_;
    int N_ = 0;          // candidate viewpoints
    int nKeys_ = 0;      // (N+1) * N directed edge keys, row N of the position table = start location
    double maxAllowedEnergy_ = -1;
    std::vector<Plan> levelPlans_, completePlans_;

    // device tables
    double* d_pos_ = nullptr;      // (N+1) x 3
    uint8_t* d_coll_ = nullptr;    // nKeys (padded to 4)
    int* d_rowmap_ = nullptr;      // nKeys
    uint8_t* d_used_ = nullptr;    // nKeys
    uint8_t* d_evictUsed_ = nullptr;  // nKeys, made on the first eviction
    int* d_newKeys_ = nullptr;     // nKeys
    int* d_scalars_ = nullptr;     // [0] newCount, [1] freeTop, [2] live, [3] scratch
    int* h_scalars_ = nullptr;     // pinned mirror
    uint32_t* d_rows_ = nullptr;   // row pool: (capacity + 1) rows, the last one is scratch
    int64_t rowWords_ = 0, rowCapacity_ = 0;
    int* d_freeStack_ = nullptr;
    int64_t liveRows_ = 0;
    int nTiles_ = 0;

    // render batch buffers
    int batchCap_ = 0;
    uint8_t* d_flags_ = nullptr;
    int *d_rowOfSlot_ = nullptr, *d_counts_ = nullptr, *d_snapOffset_ = nullptr;
    Snapshot* d_snaps_ = nullptr;
    // unique camera poses of a batch and the edge rows each of them serves (pose_dedupe.cu)
    Snapshot* d_snapsU_ = nullptr;
    int *d_ntilesU_ = nullptr, *d_slotList_ = nullptr, *d_dedupeIWork_ = nullptr;
    unsigned long long* d_dedupeWork_ = nullptr;
    void* d_dedupeTmp_ = nullptr;
    size_t dedupeTmpBytes_ = 0;
    bool poseDedupe_ = true;
    int64_t nSnapshotsRendered_ = 0;
    int *d_ntiles_ = nullptr, *d_tileOffset_ = nullptr;
    int64_t snapCap_ = 0;
    void* d_scanTmp_ = nullptr;
    size_t scanTmpBytes_ = 0;
    unsigned long long* d_visCounters_ = nullptr;  // [0] rays, [1] work cursor, [2] nodes fetched, [3] triangles staged, [4] tiles, [5] bin entries read, [6..7] diagnostics, [8] supertiles from records, [9] supertiles by leaf walk
    // leaf-bin variant of K4 (visibility_bins.cu)
    bool useBins_ = false;
    BinBuffers bins_;
    int* d_superOffset_ = nullptr;
    int* d_countsAll_ = nullptr;  // snapshots per new edge (nKeys + 1)
    // [id, angleOffset] genes: the memo of their edges lives on the host, key = (edge key << 32) | float bits of the offset
    std::unordered_map<uint64_t, int> angleMemo_;
    float* d_angleBuf_ = nullptr;   // nKeys
    int* d_edgeRows_ = nullptr;
    int64_t* d_edgeOffsets_ = nullptr;
    int* d_rowList_ = nullptr;
    int64_t edgeRowsCap_ = 0, edgeOffCap_ = 0, rowListCap_ = 0;
    // postProcessing = true: which keys ended up with the reversed heading (calculateViewingDirectionTowardsPrimitives)
    std::unordered_map<int, bool> headingReversed_;
    int64_t* d_pseudoOffsets_ = nullptr;
    int64_t* d_pseudoEdgeOffsets_ = nullptr;
    uint8_t* d_pseudoActive_ = nullptr;
    double* d_pseudoFitness_ = nullptr;
    int pseudoCap_ = 0;
    double* d_partial_ = nullptr;
    int64_t partialCap_ = 0;
    // shared rendering across ranks
    int* d_sortedKeys_ = nullptr;  // nKeys
    void* d_sortTmp_ = nullptr;
    size_t sortTmpBytes_ = 0;
    uint32_t* d_exchange_ = nullptr;  // world x slab x rowWords
    size_t exchangeWords_ = 0;
    uint8_t* d_activeShared_ = nullptr;
    double* d_fitShared_ = nullptr;
    int64_t sharedCap_ = 0;
    int64_t maxSnapsPerBatch_ = 1 << 20;
    bool memoDiverged_ = false;  // the shared rendering was skipped once: the ranks' memos may differ, the next shared call clears them
    float* d_angleOne_ = nullptr;

    // plan buffers (host-pointer entry point)
    int32_t* d_ids_ = nullptr;
    int64_t* d_offsets_ = nullptr;
    double* d_fitness_ = nullptr;
    uint8_t* d_active_ = nullptr;
    int64_t popCap_ = 0, geneCap_ = 0, activeCap_ = 0;
    uint32_t* d_union_ = nullptr;

    // timing
    cudaEvent_t evStart_ = nullptr, evStop_ = nullptr;
    struct EvPair { cudaEvent_t a, b; int group; };
    std::vector<EvPair> pendingEv_, freeEv_;
    double groupMs_[KG_COUNT] = {0, 0, 0, 0, 0, 0};
    int64_t groupLaunches_[KG_COUNT] = {0, 0, 0, 0, 0, 0};
    bool profiling_ = true;
    uint32_t* d_flushBuf_ = nullptr;
    int64_t flushWords_ = 0;
    // counters
    int64_t nLaunches_ = 0, nSnapshots_ = 0, nEdgesRendered_ = 0, nCollisionQueries_ = 0, nReduceLaunches_ = 0, nPlansReduced_ = 0,
            nRowsRead_ = 0, nEvictions_ = 0;
};


}  // namespace ipp
