// comm.cu -- see comm.h.  NCCL is loaded with dlopen so that libipp.so has no link-time dependency on it;
// the population is sharded by individual, the mesh and BVH are replicated per GPU, and the only exchange
// step is one ncclAllGather of the fitness rows per generation.  plan_reduce writes each
// rank's rows straight into its slot of the gather buffer (in-place all-gather, no staging copy).
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "evaluator.h"

namespace ipp {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& api() {
    static NcclApi a;
    if (!a.lib) {
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) throw std::runtime_error(std::string("libipp: cannot load NCCL: ") + dlerror());
        auto sym = [&](const char* n) {
            void* p = dlsym(a.lib, n);
            if (!p) throw std::runtime_error(std::string("libipp: NCCL symbol missing: ") + n);
            return p;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    }
    return a;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string("libipp: NCCL error in ") + what + ": " + api().GetErrorString(r));
}
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* d_small = nullptr;  // all-reduce scratch
    double* d_gather = nullptr; // world * per * 4 fitness doubles
    int64_t gatherCap = 0;
    int32_t* d_ids = nullptr;
    int64_t* d_offsets = nullptr;
    int64_t idsCap = 0, offCap = 0;
    float* d_angles = nullptr;
    int64_t anglesCap = 0;
    int32_t* d_allIds = nullptr;  // whole population (shared edge rendering)
    int64_t* d_allOffsets = nullptr;
    int64_t allIdsCap = 0, allOffCap = 0;
    bool shareEdges = true;
};

void comm_unique_id(uint8_t out128[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, 128);
}

Comm* comm_create(int rank, int world, const uint8_t id128[128], cudaStream_t st) {
    (void)st;
    Comm* c = new Comm;
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    nccl_check(api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    IPP_CUDA_CHECK(cudaMalloc(&c->d_small, 64 * sizeof(double)));
    if (const char* e = getenv("IPP_SHARE_EDGES")) c->shareEdges = atoi(e) != 0;
    return c;
}
void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm) api().CommDestroy(c->comm);
    cudaFree(c->d_small);
    cudaFree(c->d_gather);
    cudaFree(c->d_ids);
    cudaFree(c->d_offsets);
    cudaFree(c->d_angles);
    cudaFree(c->d_allIds);
    cudaFree(c->d_allOffsets);
    delete c;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }

void comm_allreduce(Comm* c, double* inout, int n, int op, cudaStream_t st) {
    if (n > 64) throw std::invalid_argument("libipp: all-reduce of at most 64 doubles");
    IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_small, inout, n * sizeof(double), cudaMemcpyHostToDevice, st));
    nccl_check(api().AllReduce(c->d_small, c->d_small, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum, c->comm, st), "ncclAllReduce");
    IPP_CUDA_CHECK(cudaMemcpyAsync(inout, c->d_small, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    IPP_CUDA_CHECK(cudaStreamSynchronize(st));
}

namespace {
struct NcclCollectives : Collectives {
    Comm* c;
    explicit NcclCollectives(Comm* cc) : c(cc) {
        rank = cc->rank;
        world = cc->world;
    }
    void allGatherInPlace(void* d_buf, size_t bytesPerRank, cudaStream_t st) override {
        nccl_check(api().AllGather((const char*)d_buf + (size_t)rank * bytesPerRank, d_buf, bytesPerRank, ncclChar, c->comm, st), "ncclAllGather (rows)");
    }
    void allReduceMax(double* inout, int n, cudaStream_t st) override { comm_allreduce(c, inout, n, 1, st); }
};
}  // namespace

void comm_eval_device_and_allgather(Evaluator& ev, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t mine,
                                    int64_t nGenes, int64_t per, bool memo, bool noLimit, double* d_gather, bool shardBad) {
    Comm* c = ev.comm_;
    cudaStream_t st = ev.stream();
    if (mine > per) throw std::invalid_argument("libipp: shard larger than the per-rank slot");
    double* d_mine = d_gather + (size_t)c->rank * per * 4;
    // A failure on one rank only (a bad box id in its shard, an allocation) must not leave the others waiting in the collective:
    // the ranks agree on an error flag first and throw together.
    std::string localError;
    if (shardBad) localError = "libipp: box id out of range";  // found on the host by the caller: nothing is evaluated
    else try {
        // rows of the last (short) shard that hold no plan are zero
        if (mine < per) IPP_CUDA_CHECK(cudaMemsetAsync(d_mine + mine * 4, 0, (size_t)(per - mine) * 4 * sizeof(double), st));
        // plan_reduce writes this rank's rows straight into its slot of the gather buffer: no staging copy before the collective
        if (mine > 0) ev.evaluatePopulationDevice(d_ids, d_angles, d_offsets, mine, nGenes, memo, noLimit, d_mine, nullptr);
    } catch (const std::exception& e) {
        localError = e.what();
        if (localError.empty()) localError = "unknown error";
    }
    double failed = localError.empty() ? 0.0 : 1.0;
    comm_allreduce(c, &failed, 1, 1, st);
    if (failed != 0.0) {
        if (localError.find("out of range") != std::string::npos) throw std::out_of_range(localError);
        throw std::runtime_error(localError.empty() ? "libipp: another rank failed while it evaluated its shard of the population" : localError);
    }
    // the one exchange step: every rank receives every shard's fitness rows over NVLink
    nccl_check(api().AllGather(d_mine, d_gather, (size_t)per * 4, ncclDouble, c->comm, st), "ncclAllGather");
}

void comm_eval_and_allgather(Evaluator& ev, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, int64_t per,
                             int64_t lo, int64_t hi, bool memo, bool noLimit, double* fitness) {
    Comm* c = ev.comm_;
    cudaStream_t st = ev.stream();
    const int64_t mine = hi - lo;
    const int64_t g0 = offsets[lo], nGenes = offsets[hi] - g0;
    // A bad box id (assert at PlanInterpreterBoxOrder.cpp:282) must make every rank raise, not leave one of them outside a
    // collective.  A rank checks its own shard here (1/world of the genes); the whole population is only checked -- by every rank,
    // on the same data, so they agree -- before it is marked for the shared rendering below.
    const int N = ev.numBoxes();
    auto anyOutOfRange = [N](const int32_t* p, int64_t n) {
        int bad = 0;  // branch-free: the compiler vectorises it (3 M genes: 0.6 ms)
        for (int64_t k = 0; k < n; ++k) bad += (uint32_t)p[k] >= (uint32_t)N;
        return bad != 0;
    };
    const bool shardBad = mine > 0 && anyOutOfRange(ids + g0, nGenes);
    if (c->world * per * 4 > c->gatherCap) {
        cudaFree(c->d_gather);
        c->gatherCap = c->world * per * 4 + 1024;
        IPP_CUDA_CHECK(cudaMalloc(&c->d_gather, (size_t)c->gatherCap * sizeof(double)));
    }
    if (nGenes + 1 > c->idsCap) {
        cudaFree(c->d_ids);
        c->idsCap = nGenes + nGenes / 4 + 16;
        IPP_CUDA_CHECK(cudaMalloc(&c->d_ids, (size_t)c->idsCap * sizeof(int32_t)));
    }
    if (mine + 2 > c->offCap) {
        cudaFree(c->d_offsets);
        c->offCap = mine + mine / 4 + 16;
        IPP_CUDA_CHECK(cudaMalloc(&c->d_offsets, (size_t)c->offCap * sizeof(int64_t)));
    }
    std::vector<int64_t> off((size_t)mine + 1);
    for (int64_t i = 0; i <= mine; ++i) off[i] = offsets[lo + i] - g0;
    if (angles && nGenes + 1 > c->anglesCap) {
        cudaFree(c->d_angles);
        c->anglesCap = nGenes + nGenes / 4 + 16;
        IPP_CUDA_CHECK(cudaMalloc(&c->d_angles, (size_t)c->anglesCap * sizeof(float)));
    }
    if (mine > 0) {
        IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_ids, ids + g0, (size_t)nGenes * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_offsets, off.data(), (size_t)(mine + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (angles) IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_angles, angles + g0, (size_t)nGenes * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    // The edges of the WHOLE population are rendered once across the ranks and exchanged (second exchange step of this path: the
    // fresh bitmap rows), so that a rank does not render an edge its shard shares with another rank's shard.  Skipped when no
    // shard uses an edge its rank has not got: the steady state of a memoised run costs one probe kernel and one all-reduce.
    // (memoization == false keeps the reference's meaning -- nothing rendered for this call is remembered -- and renders per shard)
    // ([id, offset] genes and post-processing evaluators render per shard: the first keeps its memo on the host, keyed by the
    // offset bits; the second chooses every heading by rendering both candidates.)
    if (c->world > 1 && pop > 0 && c->shareEdges && memo && !angles && !ev.postProcessing_) {
        double v[2] = {shardBad ? 0.0 : (double)ev.countUnseenEdges(c->d_ids, c->d_offsets, mine, noLimit), shardBad ? 1.0 : 0.0};
        comm_allreduce(c, v, 2, 1, st);
        if (v[1] != 0.0) throw std::out_of_range("libipp: box id out of range");
        if (v[0] > 0) {
            const int64_t allGenes = offsets[pop];
            if (anyOutOfRange(ids, allGenes)) throw std::out_of_range("libipp: box id out of range");
            if (allGenes + 1 > c->allIdsCap) {
                cudaFree(c->d_allIds);
                c->allIdsCap = allGenes + allGenes / 4 + 16;
                IPP_CUDA_CHECK(cudaMalloc(&c->d_allIds, (size_t)c->allIdsCap * sizeof(int32_t)));
            }
            if (pop + 2 > c->allOffCap) {
                cudaFree(c->d_allOffsets);
                c->allOffCap = pop + pop / 4 + 16;
                IPP_CUDA_CHECK(cudaMalloc(&c->d_allOffsets, (size_t)c->allOffCap * sizeof(int64_t)));
            }
            IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_allIds, ids, (size_t)allGenes * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            IPP_CUDA_CHECK(cudaMemcpyAsync(c->d_allOffsets, offsets, (size_t)(pop + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
            NcclCollectives cx(c);
            ev.renderEdgesShared(c->d_allIds, c->d_allOffsets, pop, noLimit, cx);
        }
    }
    comm_eval_device_and_allgather(ev, c->d_ids, angles ? c->d_angles : nullptr, c->d_offsets, mine, nGenes, per, memo, noLimit, c->d_gather,
                                   shardBad);
    IPP_CUDA_CHECK(cudaMemcpyAsync(fitness, c->d_gather, (size_t)pop * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    IPP_CUDA_CHECK(cudaStreamSynchronize(st));
}

}  // namespace ipp
