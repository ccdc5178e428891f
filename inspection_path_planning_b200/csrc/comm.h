// comm.h -- K7: NCCL plumbing for the population-sharded multi-GPU path (one process per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipp {
class Evaluator;
struct Comm;
void comm_unique_id(uint8_t out128[128]);
Comm* comm_create(int rank, int world, const uint8_t id128[128], cudaStream_t st);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_world(const Comm* c);
// op: 0 sum, 1 max; host buffer in/out
void comm_allreduce(Comm* c, double* inout, int n, int op, cudaStream_t st);
// evaluate plans [lo, hi) of the population on this rank and all-gather `per` fitness rows per rank
void comm_eval_and_allgather(Evaluator& ev, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, int64_t per,
                             int64_t lo, int64_t hi, bool memo, bool noLimit, double* fitness);
// device-resident form: this rank's `mine` plans (d_ids / d_offsets on the device) are evaluated into slot `rank` of
// d_gather (world * per rows of 4 doubles) and the slots are all-gathered in place; asynchronous on the evaluator's stream
void comm_eval_device_and_allgather(Evaluator& ev, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t mine,
                                    int64_t nGenes, int64_t per, bool memo, bool noLimit, double* d_gather, bool shardBad = false);
}  // namespace ipp
