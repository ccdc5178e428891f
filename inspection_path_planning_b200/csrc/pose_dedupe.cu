// pose_dedupe.cu -- camera poses that several plan edges share are rendered once.
//
// The reference renders every sample of every edge (CameraEstimator.cpp:131-178).  Two edges that meet at a viewpoint with the
// same sensor heading -- A->B and B->A whenever their headings agree, every edge into or out of a viewpoint along the same
// horizontal direction (the heading only depends on the xy direction of the edge, OsgHelpers.cpp:421-446, and the candidate grid
// stacks viewpoints above each other) -- put bit-identical camera poses there: same float eye, same fp64 look-at basis, hence the
// same item buffer.  13 % of the poses of BASELINE config 2's random populations and ~30 % of a saturated edge set are such
// repeats.  Here the poses of a batch are hashed, sorted (CUB radix sort, stable) and compared bit for bit; K4 renders one
// representative per group, in the order of first appearance, and its winners are flagged in the rows of all the edges of the
// group (Snapshot::slot = first entry of the group in `slotList`, Snapshot::slotCount entries).  Results are unchanged bit for bit.
#include <cub/cub.cuh>

#include "kernels.h"

namespace ipp {

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}

// the fields that determine the image: eye (3 floats) and the look-at basis (9 doubles)
__device__ __forceinline__ bool same_pose(const Snapshot& a, const Snapshot& b) {
    bool eq = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        eq = eq && __float_as_int(a.eye[k]) == __float_as_int(b.eye[k]);
        eq = eq && __double_as_longlong(a.f[k]) == __double_as_longlong(b.f[k]);
        eq = eq && __double_as_longlong(a.s[k]) == __double_as_longlong(b.s[k]);
        eq = eq && __double_as_longlong(a.u[k]) == __double_as_longlong(b.u[k]);
    }
    return eq;
}

__global__ void k_pose_hash(const Snapshot* __restrict__ snaps, int n, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Snapshot& s = snaps[i];
    unsigned long long h = 0x243F6A8885A308D3ull;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        h = mix64(h, (unsigned long long)(unsigned)__float_as_int(s.eye[k]));
        h = mix64(h, (unsigned long long)__double_as_longlong(s.f[k]));
        h = mix64(h, (unsigned long long)__double_as_longlong(s.s[k]));
        h = mix64(h, (unsigned long long)__double_as_longlong(s.u[k]));
    }
    keys[i] = h;
    idx[i] = i;
}

// sorted position i starts a group unless the pose before it (same hash, stable sort: lower original index) is the same pose
__global__ void k_pose_heads(const Snapshot* __restrict__ snaps, const int* __restrict__ idx, int n, int* __restrict__ headPos) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = i == 0 || !same_pose(snaps[idx[i]], snaps[idx[i - 1]]);
    headPos[i] = head ? i : 0;  // an inclusive max-scan turns this into the position of the group's head
}

// per sorted position: the edge row of this member; the last member of a group tells the group's leader (the member with the
// lowest original index = the head) where the group starts in slotList and how long it is
__global__ void k_pose_groups(const Snapshot* __restrict__ snaps, const int* __restrict__ idx, const int* __restrict__ headPos, int n,
                              int* __restrict__ slotList, int* __restrict__ leaderFlag, int* __restrict__ groupStart, int* __restrict__ groupCount) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = idx[i];
    slotList[i] = snaps[j].slot;
    const int hp = headPos[i];
    leaderFlag[j] = hp == i ? 1 : 0;
    const bool last = i == n - 1 || headPos[i + 1] == i + 1;
    if (last) {
        const int leader = idx[hp];
        groupStart[leader] = hp;
        groupCount[leader] = i - hp + 1;
    }
}

__global__ void k_pose_emit(const Snapshot* __restrict__ snaps, const int* __restrict__ leaderFlag, const int* __restrict__ uid,
                            const int* __restrict__ groupStart, const int* __restrict__ groupCount, const int* __restrict__ ntilesIn, int n,
                            Snapshot* __restrict__ out, int* __restrict__ ntilesOut, int* __restrict__ nUniq) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (leaderFlag[j]) {
        Snapshot s = snaps[j];
        s.slot = groupStart[j];
        s.slotCount = groupCount[j];
        out[uid[j]] = s;
        ntilesOut[uid[j]] = ntilesIn[j];
    }
    if (j == n - 1) *nUniq = uid[j] + leaderFlag[j];
}

__global__ void k_pose_identity(const Snapshot* __restrict__ snaps, const int* __restrict__ ntilesIn, int n, Snapshot* __restrict__ out,
                                int* __restrict__ ntilesOut, int* __restrict__ slotList, int* __restrict__ nUniq) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Snapshot s = snaps[j];
    slotList[j] = s.slot;
    s.slot = j;
    s.slotCount = 1;
    out[j] = s;
    ntilesOut[j] = ntilesIn[j];
    if (j == 0) *nUniq = n;
}

struct MaxOp {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

}  // namespace

size_t pose_dedupe_tmp_bytes(int n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, n);
    cub::DeviceScan::InclusiveScan(nullptr, b, (const int*)nullptr, (int*)nullptr, MaxOp(), n);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const int*)nullptr, (int*)nullptr, n);
    return std::max(a, std::max(b, c)) + 256;
}

// d_work: 2 n 64-bit words; d_iwork: 7 n ints.  Writes the unique poses (order of first appearance), their tile counts (BVH-walk
// variant), the slot list and the number of unique poses (*d_nUniq).
void launch_pose_dedupe(const Snapshot* d_snaps, const int* d_ntiles, int n, bool enabled, Snapshot* d_out, int* d_ntilesOut, int* d_slotList,
                        int* d_nUniq, unsigned long long* d_work, int* d_iwork, void* d_tmp, size_t tmpBytes, cudaStream_t st) {
    if (n <= 0) return;
    const int B = 256, G = (n + B - 1) / B;
    if (!enabled) {
        k_pose_identity<<<G, B, 0, st>>>(d_snaps, d_ntiles, n, d_out, d_ntilesOut, d_slotList, d_nUniq);
        return;
    }
    unsigned long long *keys = d_work, *keys2 = d_work + n;
    int *idx = d_iwork, *idx2 = d_iwork + n, *headPos = d_iwork + 2 * (size_t)n, *leaderFlag = d_iwork + 3 * (size_t)n,
        *groupStart = d_iwork + 4 * (size_t)n, *groupCount = d_iwork + 5 * (size_t)n, *uid = d_iwork + 6 * (size_t)n;
    k_pose_hash<<<G, B, 0, st>>>(d_snaps, n, keys, idx);
    cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, keys, keys2, idx, idx2, n, 0, 64, st);
    k_pose_heads<<<G, B, 0, st>>>(d_snaps, idx2, n, headPos);
    cub::DeviceScan::InclusiveScan(d_tmp, tmpBytes, headPos, headPos, MaxOp(), n, st);
    k_pose_groups<<<G, B, 0, st>>>(d_snaps, idx2, headPos, n, d_slotList, leaderFlag, groupStart, groupCount);
    cub::DeviceScan::ExclusiveSum(d_tmp, tmpBytes, leaderFlag, uid, n, st);
    k_pose_emit<<<G, B, 0, st>>>(d_snaps, leaderFlag, uid, groupStart, groupCount, d_ntiles, n, d_out, d_ntilesOut, d_nUniq);
}

}  // namespace ipp
