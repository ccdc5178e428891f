// common.cuh -- shared declarations of the libipp.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#define IPP_CUDA_CHECK(expr)                                                                            \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " +   \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");       \
    } while (0)

namespace ipp {

// Compiled-in constants of the reference: Utility_Functions/src/Constants.h:19-60
constexpr double kNumViewpointsInsideVolume = 1000.0;
constexpr double kBoxSafetyMargin = 2.0;
constexpr double kBoxPadding = 4.0;
constexpr double kMinDistFromBottom = 3.0;
constexpr bool kDownwardCameraActive = true;
constexpr double kTranslationEnergyFactor = 0.1;
constexpr double kRotationEnergyFactor = 1.0;
constexpr double kMaxEnergyMultiplier = 1.0;
constexpr double kCollisionPenaltyModifier = 2.0;
constexpr double kEdgeSafetyBuffer = 1.5;
constexpr double kCameraFarPlaneDist = 10.0;

constexpr int kLeafMax = 8;          // triangles per BVH leaf
constexpr int kLeafShift = 4;        // leaf link: ~link = (first << kLeafShift) | count
constexpr int kLeafMask = 15;
constexpr int kMaxBinLeaves = 16384;    // clusters a camera pose can sort (visibility_bins.cu)
constexpr int kMaxClusterTris = 256;    // largest cluster of the binned visibility kernel (maximal subtrees average ~0.6 of it)
constexpr int kMaxDepth = 96;        // Karras tree over 63-bit Morton codes + 31-bit index tie-break: depth <= 96
constexpr int kTileWords = 128;      // bitmap words per plan_reduce tile (one uint4 per lane)
constexpr int kTileTris = kTileWords * 32;

// One 64-byte BVH node: the boxes of both children and their links.
//   q0 = (c0min.x, c0min.y, c0min.z, c0max.x)  q1 = (c0max.y, c0max.z, c1min.x, c1min.y)
//   q2 = (c1min.z, c1max.x, c1max.y, c1max.z)  q3 = (link0, link1, -, -) as int bits
// link >= 0: internal node index; link < 0: leaf, ~link = (first << kLeafShift) | count, count in 0..kLeafMax
struct BvhNode {
    float4 q0, q1, q2, q3;
};

struct DeviceScene {
    int T = 0;        // triangles
    int Tpad = 0;     // T rounded up to 32
    int nNodes = 0;   // BVH nodes (root = 0)
    // BVH-ordered triangles, canonical (lexicographic) vertex order; a.w = original id as int bits
    float4 *ta = nullptr, *tb = nullptr, *tc = nullptr;
    float4* tn = nullptr;  // geometric normal (B-A) x (C-A) of the canonical order, computed in fp64
    // BVH-ordered triangles, file vertex order (exact collision predicate); a.w = original id
    float4 *oa = nullptr, *ob = nullptr, *oc = nullptr;
    BvhNode* nodes = nullptr;
    // cluster table of the binned visibility kernel: maximal BVH subtrees of at most leafCap triangles (leafCap = kLeafMax: the leaves
    // of the node array), node order.  Padded box; leafLo.w = first triangle (BVH order), leafHi.w = count, both as int bits
    float4 *leafLo = nullptr, *leafHi = nullptr;
    int nLeaves = 0, leafCap = 0;
    double* area = nullptr;   // [areaPad] file order, zero padded to a multiple of kTileTris
    float bbmin[3], bbmax[3]; // float bounding box of all vertices
};

// Per-camera-pose record consumed by the visibility kernel
struct Snapshot {
    double f[3], s[3], u[3];  // look-at basis (world space)
    float eye[3];
    int slot;                 // as made by K3: flag row (edge) within the current batch; as seen by K4: first entry of this pose's
    int slotCount;            // flag rows in the batch's slot list, and how many there are (pose_dedupe.cu: edges that share a pose)
    int rx0, ry0, rnx, rny;   // rectangle of 32x32 macro tiles that may see the scene
};
static_assert(sizeof(Snapshot) <= 128, "the visibility kernels copy a Snapshot with one 4-byte load per lane");

struct CameraParams {
    double tanX, tanY;
    float zNear, zFar;
    int W, H;
    bool front, down;
    double sampleInterval;
};

}  // namespace ipp
