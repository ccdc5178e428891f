// kernels.h -- launch wrappers of every CUDA kernel in libipp.so (one per translation unit group).
#pragma once
#include "common.cuh"

namespace ipp {

struct PlanMarkArgs {
    const int* ids;
    const int64_t* offsets;
    int64_t pop;
    int N;
    bool hasStart, loops;
    const uint8_t* active;  // nullable: only plans with active[p] != 0
    uint8_t* coll;          // collision flags per key: 0xFF unknown, 0xFE pending, 0 / 1
    int* rowmap;            // bitmap row per key: -1 unknown, -2 pending, >= 0 row
    uint8_t* used;          // nullable: set to 1 for every key touched
    int* newKeys;
    int* newCount;
    int* errFlag;  // set to 1 when a box id is out of range
};

struct PlanEnergyArgs {
    const int* ids;
    const int64_t* offsets;
    int64_t pop;
    int N;
    bool hasStart, loops;
    const double* pos;  // (N+1) x 3, row N = start location
    const uint8_t* coll;
    double collisionPenalty, maxEnergy;
    bool disableLimit;
    double* fitness;  // pop x 4
    uint8_t* active;  // pop
};

struct PlanReduceArgs {
    const int* ids;
    const int64_t* offsets;
    int64_t pop;
    int N;
    bool hasStart, loops;
    const uint8_t* active;
    const int* rowmap;
    const uint32_t* rows;  // row pool
    int64_t rowWords;      // words per row (multiple of 4)
    const double* area;    // padded to nTiles * kTileTris
    int nTiles;
    double totalArea;
    double* fitness;
    // optional scratch of pop * ceil(nTiles / 8) doubles: lets a very small population on a large mesh split the tiles of a plan over
    // several CTAs (area sums per block of 8 tiles, added up in block order by a second kernel)
    double* partial = nullptr;
    const int* edgeRows = nullptr;         // optional ([id, angle] genes, heading search): row of edge e of plan p at
    const int64_t* edgeOffsets = nullptr;  // edgeRows[edgeOffsets[p] + e] instead of rowmap[key]
};

// exact.cu (-fmad=false)
void launch_mesh_prepare(const float* d_verts, int T, double* d_area, uint8_t* d_degenerate, float* d_canon, float* d_centroid, cudaStream_t st);
void launch_collide_keys(const DeviceScene& sc, const int* d_keys, int n, const double* d_pos, int N, uint8_t* d_coll, cudaStream_t st);
void launch_collide_xyz(const DeviceScene& sc, const double* d_edges, int64_t n, uint8_t* d_out, cudaStream_t st);
void launch_boxes_hit(const DeviceScene& sc, const double* d_centres, int64_t n, double s, uint8_t* d_out, cudaStream_t st);
void launch_count_snapshots(const int* d_keys, int n, const double* d_pos, int N, const CameraParams& cam, int* d_counts, cudaStream_t st);
void launch_fill_snapshots(const int* d_keys, const float* d_angles, int n, const double* d_pos, int N, const double center[3],
                           const CameraParams& cam, const DeviceScene& sc, const int* d_snapOffset, Snapshot* d_snaps, int* d_ntiles,
                           cudaStream_t st);
void launch_single_snapshot(Snapshot* d_snap, const double eye[3], const double center[3], const double up[3], const CameraParams& cam,
                            const DeviceScene& sc, int* d_ntiles, cudaStream_t st);
void launch_mark_collisions(const PlanMarkArgs& a, cudaStream_t st);
void launch_energy(const PlanEnergyArgs& a, cudaStream_t st);
void launch_mark_rows(const PlanMarkArgs& a, cudaStream_t st);
// counts (into a.newCount) the edges of the active plans that have no bitmap row yet, without claiming them
void launch_probe_rows(const PlanMarkArgs& a, cudaStream_t st);
void launch_gc_rows(int* d_rowmap, const uint8_t* d_used, int nKeys, int* d_freeStack, int* d_freeTop, int* d_live, cudaStream_t st);
void launch_assign_rows(const int* d_keys, int n, int* d_rowmap, int* d_rowsOut, const int* d_freeStack, int* d_freeTop, int scratchRow,
                        int* d_errFlag, cudaStream_t st);
void launch_unclaim_rows(const int* d_keys, int n, int* d_rowmap, cudaStream_t st);
void launch_unpublish_rows(const int* d_keys, int n, int* d_rowmap, int* d_freeStack, int* d_freeTop, cudaStream_t st);
void launch_publish_rows(const int* d_keys, const int* d_rows, int n, int* d_rowmap, cudaStream_t st);
// bitmap rows of n keys <-> a contiguous staging buffer (n x rowWords), for the exchange of freshly rendered rows between ranks
void launch_rows_gather(const int* d_keys, int n, const int* d_rowmap, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_out, cudaStream_t st);
void launch_rows_scatter(const uint32_t* d_in, const int* d_rowOfSlot, int n, uint32_t* d_rows, int64_t rowWords, cudaStream_t st);

// lbvh.cu
// Builds the BVH over the non-degenerate triangles.  d_verts / d_canon: T x 9 floats (file / canonical vertex order).
void build_bvh(const float* d_verts, const float* d_canon, const float* d_centroid, const uint8_t* d_degenerate, int T,
               const float bbmin[3], const float bbmax[3], DeviceScene& sc, cudaStream_t st);
void free_bvh(DeviceScene& sc);

// visibility.cu
// Casts one ray through every pixel centre of the macro tiles listed for the snapshots and sets flags[slot * Tpad + id] = 1
// for the nearest triangle.  d_tileOffset = exclusive scan of tiles per snapshot, totalTiles its sum.
void launch_visibility(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, const int* d_tileOffset, int nSnaps,
                       int64_t totalTiles, uint8_t* d_flags, const int* d_slotList, int32_t* d_idImage /*nullable test hook*/,
                       unsigned long long* d_rayCounter, cudaStream_t st);
// flags[slot][Tpad] (bytes) -> bitmap rows; clears the flags
void launch_pack_rows(uint8_t* d_flags, int nSlots, int Tpad, const int* d_rowOfSlot, uint32_t* d_rows, int64_t rowWords, cudaStream_t st);
void launch_bvh_selftest(const DeviceScene& sc, const float* d_rays, int64_t n, int32_t* d_idsBvh, int32_t* d_idsBrute, float* d_t /*nullable: (tBvh, tBrute) per ray*/,
                         cudaStream_t st);
int visibility_grid_blocks();

// visibility_bins.cu -- the same sampling with depth-sorted leaf bins (meshes of <= kMaxBinLeaves leaves, images <= 1024 a side)
struct BinBuffers {
    unsigned long long* table = nullptr;      // [poses][stride] packed leaf entries, front to back
    int* tableCount = nullptr;                // [poses] entries per pose
    uint32_t* tileMask = nullptr;             // [poses][32] occupied 32x32 tiles, one word per tile row
    unsigned long long* superMask = nullptr;  // [poses] occupied 128x128 supertiles (bit = sy * 8 + sx)
    int* nSuper = nullptr;                    // [poses + 1] popcount of superMask
    unsigned long long* scratch = nullptr;    // [gridBlocks * warps][stride] per-warp supertile lists
    float4* records = nullptr;                // [gridBlocks * warps][bins_record_floats_per_warp() / 4] set-up triangles of a supertile
    int stride = 0, gridBlocks = 0;
};
bool bins_supported(const DeviceScene& sc, const CameraParams& cam);
int bins_grid_blocks();
int bins_warps_per_block();
size_t bins_record_floats_per_warp();
int bins_table_stride(int nLeaves);
void launch_leaf_table(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, int nSnaps, const BinBuffers& b,
                       cudaStream_t st);
// d_superOffset = exclusive scan of BinBuffers::nSuper (nSnaps + 1 entries), totalSuper its last entry
void launch_visibility_bins(const DeviceScene& sc, const CameraParams& cam, const Snapshot* d_snaps, const int* d_superOffset, int nSnaps,
                            int64_t totalSuper, const BinBuffers& b, uint8_t* d_flags, const int* d_slotList, int32_t* d_idImage,
                            unsigned long long* d_counters, cudaStream_t st);
// pose_dedupe.cu -- camera poses several edges of a batch share are rendered once (Snapshot::slot / slotCount -> d_slotList)
size_t pose_dedupe_tmp_bytes(int n);
void launch_pose_dedupe(const Snapshot* d_snaps, const int* d_ntiles, int n, bool enabled, Snapshot* d_out, int* d_ntilesOut, int* d_slotList,
                        int* d_nUniq, unsigned long long* d_work, int* d_iwork, void* d_tmp, size_t tmpBytes, cudaStream_t st);
size_t sort_tmp_bytes(int n);
// the same pseudo-random order of a key set on every rank (contiguous chunks of it are random samples of the set)
void scrambled_order_i32(const int* d_in, int* d_out, int n, void* d_tmp, size_t tmpBytes, cudaStream_t st);
size_t scan_tmp_bytes(int n);
void exclusive_scan_i32(const int* d_in, int* d_out, int n, void* d_tmp, size_t tmpBytes, cudaStream_t st);

// plan_reduce.cu (-fmad=false)
void launch_plan_reduce(const PlanReduceArgs& a, int numSMs, cudaStream_t st);
// which streaming kernel serves ordinary populations: -1 = the library's choice, 0 = plain loads, 1.. = bulk-copy configurations
void set_plan_reduce_variant(int v);
const char* plan_reduce_variant_name(int v);
void launch_union_rows(const uint8_t* d_used, int nKeys, const int* d_rowmap, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_union,
                       cudaStream_t st);
// OR of an explicit list of rows -> union bitmap; rows back to the free stack
void launch_union_row_list(const int* d_rowList, int n, const uint32_t* d_rows, int64_t rowWords, uint32_t* d_union, cudaStream_t st);
void launch_release_rows(const int* d_rowList, int n, int* d_freeStack, int* d_freeTop, cudaStream_t st);
void launch_fill_u8(uint8_t* p, int64_t n, uint8_t v, cudaStream_t st);
void launch_fill_i32(int* p, int64_t n, int v, cudaStream_t st);
void launch_iota_i32(int* p, int64_t n, cudaStream_t st);
void launch_flush(uint32_t* p, int64_t nWords, cudaStream_t st);
void launch_read_probe(const uint32_t* p, int64_t nWords, int reps, unsigned* d_sink, int numSMs, cudaStream_t st);

}  // namespace ipp
