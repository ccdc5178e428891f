/* ipp.h -- C ABI of libipp.so, the B200-native plan-fitness evaluator.
 *
 * Drop-in boundary for the SWIG-exposed evaluator of kaiolae/Inspection-Path-Planning
 * (plan_optimizer/cpp_wrapper/cpp_binding.i:1-40).  Every entry point names the
 * reference interface it replaces.  Plain C types only: pointers, sizes, ints,
 * doubles.  All functions returning int return 0 on success and non-zero on
 * error; the message is available from ipp_last_error() (thread local).
 *
 * There is no CPU fallback: every compute call needs a CUDA device (sm_100a).
 *
 * Conventions
 *   - a "plan" is a sequence of genes; a gene is [boxId] (width 1) or
 *     [boxId, angleOffsetRad] (width 2), exactly as in the reference
 *     (plan_evaluator/Evaluator/src/PlanInterpreterBoxOrder.cpp:278-299).
 *   - populations travel in CSR form: `ids` (int32, one per gene), optional
 *     `angles` (float, one per gene, NULL = width-1 genes) and `offsets`
 *     (int64, pop+1 entries).
 *   - fitness rows are 4 doubles: {coverageScore, energy, pathLength,
 *     collidingEdges}.  Columns 0-1 are the reference's evaluatePlan() pair
 *     (PlanCoverageEstimator.cpp:201-282); columns 2-3 are extra.
 *   - seen-triangle bitmaps are uint32 words, bit (id & 31) of word (id >> 5),
 *     id = triangle index in mesh file order (the reference's colour ID,
 *     Utility_Functions/src/TriangleData.cpp:23-39).
 */
#ifndef IPP_H_
#define IPP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipp_evaluator ipp_t;

/* plotting_style enum, cpp_binding.i:20 */
enum { IPP_PLOT_NORMAL = 0, IPP_PLOT_CIRCULATING = 1, IPP_PLOT_IMAGE = 2, IPP_PLOT_NOTHING = 3 };

const char* ipp_last_error(void);
/* number of CUDA devices visible (0 if none / no driver) */
int ipp_device_count(void);

/* PlanCoverageEstimator::PlanCoverageEstimator  (cpp_binding.i:29,
 * PlanCoverageEstimator.cpp:167-195).  sensor_specs[8] = {sampleInterval, fov_x,
 * fov_y, imageHeight, zNear, zFar, frontCam, downCam} (CameraEstimator.h:83-92).
 * start_xyz may be NULL; it is copied (the reference borrows the pointer).
 * device = CUDA ordinal.  Loads .stl / .obj (and the flat .ippm test format),
 * builds the BVH, the candidate-viewpoint grid, the circling seed plans and the
 * energy limit. */
int ipp_create(const char* scene_path, const double* sensor_specs, int post_processing, const double* start_xyz,
               int plan_loops_around, int printer_friendly, int device, ipp_t** out);
/* ~PlanCoverageEstimator */
void ipp_destroy(ipp_t* h);

/* getNumberOfBoxes (cpp_binding.i:32) */
int ipp_num_boxes(const ipp_t* h);
/* SceneKeeper::getTriangleCount */
int64_t ipp_num_triangles(const ipp_t* h);
/* words per seen-triangle bitmap = ceil(T/32) */
int64_t ipp_bitmap_words(const ipp_t* h);
/* getMaxAllowedEnergy (cpp_binding.i:36); -1 when post_processing */
double ipp_max_allowed_energy(const ipp_t* h);
/* scene constants derived at load (SceneKeeper.cpp:65-71, TriangleData.cpp:65-72) */
double ipp_total_area(const ipp_t* h);
double ipp_collision_penalty(const ipp_t* h);
void ipp_scene_info(const ipp_t* h, double* center3, double* bbox6);
void ipp_image_size(const ipp_t* h, int* width, int* height);
/* The mesh ingest on its own (host code, no device needed): osgDB::readNodeFile + TriangleFunctor enumeration
 * (SceneKeeper.cpp:73-86, :29): binary / ASCII .stl, .obj (Y-up -> Z-up, polygons fan-split) and the flat .ippm test format ->
 * 9 floats per triangle in file order.  Call with out_Tx9 == NULL to obtain the count.  A file that cannot be read is an error
 * (std::invalid_argument in the reference, SceneKeeper.cpp:79-82). */
int ipp_load_mesh(const char* scene_path, float* out_Tx9, int64_t cap_triangles, int64_t* n_triangles);
/* per-triangle Heron areas, file order (TriangleData.cpp:56-61) */
int ipp_areas(const ipp_t* h, double* out_T);
/* candidate viewpoint centres, N x 3 (PlanInterpreterBoxOrder::divideIntoBoxes :193-251) */
int ipp_box_centres(const ipp_t* h, double* out_Nx3);
/* boxVolume[z][x][y] classification grid: -1 too close, -2 nothing visible, else box id */
void ipp_grid_dims(const ipp_t* h, int* nz, int* nx, int* ny);
int ipp_grid(const ipp_t* h, int32_t* out);

/* getSimplePlans (cpp_binding.i:34; PlanInterpreterBoxOrder.cpp:111-160) in CSR form.
 * Call with ids == NULL to obtain the sizes. */
int ipp_simple_plans(ipp_t* h, int* n_plans, int* n_genes, double* ids, int32_t* offsets);

/* evaluatePlan (cpp_binding.i:31).  genes: n x width doubles, width 1 ([id]) or 2 ([id, angleOffsetRad],
 * PlanInterpreterBoxOrder.cpp:278-299).  how_to_plot must be IPP_PLOT_NOTHING (the GL viewers are out of
 * scope).  out4 receives one fitness row. */
int ipp_evaluate_plan(ipp_t* h, const double* genes, int n, int width, int memoization, int how_to_plot,
                      int disable_energy_limit, double* out4);

/* The one batched entry point: evaluatePlan over a whole population
 * (replaces toolbox.map(evaluate, ...) at plan_optimizer/optimize_coverage.py:141,191).
 * Host pointers; copies in and out are part of the call.  angles: NULL for [id] genes, otherwise one heading
 * offset per gene (float, as the reference narrows it); edges with offsets are memoised under (previous id, id,
 * offset).  union_bitmap may be NULL; otherwise it receives the OR of the seen sets of all plans that passed
 * the energy gate (W words). */
int ipp_evaluate_population(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop,
                            int memoization, int disable_energy_limit, double* fitness_popx4, uint32_t* union_bitmap);

/* Same computation with every buffer already resident in device memory
 * (d_* are CUDA device pointers obtained from ipp_device_alloc).  Asynchronous
 * on the evaluator's stream; call ipp_sync() before reading results. */
int ipp_evaluate_population_device(ipp_t* h, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets,
                                   int64_t pop, int64_t n_genes, int memoization, int disable_energy_limit,
                                   double* d_fitness_popx4);

/* updateMemoisedEdges (cpp_binding.i:33; PlanCoverageEstimator.cpp:284-327):
 * drop every cached edge bitmap whose key is not in the population; returns the
 * surviving count in *remaining. */
int ipp_update_memoised_edges(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop,
                              int64_t* remaining);
/* number of edge bitmaps currently memoised */
int64_t ipp_memo_size(const ipp_t* h);
/* The reference's memo is a std::map that never refuses an entry (PlanCoverageEstimator.cpp:70,102-110); the edge-bitmap row pool
 * here is finite (70 % of the free device memory, or IPP_CACHE_GB).  When the new edges of a call do not fit, the rows of edges the
 * call does not use are dropped, and a population that still needs more rows than the pool holds is evaluated in pieces; results
 * never depend on what was cached.  This returns how often rows were dropped that way since the evaluator was made. */
int64_t ipp_evictions(const ipp_t* h);
/* Chooses the streaming kernel of the plan reduce (K6) for A/B measurements, process wide: -1 = the library's choice, 0 = plain
 * vector loads, 1.. = configurations of the bulk-asynchronous-copy kernel.  Returns a description of the variant.  Every variant
 * returns the same bits. */
const char* ipp_plan_reduce_variant(int v);
/* drop all memoised edges and collision flags */
int ipp_clear_memo(ipp_t* h);

/* interpretAndExportPlan (cpp_binding.i:38; PlanInterpreterBoxOrder.cpp:301-324).
 * pos: up to (n+1) x 3, dirs: up to (n+1) x 3. */
int ipp_interpret_plan(const ipp_t* h, const double* genes, int n, int width, double* pos, double* dirs, int* n_pos,
                       int* n_dirs);

/* ---- test hooks for per-kernel parity ---- */
/* seen-triangle bitmap of one plan edge prev->curr (prev = -1: the start location),
 * GetColorsDuringTraversal (CameraEstimator.cpp:131-178). */
int ipp_edge_bitmap(ipp_t* h, int prev, int curr, double angle, uint32_t* out_W);
/* edgeComesTooCloseToStructure (OsgHelpers.cpp:246-262) for n edges given as xyz pairs (n x 6 doubles) */
int ipp_edges_collide(ipp_t* h, const double* edges_nx6, int64_t n, uint8_t* out_flags);
/* checkPolytopeForIntersections(generateBoxPolytope(c, half)) (OsgHelpers.cpp:295-332) for n cubes (n x 3) */
int ipp_boxes_hit_mesh(ipp_t* h, const double* centres_nx3, int64_t n, double half_side, uint8_t* out_flags);
/* one camera pose -> per-pixel nearest triangle id (-1 background), row-major [y][x] */
int ipp_render_ids(ipp_t* h, const double* eye3, const double* center3, const double* up3, int32_t* out_HxW);
/* the depth-sorted cluster table the binned visibility kernel builds for one camera pose: up to cap packed entries
 * [depth:16 | tx0:5 tx1:5 ty0:5 ty1:5 | cluster:28], *count = entries of the pose, out_tile_mask32 = occupied tiles; out_first /
 * out_count (nullable, cap entries): the run of BVH-ordered triangles each entry's cluster stands for */
int ipp_leaf_table(ipp_t* h, const double* eye3, const double* center3, const double* up3, uint64_t* out, int64_t cap, int64_t* count,
                   uint32_t* out_tile_mask32, int32_t* out_first, int32_t* out_count);
/* triangle id (file order) at every position of the BVH-ordered triangle arrays: out_T[position] = id */
int ipp_bvh_order(ipp_t* h, int32_t* out_T);
/* BVH self-check: closest-hit of n rays (origin xyz, dir xyz) through the BVH and by brute force; returns
 * the number of mismatching ids */
int ipp_bvh_selftest(ipp_t* h, const float* rays_nx6, int64_t n, int32_t* out_ids, int64_t* mismatches);
/* same, also returning the brute-force ids and (t_bvh, t_brute) per ray */
int ipp_bvh_selftest_ex(ipp_t* h, const float* rays_nx6, int64_t n, int32_t* out_ids, int32_t* out_brute_ids, float* out_t_nx2,
                        int64_t* mismatches);

/* ---- device memory, stream, timing (plumbing for bench.py) ---- */
int ipp_device_alloc(ipp_t* h, int64_t bytes, void** out);
int ipp_device_free(ipp_t* h, void* p);
int ipp_host_alloc(int64_t bytes, void** out); /* pinned */
int ipp_host_free(void* p);
int ipp_memcpy_h2d(ipp_t* h, void* dst, const void* src, int64_t bytes);
int ipp_memcpy_d2h(ipp_t* h, void* dst, const void* src, int64_t bytes);
int ipp_sync(ipp_t* h);
/* CUDA events on the evaluator's stream */
int ipp_timer_start(ipp_t* h);
int ipp_timer_stop(ipp_t* h, double* ms);
/* overwrite a buffer larger than L2 */
int ipp_flush_l2(ipp_t* h);
/* measured read bandwidth (GB/s) of `bytes` bytes (at most 256 MiB) read `reps` times with plain 16-byte loads by a grid over all
 * SMs: a size inside L2 gives the L2 read peak bench.py sets the visibility kernel's lts__t_bytes rate against, a size beyond it
 * the HBM read peak */
int ipp_read_bandwidth(ipp_t* h, int64_t bytes, int reps, double* gbs);
/* counters since the last reset: [0] kernel launches, [1] pixel-centre rays issued, [2] camera poses of the edges rendered (what
 * the reference renders: CameraEstimator.cpp:131-178), [3] edges rendered, [4] collision queries, [5] plan_reduce launches,
 * [6] plans reduced, [7] camera poses actually rendered (poses that several edges of a batch share are rendered once) */
int ipp_counters(ipp_t* h, int64_t* out8, int reset);
/* statistics of the edge-visibility kernel since the last reset (reset by ipp_counters too):
 * [0] acceleration-structure data fetched in 64-byte units (BVH nodes of the tile-walking variant, 8 leaf-bin entries of the
 * binned variant), [1] triangles staged (48 B each), [2] 32x32-pixel tiles rasterised, [3] pixel-centre rays, and, only when the
 * library was built with -DIPP_VIS_DIAG (they cost 2 % of the kernel; 0 otherwise), for the binned variant [4] staged triangles that
 * survived set-up (can produce a fragment in their tile), [5] tiles in which some ray never hit, so that the leaf list was walked
 * to its end; binned variant, always: [6] 128x128 supertiles rasterised, [7] how often a supertile (or a part of one) held more
 * triangles than the record scratch of its warp and was split into halves */
int ipp_visibility_stats(ipp_t* h, int64_t* out8, int reset);
/* which edge-visibility kernel renders the item buffers: 1 = depth-sorted leaf bins per 128x128 supertile (the default when the
 * mesh has <= 16384 BVH leaves and the image is <= 1024 pixels a side), 0 = BVH frustum walk per tile.  Same sampling rule and
 * rasteriser either way.  *out_active (nullable) receives the variant in force after the call; use_bins < 0 only queries. */
int ipp_visibility_variant(ipp_t* h, int use_bins, int* out_active);
/* accumulated CUDA-event time of a named kernel group since the last reset:
 * 0 = plan_reduce, 1 = edge_visibility, 2 = edge_collision, 3 = pack, 4 = energy+mark, 5 = leaf tables of the binned variant */
int ipp_kernel_time(ipp_t* h, int which, double* ms, int64_t* launches, int reset);
int ipp_set_profiling(ipp_t* h, int on);

/* ---- multi-GPU: one process per GPU, population sharded by individual ---- */
/* 128-byte NCCL unique id (rank 0 creates, the host side distributes it) */
int ipp_comm_unique_id(uint8_t* out128);
int ipp_comm_init(ipp_t* h, int rank, int world, const uint8_t* id128);
int ipp_comm_destroy(ipp_t* h);
/* Every rank passes the same whole population.  The edges of the population that no rank has rendered yet are rendered once
 * across the ranks (contiguous chunks of the sorted new keys) and their bitmap rows all-gathered, so that every rank's edge memo
 * is complete; then this rank evaluates its contiguous shard and the fitness rows are all-gathered (ncclAllGather over NVLink):
 * every rank holds all pop x 4.  The ranks must have evaluated the same populations before (same memo contents; checked). */
int ipp_evaluate_population_sharded(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets,
                                    int64_t pop, int memoization, int disable_energy_limit, double* fitness_popx4);
/* Device-resident form of the same step: this rank's `mine` plans (CSR on the device, offsets starting at 0) are
 * evaluated into slot `rank` of d_fitness_all (world * per_rank rows of 4 doubles, device memory) and the slots are
 * all-gathered in place.  Asynchronous on the evaluator's stream. */
int ipp_evaluate_population_sharded_device(ipp_t* h, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets,
                                           int64_t mine, int64_t n_genes, int64_t per_rank, int memoization,
                                           int disable_energy_limit, double* d_fitness_all);
int ipp_comm_barrier(ipp_t* h);
int ipp_comm_allreduce_max(ipp_t* h, double* inout, int n);
int ipp_comm_allreduce_sum(ipp_t* h, double* inout, int n);

#ifdef __cplusplus
}
#endif
#endif /* IPP_H_ */
