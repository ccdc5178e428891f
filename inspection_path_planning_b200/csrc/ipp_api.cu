// ipp_api.cu -- the extern "C" surface of libipp.so declared in include/ipp.h.
#include <cstring>
#include <string>

#include "../../include/ipp.h"
#include "comm.h"
#include "evaluator.h"
#include "mesh_io.h"

using ipp::Evaluator;

struct ipp_evaluator {
    Evaluator* ev;
};

namespace {
thread_local std::string g_err;
inline Evaluator& E(ipp_t* h) { return *h->ev; }
inline const Evaluator& E(const ipp_t* h) { return *h->ev; }
}  // namespace

#define IPP_TRY try {
#define IPP_CATCH                     \
    }                                 \
    catch (const std::exception& e) { \
        g_err = e.what();             \
        cudaGetLastError();           \
        return 1;                     \
    }                                 \
    catch (...) {                     \
        g_err = "unknown error";      \
        return 1;                     \
    }                                 \
    return 0;

extern "C" {

const char* ipp_last_error(void) { return g_err.c_str(); }

int ipp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ipp_create(const char* scene_path, const double* sensor_specs, int post_processing, const double* start_xyz, int plan_loops_around,
               int printer_friendly, int device, ipp_t** out) {
    (void)printer_friendly;  // only affects the reference's plotting colours
    IPP_TRY
    if (!scene_path || !sensor_specs || !out) throw std::invalid_argument("libipp: null argument");
    Evaluator* ev = new Evaluator(scene_path, sensor_specs, post_processing != 0, start_xyz, plan_loops_around != 0, device);
    ipp_t* h = new ipp_t;
    h->ev = ev;
    *out = h;
    IPP_CATCH
}

void ipp_destroy(ipp_t* h) {
    if (!h) return;
    if (h->ev->comm_) ipp::comm_destroy(h->ev->comm_);
    delete h->ev;
    delete h;
}

int ipp_num_boxes(const ipp_t* h) { return E(h).numBoxes(); }
int64_t ipp_num_triangles(const ipp_t* h) { return E(h).numTriangles(); }
int64_t ipp_bitmap_words(const ipp_t* h) { return E(h).bitmapWords(); }
double ipp_max_allowed_energy(const ipp_t* h) { return E(h).maxAllowedEnergy(); }
double ipp_total_area(const ipp_t* h) { return E(h).totalArea_; }
double ipp_collision_penalty(const ipp_t* h) { return E(h).collisionPenalty_; }
void ipp_scene_info(const ipp_t* h, double* center3, double* bbox6) {
    for (int k = 0; k < 3; ++k) {
        center3[k] = E(h).sceneCenter_[k];
        bbox6[k] = E(h).bbmin_[k];
        bbox6[3 + k] = E(h).bbmax_[k];
    }
}
void ipp_image_size(const ipp_t* h, int* width, int* height) {
    *width = E(h).cam_.W;
    *height = E(h).cam_.H;
}
int ipp_load_mesh(const char* scene_path, float* out_Tx9, int64_t cap_triangles, int64_t* n_triangles) {
    IPP_TRY
    std::vector<float> tris;
    ipp::load_mesh(scene_path, tris);
    *n_triangles = (int64_t)(tris.size() / 9);
    if (out_Tx9) memcpy(out_Tx9, tris.data(), (size_t)std::min<int64_t>(*n_triangles, cap_triangles) * 9 * sizeof(float));
    IPP_CATCH
}
int ipp_areas(const ipp_t* h, double* out) {
    memcpy(out, E(h).areaHost_.data(), E(h).areaHost_.size() * sizeof(double));
    return 0;
}
int ipp_box_centres(const ipp_t* h, double* out) {
    memcpy(out, E(h).grid_.centres.data(), E(h).grid_.centres.size() * sizeof(double));
    return 0;
}
void ipp_grid_dims(const ipp_t* h, int* nz, int* nx, int* ny) {
    *nz = E(h).grid_.nz;
    *nx = E(h).grid_.nx;
    *ny = E(h).grid_.ny;
}
int ipp_grid(const ipp_t* h, int32_t* out) {
    memcpy(out, E(h).grid_.cell.data(), E(h).grid_.cell.size() * sizeof(int32_t));
    return 0;
}

int ipp_simple_plans(ipp_t* h, int* n_plans, int* n_genes, double* ids, int32_t* offsets) {
    IPP_TRY
    const auto& sp = E(h).simplePlans();
    *n_plans = (int)sp.size();
    int tot = 0;
    for (const auto& p : sp) tot += (int)p.size();
    *n_genes = tot;
    if (ids) {
        int k = 0;
        offsets[0] = 0;
        for (size_t i = 0; i < sp.size(); ++i) {
            for (const auto& g : sp[i]) ids[k++] = g[0];
            offsets[i + 1] = k;
        }
    }
    IPP_CATCH
}

int ipp_evaluate_plan(ipp_t* h, const double* genes, int n, int width, int memoization, int how_to_plot, int disable_energy_limit,
                      double* out4) {
    IPP_TRY
    if (how_to_plot == IPP_PLOT_IMAGE)  // PlanCoverageEstimator.cpp:206-208
        throw std::logic_error("ERROR! Asked to plot image, but not given an OSG-drawable to plot to.");
    if (how_to_plot != IPP_PLOT_NOTHING) throw std::logic_error("libipp: only how_to_plot == nothing is supported (no GL viewers)");
    if (width != 1 && width != 2) throw std::invalid_argument("libipp: a gene is [id] or [id, angleOffset]");
    std::vector<int32_t> ids((size_t)n);
    std::vector<float> angles((size_t)n, 0.f);
    for (int i = 0; i < n; ++i) {
        size_t id = (size_t)(genes[(size_t)i * width] + 0.5);  // PlanInterpreterBoxOrder.cpp:280 (a -1 gene decodes to 0)
        if (id >= (size_t)E(h).numBoxes()) throw std::out_of_range("libipp: box id out of range");
        ids[i] = (int32_t)id;
        if (width == 2) angles[i] = (float)genes[(size_t)i * width + 1];  // float sensor_dir_offset, :283-286
    }
    int64_t off[2] = {0, n};
    if (n == 0) {
        out4[0] = 1.0; out4[1] = 0; out4[2] = 0; out4[3] = 0;  // evaluateEmptyPlan
    } else {
        E(h).evaluatePopulationHost(ids.data(), width == 2 ? angles.data() : nullptr, off, 1, memoization != 0, disable_energy_limit != 0, out4,
                                    nullptr);
    }
    IPP_CATCH
}

int ipp_evaluate_population(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, int memoization,
                            int disable_energy_limit, double* fitness, uint32_t* union_bitmap) {
    IPP_TRY
    E(h).evaluatePopulationHost(ids, angles, offsets, pop, memoization != 0, disable_energy_limit != 0, fitness, union_bitmap);
    IPP_CATCH
}

int ipp_evaluate_population_device(ipp_t* h, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t pop,
                                   int64_t n_genes, int memoization, int disable_energy_limit, double* d_fitness) {
    IPP_TRY
    E(h).evaluatePopulationDevice(d_ids, d_angles, d_offsets, pop, n_genes, memoization != 0, disable_energy_limit != 0, d_fitness, nullptr);
    IPP_CATCH
}

int ipp_update_memoised_edges(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop, int64_t* remaining) {
    IPP_TRY
    *remaining = E(h).updateMemoisedEdges(ids, angles, offsets, pop);
    IPP_CATCH
}
int64_t ipp_memo_size(const ipp_t* h) { return E(h).memoSize(); }
int64_t ipp_evictions(const ipp_t* h) { return E(h).evictions(); }
int ipp_read_bandwidth(ipp_t* h, int64_t bytes, int reps, double* gbs) {
    IPP_TRY
    *gbs = E(h).readBandwidth(bytes, reps);
    IPP_CATCH
}
const char* ipp_plan_reduce_variant(int v) {
    ipp::set_plan_reduce_variant(v);
    return ipp::plan_reduce_variant_name(v);
}
int ipp_clear_memo(ipp_t* h) {
    IPP_TRY
    E(h).clearMemo();
    IPP_CATCH
}

int ipp_interpret_plan(const ipp_t* h, const double* genes, int n, int width, double* pos, double* dirs, int* n_pos, int* n_dirs) {
    IPP_TRY
    // (a post-processing evaluator may have to render the edges of the plan to decide their headings: not const underneath)
    const_cast<Evaluator&>(E(h)).interpretPlan(genes, n, width, pos, dirs, n_pos, n_dirs);
    IPP_CATCH
}

int ipp_edge_bitmap(ipp_t* h, int prev, int curr, double angle, uint32_t* out) {
    IPP_TRY
    E(h).edgeBitmap(prev, curr, angle, out);
    IPP_CATCH
}
int ipp_edges_collide(ipp_t* h, const double* edges, int64_t n, uint8_t* out) {
    IPP_TRY
    E(h).edgesCollide(edges, n, out);
    IPP_CATCH
}
int ipp_boxes_hit_mesh(ipp_t* h, const double* centres, int64_t n, double half, uint8_t* out) {
    IPP_TRY
    E(h).boxesHitMesh(centres, n, half, out);
    IPP_CATCH
}
int ipp_render_ids(ipp_t* h, const double* eye, const double* center, const double* up, int32_t* out) {
    IPP_TRY
    E(h).renderIds(eye, center, up, out);
    IPP_CATCH
}
int ipp_bvh_selftest(ipp_t* h, const float* rays, int64_t n, int32_t* out_ids, int64_t* mismatches) {
    IPP_TRY
    E(h).bvhSelftest(rays, n, out_ids, mismatches);
    IPP_CATCH
}

int ipp_bvh_selftest_ex(ipp_t* h, const float* rays, int64_t n, int32_t* out_ids, int32_t* out_brute, float* out_t2, int64_t* mismatches) {
    IPP_TRY
    E(h).bvhSelftest(rays, n, out_ids, mismatches, out_brute, out_t2);
    IPP_CATCH
}

int ipp_device_alloc(ipp_t* h, int64_t bytes, void** out) {
    IPP_TRY
    E(h).bind();
    IPP_CUDA_CHECK(cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1)));
    IPP_CATCH
}
int ipp_device_free(ipp_t* h, void* p) {
    IPP_TRY
    E(h).bind();
    IPP_CUDA_CHECK(cudaFree(p));
    IPP_CATCH
}
int ipp_host_alloc(int64_t bytes, void** out) {
    IPP_TRY
    IPP_CUDA_CHECK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    IPP_CATCH
}
int ipp_host_free(void* p) {
    IPP_TRY
    IPP_CUDA_CHECK(cudaFreeHost(p));
    IPP_CATCH
}
int ipp_memcpy_h2d(ipp_t* h, void* dst, const void* src, int64_t bytes) {
    IPP_TRY
    E(h).bind();
    IPP_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, E(h).stream()));
    IPP_CATCH
}
int ipp_memcpy_d2h(ipp_t* h, void* dst, const void* src, int64_t bytes) {
    IPP_TRY
    E(h).bind();
    IPP_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, E(h).stream()));
    E(h).sync();
    IPP_CATCH
}
int ipp_sync(ipp_t* h) {
    IPP_TRY
    E(h).bind();
    E(h).sync();
    IPP_CATCH
}
int ipp_timer_start(ipp_t* h) {
    IPP_TRY
    E(h).timerStart();
    IPP_CATCH
}
int ipp_timer_stop(ipp_t* h, double* ms) {
    IPP_TRY
    *ms = E(h).timerStop();
    IPP_CATCH
}
int ipp_flush_l2(ipp_t* h) {
    IPP_TRY
    E(h).flushL2();
    IPP_CATCH
}
int ipp_counters(ipp_t* h, int64_t* out8, int reset) {
    IPP_TRY
    E(h).counters(out8, reset != 0);
    IPP_CATCH
}
int ipp_visibility_stats(ipp_t* h, int64_t* out8, int reset) {
    IPP_TRY
    E(h).visibilityStats(out8, reset != 0);
    IPP_CATCH
}
int ipp_leaf_table(ipp_t* h, const double* eye3, const double* center3, const double* up3, uint64_t* out, int64_t cap, int64_t* count,
                   uint32_t* out_tile_mask32, int32_t* out_first, int32_t* out_count) {
    IPP_TRY
    E(h).leafTable(eye3, center3, up3, out, cap, count, out_tile_mask32, out_first, out_count);
    IPP_CATCH
}
int ipp_bvh_order(ipp_t* h, int32_t* out_T) {
    IPP_TRY
    E(h).bvhOrder(out_T);
    IPP_CATCH
}
int ipp_visibility_variant(ipp_t* h, int use_bins, int* out_active) {
    IPP_TRY
    if (use_bins >= 0) E(h).setUseBins(use_bins != 0);
    if (out_active) *out_active = E(h).usesBins() ? 1 : 0;
    IPP_CATCH
}
int ipp_kernel_time(ipp_t* h, int which, double* ms, int64_t* launches, int reset) {
    IPP_TRY
    E(h).kernelTime(which, ms, launches, reset != 0);
    IPP_CATCH
}
int ipp_set_profiling(ipp_t* h, int on) {
    E(h).setProfiling(on != 0);
    return 0;
}

// ---- multi-GPU ----
int ipp_comm_unique_id(uint8_t* out128) {
    IPP_TRY
    ipp::comm_unique_id(out128);
    IPP_CATCH
}
int ipp_comm_init(ipp_t* h, int rank, int world, const uint8_t* id128) {
    IPP_TRY
    E(h).bind();
    if (E(h).comm_) throw std::logic_error("libipp: communicator already initialised");
    E(h).comm_ = ipp::comm_create(rank, world, id128, E(h).stream());
    IPP_CATCH
}
int ipp_comm_destroy(ipp_t* h) {
    IPP_TRY
    if (E(h).comm_) {
        ipp::comm_destroy(E(h).comm_);
        E(h).comm_ = nullptr;
    }
    IPP_CATCH
}
int ipp_evaluate_population_sharded(ipp_t* h, const int32_t* ids, const float* angles, const int64_t* offsets, int64_t pop,
                                    int memoization, int disable_energy_limit, double* fitness) {
    IPP_TRY
    Evaluator& ev = E(h);
    ev.bind();
    if (!ev.comm_) throw std::logic_error("libipp: ipp_comm_init has not been called");
    int rank = ipp::comm_rank(ev.comm_), world = ipp::comm_world(ev.comm_);
    // contiguous shard of ceil(pop / world) individuals per rank
    int64_t per = (pop + world - 1) / world;
    int64_t lo = std::min<int64_t>(pop, per * rank), hi = std::min<int64_t>(pop, lo + per);
    ipp::comm_eval_and_allgather(ev, ids, angles, offsets, pop, per, lo, hi, memoization != 0, disable_energy_limit != 0, fitness);
    IPP_CATCH
}
int ipp_evaluate_population_sharded_device(ipp_t* h, const int32_t* d_ids, const float* d_angles, const int64_t* d_offsets, int64_t mine,
                                           int64_t n_genes, int64_t per_rank, int memoization, int disable_energy_limit,
                                           double* d_fitness_all) {
    IPP_TRY
    Evaluator& ev = E(h);
    ev.bind();
    if (!ev.comm_) throw std::logic_error("libipp: ipp_comm_init has not been called");
    ipp::comm_eval_device_and_allgather(ev, d_ids, d_angles, d_offsets, mine, n_genes, per_rank, memoization != 0, disable_energy_limit != 0,
                                        d_fitness_all);
    IPP_CATCH
}
int ipp_comm_barrier(ipp_t* h) {
    IPP_TRY
    E(h).bind();
    if (!E(h).comm_) throw std::logic_error("libipp: ipp_comm_init has not been called");
    double z = 0;
    ipp::comm_allreduce(E(h).comm_, &z, 1, 0, E(h).stream());
    IPP_CATCH
}
int ipp_comm_allreduce_max(ipp_t* h, double* inout, int n) {
    IPP_TRY
    E(h).bind();
    if (!E(h).comm_) throw std::logic_error("libipp: ipp_comm_init has not been called");
    ipp::comm_allreduce(E(h).comm_, inout, n, 1, E(h).stream());
    IPP_CATCH
}
int ipp_comm_allreduce_sum(ipp_t* h, double* inout, int n) {
    IPP_TRY
    E(h).bind();
    if (!E(h).comm_) throw std::logic_error("libipp: ipp_comm_init has not been called");
    ipp::comm_allreduce(E(h).comm_, inout, n, 0, E(h).stream());
    IPP_CATCH
}

}  // extern "C"
