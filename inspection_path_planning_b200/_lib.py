"""ctypes binding of libipp.so (include/ipp.h).  No torch, no CPU fallback: if the shared library is
missing this module raises, and every compute call needs a CUDA device."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libipp.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
H = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/ipp.h declares
PROTOTYPES = {
    "ipp_last_error": (C.c_char_p, []),
    "ipp_device_count": (C.c_int, []),
    "ipp_create": (C.c_int, [C.c_char_p, c_f64p, C.c_int, c_f64p, C.c_int, C.c_int, C.c_int, C.POINTER(H)]),
    "ipp_destroy": (None, [H]),
    "ipp_num_boxes": (C.c_int, [H]),
    "ipp_num_triangles": (C.c_int64, [H]),
    "ipp_bitmap_words": (C.c_int64, [H]),
    "ipp_max_allowed_energy": (C.c_double, [H]),
    "ipp_total_area": (C.c_double, [H]),
    "ipp_collision_penalty": (C.c_double, [H]),
    "ipp_scene_info": (None, [H, c_f64p, c_f64p]),
    "ipp_image_size": (None, [H, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ipp_areas": (C.c_int, [H, c_f64p]),
    "ipp_box_centres": (C.c_int, [H, c_f64p]),
    "ipp_grid_dims": (None, [H, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ipp_grid": (C.c_int, [H, c_i32p]),
    "ipp_simple_plans": (C.c_int, [H, C.POINTER(C.c_int), C.POINTER(C.c_int), c_f64p, c_i32p]),
    "ipp_evaluate_plan": (C.c_int, [H, c_f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p]),
    "ipp_evaluate_population": (C.c_int, [H, c_i32p, c_f32p, c_i64p, C.c_int64, C.c_int, C.c_int, c_f64p, c_u32p]),
    "ipp_evaluate_population_device": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "ipp_update_memoised_edges": (C.c_int, [H, c_i32p, c_f32p, c_i64p, C.c_int64, c_i64p]),
    "ipp_load_mesh": (C.c_int, [C.c_char_p, c_f32p, C.c_int64, c_i64p]),
    "ipp_memo_size": (C.c_int64, [H]),
    "ipp_evictions": (C.c_int64, [H]),
    "ipp_plan_reduce_variant": (C.c_char_p, [C.c_int]),
    "ipp_clear_memo": (C.c_int, [H]),
    "ipp_interpret_plan": (C.c_int, [H, c_f64p, C.c_int, C.c_int, c_f64p, c_f64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ipp_edge_bitmap": (C.c_int, [H, C.c_int, C.c_int, C.c_double, c_u32p]),
    "ipp_edges_collide": (C.c_int, [H, c_f64p, C.c_int64, c_u8p]),
    "ipp_boxes_hit_mesh": (C.c_int, [H, c_f64p, C.c_int64, C.c_double, c_u8p]),
    "ipp_render_ids": (C.c_int, [H, c_f64p, c_f64p, c_f64p, c_i32p]),
    "ipp_leaf_table": (C.c_int, [H, c_f64p, c_f64p, c_f64p, C.POINTER(C.c_uint64), C.c_int64, c_i64p, c_u32p, c_i32p, c_i32p]),
    "ipp_bvh_order": (C.c_int, [H, c_i32p]),
    "ipp_bvh_selftest": (C.c_int, [H, c_f32p, C.c_int64, c_i32p, c_i64p]),
    "ipp_bvh_selftest_ex": (C.c_int, [H, c_f32p, C.c_int64, c_i32p, c_i32p, c_f32p, c_i64p]),
    "ipp_device_alloc": (C.c_int, [H, C.c_int64, C.POINTER(C.c_void_p)]),
    "ipp_device_free": (C.c_int, [H, C.c_void_p]),
    "ipp_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "ipp_host_free": (C.c_int, [C.c_void_p]),
    "ipp_memcpy_h2d": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int64]),
    "ipp_memcpy_d2h": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int64]),
    "ipp_sync": (C.c_int, [H]),
    "ipp_timer_start": (C.c_int, [H]),
    "ipp_timer_stop": (C.c_int, [H, c_f64p]),
    "ipp_flush_l2": (C.c_int, [H]),
    "ipp_read_bandwidth": (C.c_int, [H, C.c_int64, C.c_int, c_f64p]),
    "ipp_counters": (C.c_int, [H, c_i64p, C.c_int]),
    "ipp_visibility_stats": (C.c_int, [H, c_i64p, C.c_int]),
    "ipp_visibility_variant": (C.c_int, [H, C.c_int, C.POINTER(C.c_int)]),
    "ipp_kernel_time": (C.c_int, [H, C.c_int, c_f64p, c_i64p, C.c_int]),
    "ipp_set_profiling": (C.c_int, [H, C.c_int]),
    "ipp_comm_unique_id": (C.c_int, [c_u8p]),
    "ipp_comm_init": (C.c_int, [H, C.c_int, C.c_int, c_u8p]),
    "ipp_comm_destroy": (C.c_int, [H]),
    "ipp_evaluate_population_sharded": (C.c_int, [H, c_i32p, c_f32p, c_i64p, C.c_int64, C.c_int, C.c_int, c_f64p]),
    "ipp_evaluate_population_sharded_device": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                                          C.c_int, C.c_void_p]),
    "ipp_comm_barrier": (C.c_int, [H]),
    "ipp_comm_allreduce_max": (C.c_int, [H, c_f64p, C.c_int]),
    "ipp_comm_allreduce_sum": (C.c_int, [H, c_f64p, C.c_int]),
}

_lib = None


def load():
    """Loads libipp.so (building nothing: run `python -m inspection_path_planning_b200.build` first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libipp.so is missing (%s); build it with `python -m inspection_path_planning_b200.build`. "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        f = getattr(L, name)  # AttributeError here means the library does not export what ipp.h declares
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc:
        msg = load().ipp_last_error().decode(errors="replace")
        if "out of range" in msg:
            raise IndexError(msg)
        if "valid 3D model" in msg or "invalid" in msg.lower():
            raise ValueError(msg)
        raise RuntimeError(msg)
