"""Loader of the C-ABI shared library (include/greyjack_b200.h).  No fallback of any
kind: if the library is missing or has no CUDA device, calls raise."""
from __future__ import annotations

import ctypes as C
import os

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("GREYJACK_B200_LIB") or os.path.join(_PKG_ROOT, "libgreyjack_b200.so")

GJ_OK = 0


class GjError(RuntimeError):
    pass


class ProblemDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_vars", C.c_int32),
        ("lower_bounds", C.c_void_p), ("upper_bounds", C.c_void_p),
        ("frozen", C.c_void_p), ("initial", C.c_void_p),
        ("n_groups", C.c_int32), ("group_offsets", C.c_void_p), ("group_var_ids", C.c_void_p),
        ("column_id", C.c_void_p),
        ("n_locations", C.c_int32), ("distance_matrix", C.c_void_p), ("coords", C.c_void_p),
        ("n_vehicles", C.c_int32), ("vehicle_depot", C.c_void_p),
        ("vehicle_capacity", C.c_void_p), ("work_day_start", C.c_void_p),
        ("work_day_end", C.c_void_p), ("demand", C.c_void_p), ("tw_start", C.c_void_p),
        ("tw_end", C.c_void_p), ("service_time", C.c_void_p),
        ("time_windowed", C.c_int32),
        ("weights", C.c_double * 4),
        ("score_precision", C.c_int64 * 3),
    ]


class AgentParams(C.Structure):
    _fields_ = [
        ("agent", C.c_int32), ("n_islands", C.c_int32), ("seed", C.c_uint64),
        ("neighbours_count", C.c_int64), ("late_acceptance_size", C.c_int64),
        ("population_size", C.c_int64),
        ("crossover_probability", C.c_double), ("p_best_rate", C.c_double),
        ("migration_rate", C.c_double), ("tabu_entity_rate", C.c_double),
        ("compare_to_global", C.c_int32), ("has_mutation_rate_multiplier", C.c_int32),
        ("mutation_rate_multiplier", C.c_double),
        ("has_move_probas", C.c_int32), ("move_probas", C.c_double * 6),
        ("migration_frequency", C.c_int64),
        ("reference_noop_moves", C.c_int32), ("scoring_mode", C.c_int32),
        ("chain_steps_per_launch", C.c_int32), ("has_cooling_rate", C.c_int32),
        ("cooling_rate", C.c_double), ("initial_temperature", C.c_double * 3),
    ]


class GaTrace(C.Structure):
    _fields_ = [("order_before", C.c_void_p), ("pairs", C.c_void_p), ("move_desc", C.c_void_p),
                ("cand_rows", C.c_void_p), ("cand_scores", C.c_void_p), ("replace", C.c_void_p),
                ("src", C.c_void_p)]


class Term(C.Structure):
    _fields_ = [("op", C.c_int32), ("value_offset", C.c_int32), ("value_stride", C.c_int32),
                ("seg_offset", C.c_int32), ("seg_stride", C.c_int32),
                ("key_value_coef", C.c_int32), ("key_index_coef", C.c_int32), ("variant", C.c_int32),
                ("scale", C.c_double)]


class Constraint(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("level", C.c_int32), ("n_terms", C.c_int32), ("terms", Term * 4)]


# every symbol include/greyjack_b200.h declares
EXPORTED = [
    "gj_last_error", "gj_abi_version", "gj_device_count", "gj_sizeof_problem_desc", "gj_sizeof_agent_params", "gj_launch_count",
    "gj_problem_create", "gj_problem_destroy", "gj_problem_levels", "gj_problem_n_vars",
    "gj_problem_set_constraint_weights", "gj_problem_set_exact_sums", "gj_problem_get_distance_matrix",
    "gj_host_alloc", "gj_host_free",
    "gj_score_plain", "gj_score_incremental", "gj_score_incremental_packed",
    "gj_score_plain_device", "gj_score_plain_i32_device", "gj_score_incremental_device",
    "gj_islands_create", "gj_islands_destroy", "gj_islands_step", "gj_islands_set_accomplish_rate", "gj_islands_stats", "gj_islands_trace_aux", "gj_islands_step_path", "gj_islands_set_profiling", "gj_islands_profile_read",
    "gj_islands_best", "gj_islands_current", "gj_islands_migrant_bytes",
    "gj_islands_set_external_ring", "gj_islands_export_migrants", "gj_islands_import_migrants",
    "gj_islands_global_top_bytes", "gj_islands_export_global_top", "gj_islands_import_global_top",
    "gj_program_create", "gj_program_destroy", "gj_program_add_constraint", "gj_program_remove_constraint",
    "gj_program_set_constraint_weights", "gj_program_n_constraints", "gj_program_get_score",
    "gj_ring_create", "gj_ring_destroy", "gj_ring_handle", "gj_ring_connect", "gj_ring_exchange", "gj_ring_stats", "gj_islands_trace_step", "gj_islands_trace_tabu", "gj_islands_ga_trace_generation", "gj_islands_ga_population",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GjError(f"{LIB_PATH} is missing: build it with `make -C {_PKG_ROOT}` "
                      "(python __graft_entry__.py build); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.gj_last_error.restype = C.c_char_p
    L.gj_islands_step_path.restype = C.c_char_p
    L.gj_islands_step_path.argtypes = [C.c_void_p]
    L.gj_abi_version.restype = C.c_int32
    L.gj_device_count.restype = C.c_int32
    L.gj_launch_count.restype = C.c_int64
    L.gj_problem_levels.restype = C.c_int32
    L.gj_problem_n_vars.restype = C.c_int32
    L.gj_problem_destroy.restype = None
    L.gj_host_free.restype = None
    for name in ("gj_islands_destroy",):
        if hasattr(L, name):
            getattr(L, name).restype = None
    if hasattr(L, "gj_islands_migrant_bytes"):
        L.gj_islands_migrant_bytes.restype = C.c_int64
    if hasattr(L, "gj_islands_global_top_bytes"):
        L.gj_islands_global_top_bytes.restype = C.c_int64
        L.gj_ring_destroy.restype = None
    if hasattr(L, "gj_program_destroy"):
        L.gj_program_destroy.restype = None
        L.gj_program_n_constraints.restype = C.c_int32
    _lib = L
    return L


def check(rc):
    if rc != GJ_OK:
        raise GjError(f"gj status {rc}: {load().gj_last_error().decode()}")
