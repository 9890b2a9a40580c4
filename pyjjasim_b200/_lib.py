"""
ctypes binding of libjjstep.so (the C ABI declared in include/jjstep.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``pyjjasim_b200/csrc/build.sh``.
There is no CPU fallback: if the library is missing or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# JJ_LIB_PATH: an alternative build of the same library (build variants measured side by side, e.g. -DJJ_NOISE_F64)
LIB_PATH = os.environ.get("JJ_LIB_PATH") or os.path.join(_HERE, "libjjstep.so")

JJ_SRC_IS, JJ_SRC_F, JJ_SRC_VS, JJ_SRC_T = 0, 1, 2, 3
JJ_KIND_ZERO, JJ_KIND_RANK1, JJ_KIND_DENSE = 0, 1, 2
JJ_ENGINE_AUTO, JJ_ENGINE_STREAMING, JJ_ENGINE_SUBDOMAIN = 0, 1, 3
JJ_ENONFINITE = -5

EXPORTS = ["jj_create", "jj_destroy", "jj_last_error", "jj_set_circuit", "jj_set_solver", "jj_set_problem", "jj_set_engine",
           "jj_set_state", "jj_get_state", "jj_set_source", "jj_upload_source", "jj_upload_noise",
           "jj_alloc_outputs", "jj_run", "jj_fetch_theta", "jj_fetch_current", "jj_debug_noise",
           "jj_debug_solve", "jj_stats", "jj_set_subdomain_plan", "jj_debug_subdomain_solve", "jj_sm_count",
           "jj_restart_at_rest", "jj_adopt_state_at_rest", "jj_vortex_configuration", "jj_vortex_mobility", "jj_vortex_configurations", "jj_observe_begin", "jj_observe_fetch", "jj_anneal",
           "jj_host_alloc", "jj_host_free"]

_p = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)
_f64p = C.POINTER(C.c_double)


class JJSweep(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("level_ptr", _i32p), ("group_ptr", _i32p), ("n_tiles", C.c_int32),
                ("tile_row0", _i32p), ("tile_nrows", _i32p), ("tile_lpr", _i32p), ("tile_nsteps", _i32p),
                ("tile_flags", _i32p), ("tile_col_off", _i64p), ("tile_val_off", _i64p),
                ("n_cols", C.c_int64), ("cols", _i32p), ("n_vals", C.c_int64), ("vals", _f64p),
                ("stage_rows", C.c_int32), ("tile_stage_off", _i32p)]


class JJSubProgram(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_bwd", C.c_int32), ("n_warps", C.c_int32), ("n_tiles", C.c_int32),
                ("wt_ptr", _i32p), ("ws_ptr", _i32p), ("thdr", _i32p), ("lstaged", _i32p), ("n_steps", C.c_int64),
                ("stream", C.POINTER(C.c_uint8))]


class JJSubdomainPlan(C.Structure):
    _fields_ = [("P", C.c_int32), ("NG", C.c_int32), ("n_rows", C.c_int32), ("n_loc_max", C.c_int32),
                ("stage_rows", C.c_int32), ("n_top", C.c_int32), ("n_up_pad", C.c_int32), ("n_slots", C.c_int32),
                ("tt0", C.c_int32), ("n_tt", C.c_int32), ("n_tt_pad", C.c_int32),
                ("n_loc", _i32p), ("n_halo", _i32p), ("hptr", _i32p), ("halo_top", _i32p), ("tptr", _i32p),
                ("tslot", _i32p), ("top_face", _i32p), ("Sinv_packed", _f64p),
                ("up_RB", C.c_int32), ("up_KB", C.c_int32), ("n_up_fwd", C.c_int32), ("n_up_bwd", C.c_int32),
                ("n_up_tasks", C.c_int32), ("up_phase_ptr", _i32p), ("up_phase_split", _i32p), ("up_task", _i32p), ("up_task_aoff", _i64p),
                ("n_up_cols", C.c_int64), ("up_cols", _i32p), ("n_up_vals", C.c_int64), ("up_A", _f64p),
                ("prog", C.POINTER(JJSubProgram)),
                ("junc_ptr", _i32p), ("junc_orig", _i32p), ("junc_row", _i32p), ("junc_sign", _i8p),
                ("face_K", C.c_int32), ("face_ell_j", _i32p), ("face_ell_c", _f64p), ("face_fidx", _i32p)]


class JJCircuit(C.Structure):
    _fields_ = [("Nj", C.c_int32), ("Nf", C.c_int32), ("face_ptr", _i32p), ("face_junc", _i32p),
                ("face_sign", _i8p), ("junc_face", _i32p), ("junc_sign", _i8p),
                ("Ic", _f64p), ("c0", _f64p), ("c1", _f64p), ("c2", _f64p),
                ("cpr_harmonics", C.c_int32), ("cpr_a", _f64p), ("cpr_b", _f64p)]


class JJStats(C.Structure):
    _fields_ = [("engine", C.c_int32), ("cluster_size", C.c_int32), ("tile_problems", C.c_int32),
                ("steps_done", C.c_int64), ("kernel_launches", C.c_int64), ("step_ms", C.c_double),
                ("device_bytes", C.c_int64), ("non_finite", C.c_int32)]


_lib = None


def load():
    """Load libjjstep.so and declare prototypes. Raises RuntimeError if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the time-evolution path)")
    lib = C.CDLL(LIB_PATH)
    lib.jj_create.argtypes = [C.c_int, C.POINTER(_p)]
    lib.jj_destroy.argtypes = [_p]
    lib.jj_destroy.restype = None
    lib.jj_last_error.argtypes = [_p]
    lib.jj_last_error.restype = C.c_char_p
    lib.jj_set_circuit.argtypes = [_p, C.POINTER(JJCircuit)]
    lib.jj_set_solver.argtypes = [_p, C.POINTER(JJSweep), C.POINTER(JJSweep)]
    lib.jj_set_problem.argtypes = [_p, C.c_int32, C.c_double, C.c_uint64, C.c_int64, C.c_int32]
    lib.jj_set_engine.argtypes = [_p, C.c_int32]
    lib.jj_set_state.argtypes = [_p, _f64p, _f64p]
    lib.jj_get_state.argtypes = [_p, _f64p, _f64p]
    lib.jj_set_source.argtypes = [_p, C.c_int32, C.c_int32, C.c_int32, _f64p]
    lib.jj_upload_source.argtypes = [_p, C.c_int32, C.c_int64, C.c_int32, _f64p]
    lib.jj_upload_noise.argtypes = [_p, C.c_int64, C.c_int32, _f64p]
    lib.jj_alloc_outputs.argtypes = [_p, C.c_int64, C.c_int64]
    lib.jj_run.argtypes = [_p, C.c_int64, C.c_int32, _i64p, _i64p]
    lib.jj_fetch_theta.argtypes = [_p, C.c_int64, C.c_int64, _f64p]
    lib.jj_fetch_current.argtypes = [_p, C.c_int64, C.c_int64, _f64p]
    lib.jj_debug_noise.argtypes = [_p, C.c_int64, _f64p]
    lib.jj_debug_solve.argtypes = [_p, _f64p, _f64p]
    lib.jj_stats.argtypes = [_p, C.POINTER(JJStats)]
    lib.jj_set_subdomain_plan.argtypes = [_p, C.POINTER(JJSubdomainPlan)]
    lib.jj_debug_subdomain_solve.argtypes = [_p, _f64p, _f64p]
    lib.jj_sm_count.argtypes = [C.c_int]
    lib.jj_restart_at_rest.argtypes = [_p]
    lib.jj_adopt_state_at_rest.argtypes = [_p, _p]
    lib.jj_vortex_configuration.argtypes = [_p, C.c_int64, _i32p]
    lib.jj_vortex_mobility.argtypes = [_p, C.c_int64, C.c_int64, _i64p]
    lib.jj_vortex_configurations.argtypes = [_p, C.c_int64, C.c_int64, _i32p, _i32p]
    lib.jj_anneal.argtypes = [_p, C.c_int64, C.c_int32, C.c_int32, _f64p, C.c_double, C.c_double, C.c_double, _f64p, _f64p,
                              C.POINTER(C.c_double)]
    lib.jj_observe_begin.argtypes = [_p, C.c_int64, C.c_int32]
    lib.jj_observe_fetch.argtypes = [_p, C.POINTER(C.c_int64), _i32p, _f64p, _f64p, _i32p]
    lib.jj_host_alloc.argtypes = [C.c_int, C.c_uint64, C.POINTER(_p)]
    lib.jj_host_free.argtypes = [_p]
    for name in EXPORTS:
        if name not in ("jj_destroy", "jj_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def f64(a):
    return a.ctypes.data_as(_f64p)


def i32(a):
    return a.ctypes.data_as(_i32p)


def i64(a):
    return a.ctypes.data_as(_i64p)


def i8(a):
    return a.ctypes.data_as(_i8p)


def c_f64(a):
    return np.ascontiguousarray(a, dtype=np.double)
