"""ctypes binding of libhydrograd_b200.so (include/hydrograd_b200.h).  Fails loudly if the CUDA
library has not been built -- there is no Python/CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhydrograd_b200.so")

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int64, C.c_void_p)
c_u8p = C.POINTER(C.c_uint8)

PARAM_NONE, PARAM_ZB, PARAM_MANNING, PARAM_Q, PARAM_UDE = 0, 1, 2, 3, 4
# active_param_name strings of the reference (application_commons.jl:9); "UDE" = settings.bPerform_UDE (params = NN parameters)
ACTIVE_PARAM = {None: 0, "": 0, "none": 0, "zb": 1, "ManningN": 2, "Q": 3, "UDE": 4}
UDE_MAX_HIDDEN, UDE_MAX_WIDTH = 3, 8
ERR_NAMES = {1: "HG_ERR_ARG", 2: "HG_ERR_CUDA", 3: "HG_ERR_CONVEYANCE", 4: "HG_ERR_SOLVER", 5: "HG_ERR_STATE", 6: "HG_ERR_COMM"}


class MeshDesc(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_faces", C.c_int64), ("n_ghost", C.c_int64), ("ld", C.c_int64),
                ("index_base", C.c_int32), ("cell_nfaces", c_i64p), ("cell_faces", c_i64p),
                ("cell_neighbors", c_i64p), ("cell_normals", c_f64p), ("face_is_boundary", c_u8p),
                ("face_lengths", c_f64p), ("cell_areas", c_f64p), ("cell_centroids", c_f64p)]


class BcDesc(C.Structure):
    _fields_ = [("n_inletq", C.c_int64), ("n_exith", C.c_int64), ("n_wall", C.c_int64), ("n_symm", C.c_int64),
                ("bc_ptr", c_i64p), ("ghost_ids", c_i64p), ("internal_cells", c_i64p),
                ("outward_normals", c_f64p), ("face_lengths", c_f64p),
                ("n_halo", C.c_int64), ("halo_flip", c_u8p), ("halo_area", c_f64p)]


class FieldsDesc(C.Structure):
    _fields_ = [("g", C.c_double), ("k_n", C.c_double), ("h_small", C.c_double), ("riemann_solver", C.c_char_p),
                ("hstill", c_f64p), ("hstill_ghost", c_f64p), ("zb_cells", c_f64p), ("zb_ghost", c_f64p),
                ("S0_cells", c_f64p), ("ManningN_cells", c_f64p), ("matID_cells", c_i64p), ("n_mat", C.c_int64),
                ("inletQ_TotalQ", c_f64p), ("exitH_WSE", c_f64p)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("tile_cells", C.c_int32), ("reorder", C.c_int32), ("strict", C.c_int32),
                ("path", C.c_int32), ("reserved", C.c_int32 * 11)]


class UdeDesc(C.Structure):
    _fields_ = [("choice", C.c_int32), ("n_hidden", C.c_int32), ("width", C.c_int32 * UDE_MAX_HIDDEN),
                ("activation", C.c_int32 * UDE_MAX_HIDDEN), ("layernorm", C.c_int32), ("ln_epsilon", C.c_double),
                ("h_bounds", C.c_double * 2), ("umag_bounds", C.c_double * 2), ("ks_bounds", C.c_double * 2),
                ("output_bounds", C.c_double * 2), ("n_params", C.c_int64),
                ("off_weight", C.c_int64 * (UDE_MAX_HIDDEN + 1)), ("off_bias", C.c_int64 * (UDE_MAX_HIDDEN + 1)),
                ("off_ln_scale", C.c_int64 * UDE_MAX_HIDDEN), ("off_ln_bias", C.c_int64 * UDE_MAX_HIDDEN)]


class NamedArray(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", c_f64p)]


# every symbol include/hydrograd_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "hg_default_options": (None, [C.POINTER(Options)]),
    "hg_abi_version": (C.c_int, []),
    "hg_create": (C.c_int, [C.POINTER(_vp), C.POINTER(MeshDesc), C.POINTER(BcDesc), C.POINTER(FieldsDesc),
                            C.POINTER(Options)]),
    "hg_destroy": (None, [_vp]),
    "hg_last_error": (C.c_char_p, [_vp]),
    "hg_n_cells": (C.c_int64, [_vp]),
    "hg_set_fields": (C.c_int, [_vp, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_set_manning_function": (C.c_int, [_vp, C.c_int32, c_f64p, c_f64p]),
    "hg_rhs": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, c_f64p]),
    "hg_rhs_vjp": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_rhs_jvp": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_rhs_jvp_multi": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, C.c_int64, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_solve_tsit5_sens": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                      C.c_double, C.c_double, c_f64p, C.c_int64, c_f64p, c_f64p, c_f64p, c_i64p]),
    "hg_set_state": (C.c_int, [_vp, c_f64p]),
    "hg_get_state": (C.c_int, [_vp, c_f64p]),
    "hg_set_params": (C.c_int, [_vp, c_f64p, C.c_int64, C.c_int32]),
    "hg_rhs_resident": (C.c_int, [_vp]),
    "hg_rhs_resident_phase": (C.c_int, [_vp, C.c_int32]),
    "hg_solve_tsit5": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, c_f64p, C.c_int64, c_f64p,
                                 c_i64p]),
    "hg_solve_tsit5_dense": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, c_f64p, C.c_int64, c_f64p,
                                 c_i64p]),
    "hg_vjp_resident_phase": (C.c_int, [_vp, C.c_int32]),
    "hg_get_rhs": (C.c_int, [_vp, c_f64p]),
    "hg_sync": (C.c_int, [_vp]),
    "hg_step_euler": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "hg_step_rk4": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "hg_euler_adjoint": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, C.c_int64, c_f64p, c_f64p, c_f64p,
                                   c_f64p]),
    "hg_rk_adjoint": (C.c_int, [_vp, C.c_int32, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, C.c_int64, c_f64p, c_f64p, c_f64p,
                                c_f64p]),
    "hg_rk_adjoint_steps": (C.c_int, [_vp, C.c_int32, c_f64p, c_f64p, C.c_int64, C.c_int32, c_f64p, C.c_int64, c_f64p, c_f64p, c_f64p,
                                      c_f64p]),
    "hg_last_steps": (C.c_int, [_vp, c_f64p, C.c_int64, c_i64p]),
    "hg_custom_ode_solve": (C.c_int, [_vp, c_f64p, c_f64p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double,
                                      c_f64p, C.c_int64, c_i64p]),
    "hg_halo_info": (C.c_int, [_vp, c_i64p, c_i64p]),
    "hg_halo_counts": (C.c_int, [_vp, c_i64p]),
    "hg_halo_buffers": (C.c_int, [_vp, C.POINTER(c_f64p), C.POINTER(c_f64p), c_i64p]),
    "hg_halo_pack": (C.c_int, [_vp, C.c_int32]),
    "hg_set_stream": (C.c_int, [_vp, _vp]),
    "hg_set_lambda": (C.c_int, [_vp, c_f64p]),
    "hg_vjp_resident": (C.c_int, [_vp]),
    "hg_get_vjp": (C.c_int, [_vp, c_f64p, c_f64p, c_f64p]),
    "hg_ensemble_alloc": (C.c_int, [_vp, C.c_int64, C.c_int32]),
    "hg_ensemble_set_member": (C.c_int, [_vp, C.c_int64, c_f64p, c_f64p, C.c_int64, C.c_int32]),
    "hg_ensemble_step_euler": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "hg_ensemble_rhs": (C.c_int, [_vp]),
    "hg_ensemble_get_member": (C.c_int, [_vp, C.c_int64, C.c_int32, c_f64p]),
    "hg_time_ensemble": (C.c_int, [_vp, C.c_int32, C.c_double, C.POINTER(C.c_float)]),
    "hg_partition_rcb": (C.c_int, [C.c_int64, c_f64p, c_f64p, C.c_int32, C.c_int64, c_i64p, c_i64p, C.POINTER(C.c_int32)]),
    "hg_partition_extract": (C.c_int, [C.POINTER(_vp), C.POINTER(MeshDesc), C.POINTER(BcDesc), C.POINTER(FieldsDesc), C.POINTER(C.c_int32), C.c_int32, c_i64p,
                                       C.c_char_p, C.c_int64]),
    "hg_case_load_srh2d": (C.c_int, [C.POINTER(_vp), C.c_char_p, C.c_char_p, C.c_int64]),
    "hg_case_free": (None, [_vp]),
    "hg_case_dims": (C.c_int, [_vp, c_i64p]),
    "hg_case_array": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), c_i64p, C.POINTER(C.c_int32)]),
    "hg_comm_export": (C.c_int, [_vp, C.c_void_p]),
    "hg_comm_connect": (C.c_int, [_vp, C.c_int64, C.c_void_p, c_i64p, c_i64p]),
    "hg_comm_init_shm": (C.c_int, [_vp, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "hg_comm_set_auto": (C.c_int, [_vp, C.c_int32]),
    "hg_comm_exchange": (C.c_int, [_vp, C.c_int32]),
    "hg_comm_disconnect": (C.c_int, [_vp]),
    "hg_comm_set_allreduce": (C.c_int, [_vp, ALLREDUCE_FN, C.c_void_p]),
    "hg_debug_math": (C.c_int, [_vp, C.c_int32, C.c_int64, c_f64p, c_f64p]),
    "hg_time_jvp": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "hg_time_rhs": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_float)]),
    "hg_time_vjp": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_float)]),
    "hg_kernel_launches": (C.c_int64, [_vp]),
    "hg_state_generation": (C.c_int64, [_vp]),
    "hg_mesh_stats": (C.c_int, [_vp, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "hg_plan_stats": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(BcDesc), C.POINTER(FieldsDesc), C.POINTER(Options),
                                c_i64p, c_i64p]),
    "hg_plan_pipeline": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(BcDesc), C.POINTER(FieldsDesc), C.POINTER(Options),
                                   c_i64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hg_plan_open": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(BcDesc), C.POINTER(FieldsDesc), C.POINTER(Options), C.POINTER(_vp)]),
    "hg_plan_close": (None, [_vp]),
    "hg_plan_array": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), c_i64p, C.POINTER(C.c_int32)]),
    "hg_debug_chunk_rows": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.c_int64, C.c_int64, c_i64p, c_i64p]),
    "hg_flush_l2": (C.c_int, [_vp]),
    "hg_set_ude_model": (C.c_int, [_vp, C.POINTER(UdeDesc), c_f64p]),
    "hg_set_controller_pow": (C.c_int, [_vp, C.c_int32]),
    "hg_fastpow": (C.c_double, [C.c_double, C.c_double]),
    "hg_step_ode_euler": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "hg_step_ab3": (C.c_int, [_vp, C.c_double, C.c_int64, C.c_int32]),
    "hg_format_f64": (C.c_int, [C.c_double, C.c_int32, C.c_char_p, C.c_int64]),
    "hg_json_open": (C.c_int, [C.POINTER(_vp), C.c_char_p, C.c_int32, C.c_char_p, C.c_int64]),
    "hg_json_error": (C.c_char_p, [_vp]),
    "hg_json_key": (C.c_int, [_vp, C.c_char_p]),
    "hg_json_begin_array": (C.c_int, [_vp]),
    "hg_json_end_array": (C.c_int, [_vp]),
    "hg_json_numbers": (C.c_int, [_vp, c_f64p, C.c_int64]),
    "hg_json_number": (C.c_int, [_vp, C.c_double]),
    "hg_json_string": (C.c_int, [_vp, C.c_char_p]),
    "hg_json_close": (C.c_int, [_vp, C.c_int32]),
    "hg_write_vtk_2d": (C.c_int, [C.c_char_p, C.c_int64, c_f64p, C.c_int64, C.c_int64, C.c_int32, c_i64p, c_i64p, C.c_char_p,
                                  C.c_char_p, C.c_double, C.POINTER(NamedArray), C.c_int64, C.POINTER(NamedArray), C.c_int64,
                                  C.c_char_p, C.c_int64]),
    "hg_forward_truth_fields": (C.c_int, [C.c_int64, c_f64p, c_f64p, c_f64p, c_f64p, C.c_double, C.c_double, C.c_double,
                                          c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_manning_function_cells": (C.c_int, [C.c_int32, c_f64p, C.c_int64, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]),
    "hg_dry_wet_flags": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, c_i64p, c_i64p, c_i64p, c_u8p, C.c_int64, c_f64p, c_f64p,
                                   C.c_double, c_u8p, c_u8p, c_u8p]),
    "hg_total_water_volume": (C.c_double, [C.c_int64, c_f64p, c_f64p]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python hydrograd.jl_b200/csrc/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
