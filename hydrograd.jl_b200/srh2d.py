"""SRH-2D case loading for the host mirror: thin wrapper over the C++ reader / mesh builder in the library
(csrc/hg_srh.cpp) + the reference's initial-condition set-up.

  process_SRH_2D_input(case_path, srhhydro_file_name)   utilities/process_SRH_2D_input.jl:3-225 (+ mesh, BCs, bed,
                                                        Manning set-up of solve_swe_2D.jl:46-221), flattened
  setup_initial_condition(flat, ...)                    fvm/initial_conditions/process_ICs_2D.jl:48-87, 145-155
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .api import HydrogradError

_NAMES = ["cell_nfaces", "cell_faces", "cell_neighbors", "cell_nodes", "cell_normals", "face_lengths", "cell_areas",
          "cell_centroids", "bc_ptr", "bc_ghost_ids", "bc_internal_cells", "bc_normals", "bc_lengths", "zb_cells",
          "zb_ghost", "S0_cells", "ManningN_cells", "ManningN_zone", "matID_cells", "inletQ_TotalQ", "exitH_WSE",
          "node_coords", "face_is_boundary"]
_DT = {0: (np.float64, C.c_double), 1: (np.int64, C.c_int64), 2: (np.uint8, C.c_uint8)}


def process_SRH_2D_input(case_path, srhhydro_file_name):
    """Flat tables of the C ABI for an SRH-2D case (index_base = 1); fields still missing: hstill / hstill_ghost
    (they depend on the initial condition, see setup_initial_condition)."""
    lib = L.load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    path = os.path.join(case_path, srhhydro_file_name).encode()
    rc = lib.hg_case_load_srh2d(C.byref(h), path, err, 512)
    if rc:
        raise HydrogradError(rc, err.value.decode())
    try:
        dims = np.zeros(16, dtype=np.int64)
        lib.hg_case_dims(h, dims.ctypes.data_as(L.c_i64p))
        flat = dict(n_cells=int(dims[0]), n_faces=int(dims[1]), n_ghost=int(dims[2]), ld=int(dims[3]), index_base=int(dims[4]),
                    n_inletq=int(dims[5]), n_exith=int(dims[6]), n_wall=int(dims[7]), n_symm=int(dims[8]), n_mat=int(dims[9]),
                    g=9.81, k_n=1.0, h_small=1.0e-3)        # solve_swe_2D.jl:57-107 (SI)
        for name in _NAMES:
            ptr, cnt, dt = C.c_void_p(), C.c_int64(0), C.c_int32(0)
            rc = lib.hg_case_array(h, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(dt))
            if rc:
                raise HydrogradError(rc, f"array {name} missing")
            npt, ct = _DT[dt.value]
            if cnt.value == 0:
                flat[name] = np.zeros(0, dtype=npt)
            else:
                flat[name] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(cnt.value,)).astype(npt, copy=True)
        return flat
    finally:
        lib.hg_case_free(h)


def setup_initial_condition(flat, wse, wstill, q_x=0.0, q_y=0.0):
    """process_ICs_2D.jl:48-87 / 145-155: hstill = wstill - zb, xi = wse - wstill, ghost copies of hstill; returns
    Q0 = vcat(xi, q_x, q_y) (solve_swe_2D.jl:224).  Raises like the reference when wse < zb or h < h_small."""
    N = flat["n_cells"]
    zb = flat["zb_cells"]
    wse = np.broadcast_to(np.asarray(wse, dtype=np.float64), (N,)).copy()
    wstill = np.broadcast_to(np.asarray(wstill, dtype=np.float64), (N,)).copy()
    if (wse < zb).any():
        raise ValueError("The initial condition for WSE is smaller than bed elevation")
    if ((wse - zb) < flat["h_small"]).any():
        raise ValueError("The initial condition for water depth h is smaller than h_small")
    hstill = wstill - zb
    base = flat["index_base"]
    gh = np.empty(flat["n_ghost"])
    gh[flat["bc_ghost_ids"] - base] = hstill[flat["bc_internal_cells"] - base]
    flat["hstill"], flat["hstill_ghost"] = hstill, gh
    qx = np.broadcast_to(np.asarray(q_x, dtype=np.float64), (N,))
    qy = np.broadcast_to(np.asarray(q_y, dtype=np.float64), (N,))
    return np.concatenate([wse - wstill, qx, qy])
