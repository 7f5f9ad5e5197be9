"""Output side of the path: the result files of the reference's forward / sensitivity drivers, written by the C++ writers of
the library (csrc/hg_results.cpp), byte-compatible with the files the reference commits.

  format_float                                      how Julia prints a Float64 / what JSON3.pretty leaves of it
  write_json_pretty                                 JSON3.pretty(io, Dict(...)); println(io)
  export_to_vtk_2D                                  utilities/swe_2D_tools.jl:145-214 (same argument list)
  update_ManningN_forward_simulation                parameters/process_ManningN_2D.jl:102-213 (n, h_ks, f, Re)
  process_dry_wet_flags                             fvm/discretization/process_dry_wet.jl:2-35
  swe_2D_calc_total_water_volume                    utilities/swe_2D_tools.jl:4-7
  postprocess_forward_simulation_results_swe_2D     applications/forward_simulation/process_forward_simulation_results_2D.jl:4-86
  swe_2D_save_results_SciML                         utilities/swe_2D_tools.jl:10-90
  swe_2D_save_results_custom                        utilities/swe_2D_tools.jl:92-141
  save_sensitivity_results                          applications/sensitivity/swe_2D_sensitivity.jl:60-72, 84-95
  postprocess_sensitivity_results_swe_2D            applications/sensitivity/process_sensitivity_results_2D.jl:4-80

Host only (file output is I/O bound); the states come from the device integrators (Context.solve_tsit5 / custom_ode_solve).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .api import HydrogradError

FMT = {"julia": 0, "JSON3": 1}
MANNING_TYPES = {"power_law": 1, "sigmoid": 2, "inverse": 3, "h_Umag_ks": 4}

# Julia's Dict iteration order for the key sets the reference writes (read off its committed files; it is a property of the
# key set, not of the data)
TRUTH_KEYS = ("wstill_truth", "Re_cells_truth", "friction_y_truth", "xi_truth", "h_truth", "h_ks_cells_truth", "S0_cells_truth",
              "ManningN_cells_truth", "wse_truth", "inlet_discharges_truth", "ManningN_zone_values_truth", "v_truth",
              "zb_cell_truth", "friction_factor_cells_truth", "friction_x_truth", "u_truth", "hstill_truth")
SENSITIVITY_KEYS = ("params_vector", "parameter_name", "sensitivity_results")
FORWARD_RESULTS_KEYS = ("forward_simulation_results", "zb_cells", "wstill", "hstill")
SENSITIVITY_PARAM_KEYS = ("dh_dparam", "parameter_name", "parameter_value", "dhv_dparam", "dhu_dparam")


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a, t=L.c_f64p):
    return None if a is None else a.ctypes.data_as(t)


def format_float(x, style="julia"):
    """'julia': print(x) of a Float64; 'JSON3': the same number after JSON3.pretty (whole values as integers)."""
    buf = C.create_string_buffer(40)
    n = L.load().hg_format_f64(float(x), FMT[style], buf, 40)
    if n < 0:
        raise ValueError("hg_format_f64")
    return buf.value.decode()


class _Json:
    def __init__(self, path, style):
        self.lib = L.load()
        self.h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.hg_json_open(C.byref(self.h), os.fspath(path).encode(), FMT[style], err, 512)
        if rc:
            raise HydrogradError(rc, err.value.decode())

    def ck(self, rc):
        if rc:
            msg = self.lib.hg_json_error(self.h).decode()
            self.lib.hg_json_close(self.h, 0)
            self.h = None
            raise HydrogradError(rc, msg)

    def value(self, v):
        lib = self.lib
        if isinstance(v, str):
            self.ck(lib.hg_json_string(self.h, v.encode()))
        elif isinstance(v, np.ndarray) and v.ndim >= 1 and v.dtype.kind in "fiub":
            self.ck(lib.hg_json_begin_array(self.h))
            if v.ndim == 1:
                a = _f64(v)
                self.ck(lib.hg_json_numbers(self.h, _p(a), a.size))
            else:
                for row in v:
                    self.value(row)
            self.ck(lib.hg_json_end_array(self.h))
        elif isinstance(v, (list, tuple, np.ndarray)):
            if len(v) and all(isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool) for x in v):
                self.value(np.asarray(v, dtype=np.float64))
                return
            self.ck(lib.hg_json_begin_array(self.h))
            for x in v:
                self.value(x)
            self.ck(lib.hg_json_end_array(self.h))
        elif isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool):
            self.ck(lib.hg_json_number(self.h, float(v)))
        else:
            raise TypeError(f"write_json_pretty: unsupported value of type {type(v).__name__}")


def write_json_pretty(path, obj, style="JSON3", trailing_newline=True):
    """`open(path, "w") do io; JSON3.pretty(io, obj); println(io); end` for a dict of numbers, strings and (nested) arrays, keys
    in the order of `obj`.  1-D arrays are flat lists; an n-D numpy array nests in C order -- a Julia Matrix, which this JSON3
    writes flattened column-major, is passed as A.ravel(order="F").  NaN / Inf raise, as in JSON3."""
    w = _Json(path, style)
    try:
        for k, v in obj.items():
            w.ck(w.lib.hg_json_key(w.h, str(k).encode()))
            w.value(v)
    except BaseException:
        if w.h is not None:
            w.lib.hg_json_close(w.h, 0)
        raise
    rc = w.lib.hg_json_close(w.h, int(bool(trailing_newline)))
    if rc:
        raise HydrogradError(rc, f"write_json_pretty: closing {path} failed")


def _named(names, arrays, n_cells, cols):
    keep, arr = [], (L.NamedArray * max(len(names), 1))()
    for i, (nm, a) in enumerate(zip(names, arrays)):
        a = np.asarray(a, dtype=np.float64)
        if cols == 2:
            if a.shape != (n_cells, 2):
                raise ValueError(f"vector field {nm}: expected shape ({n_cells}, 2), got {a.shape}")
            a = np.ascontiguousarray(a.T).ravel()          # column-major n_cells x 2
        else:
            a = _f64(a).ravel()
            if a.size != n_cells:
                raise ValueError(f"scalar field {nm}: expected {n_cells} values, got {a.size}")
        nb = str(nm).encode()
        keep += [a, nb]
        arr[i].name = nb
        arr[i].data = _p(a)
    return arr, keep


def export_to_vtk_2D(filename, nodeCoordinates, cellNodesList, cellNodesCount, field_name, field_type, field_value, scalar_data,
                     scalar_names, vector_data, vector_names, index_base=1):
    """swe_2D_tools.jl:145-214.  nodeCoordinates [n_nodes, 3]; cellNodesList [N, ld] (ids from index_base); scalar_data: list of
    [N] arrays; vector_data: list of [N, 2] arrays.  Like the reference, invalid field_name / field_type / field_value print a
    message and write nothing."""
    if not isinstance(field_name, str) or not isinstance(field_type, str) or isinstance(field_value, bool) or \
            not isinstance(field_value, (int, float, np.integer, np.floating)):
        print("field_name, field_type, field_value are not valid")
        return False
    xyz = _f64(nodeCoordinates).reshape(-1, 3)
    cn = np.asarray(cellNodesList, dtype=np.int64)
    cnt = _i64(cellNodesCount)
    N, ld = cn.shape
    cn = np.ascontiguousarray(cn.T).ravel()                # N x ld column-major
    sc, k1 = _named(scalar_names, scalar_data, N, 1)
    vc, k2 = _named(vector_names, vector_data, N, 2)
    err = C.create_string_buffer(512)
    rc = L.load().hg_write_vtk_2d(os.fspath(filename).encode(), xyz.shape[0], _p(xyz), N, ld, int(index_base), _p(cn, L.c_i64p),
                                  _p(cnt, L.c_i64p), field_name.encode(), field_type.encode(), float(field_value), sc,
                                  len(scalar_names), vc, len(vector_names), err, 512)
    del k1, k2
    if rc:
        raise HydrogradError(rc, err.value.decode())
    return True


def forward_truth_fields(Q, hstill, wstill, ManningN_cells, g=9.81, k_n=1.0, h_small=1.0e-3):
    """process_forward_simulation_results_2D.jl:27-48: xi, wse, h, u, v and the friction terms of a saved state."""
    Q, hstill, wstill, n = _f64(Q), _f64(hstill), _f64(wstill), _f64(ManningN_cells)
    N = hstill.size
    if Q.size != 3 * N or wstill.size != N or n.size != N:
        raise ValueError("forward_truth_fields: array lengths")
    out = {k: np.empty(N) for k in ("xi", "wse", "h", "u", "v", "friction_x", "friction_y")}
    rc = L.load().hg_forward_truth_fields(N, _p(Q), _p(hstill), _p(wstill), _p(n), float(g), float(k_n), float(h_small),
                                          *[_p(out[k]) for k in ("xi", "wse", "h", "u", "v", "friction_x", "friction_y")])
    if rc:
        raise HydrogradError(rc, "hg_forward_truth_fields")
    return out


def update_ManningN_forward_simulation(h, Umag, ks, function_type, function_parameters=None):
    """(ManningN_cells, h_ks, friction_factor, Re) of process_ManningN_2D.jl:102-118 for ManningN_function_type in
    power_law | sigmoid | inverse | h_Umag_ks."""
    if function_type not in MANNING_TYPES:
        raise ValueError(f"Unknown Manning's n function type: {function_type}. Supported types: constant, power_law, sigmoid, inverse.")
    prm = function_parameters or {}
    pv = _f64([float(prm.get(k, 0.0)) for k in ("n_lower", "n_upper", "k", "h_mid")])
    h = _f64(h)
    N = h.size
    um = _f64(Umag) if Umag is not None else None
    ksa = _f64(ks) if ks is not None else None
    n, hk, f, Re = (np.empty(N) for _ in range(4))
    rc = L.load().hg_manning_function_cells(MANNING_TYPES[function_type], _p(pv), N, _p(h), _p(um), _p(ksa), _p(n), _p(hk), _p(f), _p(Re))
    if rc:
        raise HydrogradError(rc, "hg_manning_function_cells: bad arguments (k and h_mid must be positive; h_Umag_ks needs Umag and ks)")
    return n, hk, f, Re


def process_dry_wet_flags(flat, h, zb_cells, h_small=1.0e-3):
    """(b_dry_wet, b_Adjacent_to_dry_land, b_Adjacent_to_high_dry_land) as uint8 arrays; `flat` = the mesh tables."""
    N, ld = int(flat["n_cells"]), int(flat["ld"])
    nf, cf, nb = _i64(flat["cell_nfaces"]), _i64(flat["cell_faces"]), _i64(flat["cell_neighbors"])
    fb = np.ascontiguousarray(np.asarray(flat["face_is_boundary"], dtype=np.uint8))
    h, zb = _f64(h), _f64(zb_cells)
    out = [np.zeros(N, dtype=np.uint8) for _ in range(3)]
    rc = L.load().hg_dry_wet_flags(N, ld, int(flat.get("index_base", 1)), _p(nf, L.c_i64p), _p(cf, L.c_i64p), _p(nb, L.c_i64p),
                                   _p(fb, L.c_u8p), fb.size, _p(h), _p(zb), float(h_small), *[_p(o, L.c_u8p) for o in out])
    if rc:
        raise HydrogradError(rc, "hg_dry_wet_flags: inconsistent mesh tables")
    return tuple(out)


def swe_2D_calc_total_water_volume(h, cell_areas):
    h, a = _f64(h), _f64(cell_areas)
    return float(L.load().hg_total_water_volume(h.size, _p(h), _p(a)))


def _variable_manning(fields, ks_cells, forward_settings):
    fs = {k.replace("forward_simulation_", ""): v for k, v in (forward_settings or {}).items()}
    if fs.get("ManningN_option", "constant") != "variable":
        return None
    prm = {k: v for k, v in dict(fs.get("ManningN_function_parameters", {})).items() if k != "ks"}
    umag = np.sqrt(fields["u"] ** 2 + fields["v"] ** 2)
    return update_ManningN_forward_simulation(fields["h"], umag, ks_cells, fs["ManningN_function_type"], prm)


def postprocess_forward_simulation_results_swe_2D(flat, Q_final, case_path, wstill, ManningN_zone_values_truth, inlet_discharges_truth,
                                                  zb_cell_truth=None, forward_settings=None, ks_cells=None, g=9.81, k_n=1.0,
                                                  h_small=1.0e-3, save_solution_truth_file_name="forward_simulation_solution_truth.json"):
    """Writes forward_simulation_solution_truth.json from the final state, key for key what the reference writes
    (process_forward_simulation_results_2D.jl:54-75), and returns the dict.  Faithful to one detail of the reference: the
    friction terms use the STATIC ManningN_cells (line 45 passes swe2d_extra_params.ManningN_cells) even when Manning's n is
    'variable', while ManningN_cells_truth holds the closure's values."""
    N = int(flat["n_cells"])
    wstill = np.broadcast_to(_f64(wstill), (N,)).copy()
    n_static = _f64(flat["ManningN_cells"])
    fld = forward_truth_fields(Q_final, flat["hstill"], wstill, n_static, g, k_n, h_small)
    n_cells, h_ks, f, Re = n_static, np.zeros(N), np.zeros(N), np.zeros(N)
    var = _variable_manning(fld, ks_cells, forward_settings)
    if var is not None:
        n_cells, h_ks, f, Re = var
    S0 = _f64(flat["S0_cells"]).ravel()                      # [S0x(1:N); S0y(1:N)] = the N x 2 matrix column-major
    vals = {"wstill_truth": wstill, "Re_cells_truth": Re, "friction_y_truth": fld["friction_y"], "xi_truth": fld["xi"],
            "h_truth": fld["h"], "h_ks_cells_truth": h_ks, "S0_cells_truth": S0, "ManningN_cells_truth": n_cells,
            "wse_truth": fld["wse"], "inlet_discharges_truth": _f64(inlet_discharges_truth),
            "ManningN_zone_values_truth": _f64(ManningN_zone_values_truth), "v_truth": fld["v"],
            "zb_cell_truth": _f64(flat["zb_cells"] if zb_cell_truth is None else zb_cell_truth),
            "friction_factor_cells_truth": f, "friction_x_truth": fld["friction_x"], "u_truth": fld["u"],
            "hstill_truth": _f64(flat["hstill"])}
    out = {k: vals[k] for k in TRUTH_KEYS}
    write_json_pretty(os.path.join(case_path, save_solution_truth_file_name), out)
    return out


def swe_2D_save_results_SciML(flat, states, save_path, wstill, friction_x_truth, friction_y_truth, forward_settings=None,
                              ks_cells=None, h_small=1.0e-3):
    """One forward_simulation_results_%04d.vtk per saved state (18 scalars, U and slope vectors, FIELD = the save index) and
    total_water_volume.csv, as swe_2D_tools.jl:10-98.  `flat` needs node_coords / cell_nodes (process_SRH_2D_input)."""
    N, ld = int(flat["n_cells"]), int(flat["ld"])
    xyz = _f64(flat["node_coords"]).reshape(-1, 3)
    cn = _i64(flat["cell_nodes"]).reshape(ld, N).T
    cnt = _i64(flat["cell_nfaces"])                           # polygons: as many nodes as faces
    hstill, zb = _f64(flat["hstill"]), _f64(flat["zb_cells"])
    wstill = np.broadcast_to(_f64(wstill), (N,)).copy()
    S0 = _f64(flat["S0_cells"]).reshape(2, N).T
    n_static = _f64(flat["ManningN_cells"])
    ks = np.zeros(N) if ks_cells is None else _f64(ks_cells)
    volumes = []
    for index, state in enumerate(states, start=1):
        fld = forward_truth_fields(state, hstill, wstill, n_static, h_small=h_small)
        n_cells, h_ks, f, Re = n_static, np.zeros(N), np.zeros(N), np.zeros(N)
        var = _variable_manning(fld, ks_cells, forward_settings)
        if var is not None:
            n_cells, h_ks, f, Re = var
        dw, adj, high = process_dry_wet_flags(flat, fld["h"], zb, h_small)
        volumes.append(swe_2D_calc_total_water_volume(fld["h"], flat["cell_areas"]))
        state = _f64(state)
        scalars = [fld["xi"], wstill, hstill, fld["h"], state[N:2 * N], state[2 * N:], n_cells, ks, h_ks, f, Re, zb, fld["h"] + zb,
                   friction_x_truth, friction_y_truth, dw, adj, high]
        names = ["xi", "wstill", "hstill", "h", "hu", "hv", "ManningN", "ks", "h_ks", "friction_factor", "Re", "zb_cell", "WSE",
                 "friction_x", "friction_y", "b_dry_wet", "b_Adjacent_to_dry_land", "b_Adjacent_to_high_dry_land"]
        U = np.stack([fld["u"], fld["v"]], axis=1)
        export_to_vtk_2D(os.path.join(save_path, "forward_simulation_results_%04d.vtk" % index), xyz, cn, cnt,
                         "forward_simulation_saved_index", "integer", index, scalars, names, [U, S0], ["U", "slope"],
                         index_base=int(flat.get("index_base", 1)))
    with open(os.path.join(save_path, "total_water_volume.csv"), "w") as fo:
        fo.write("total_water_volume\n")
        for v in volumes:
            fo.write(format_float(v) + "\n")
    return volumes


def swe_2D_save_results_custom(flat, sol, save_path):
    """utilities/swe_2D_tools.jl:92-141, what the forward driver writes after the customized Euler solver: one VTK per column of
    `sol` ([3N, n_saves], the layout custom_ODE_solve returns) with the scalars h, hu, hv, zb_cell, WSE and the vector U.
    Faithful to the reference, including that it reads the first block of the state as the depth (it is xi = h - hstill): the
    file labels xi as "h", divides by it for U and adds zb to it for "WSE".  Returns the "volumes" sum(xi .* areas) it computes
    (the reference does not write them: its CSV block is commented out)."""
    sol = np.asarray(sol, dtype=np.float64)
    N, ld = int(flat["n_cells"]), int(flat["ld"])
    if sol.ndim != 2 or sol.shape[0] != 3 * N:
        raise ValueError(f"sol has shape {sol.shape}, expected ({3 * N}, n_saves)")
    xyz = _f64(flat["node_coords"]).reshape(-1, 3)
    cn = _i64(flat["cell_nodes"]).reshape(ld, N).T
    cnt = _i64(flat["cell_nfaces"])
    zb = _f64(flat["zb_cells"])
    volumes = []
    for index in range(1, sol.shape[1] + 1):
        Q = np.ascontiguousarray(sol[:, index - 1])
        first, hu, hv = Q[:N], Q[N:2 * N], Q[2 * N:]
        volumes.append(swe_2D_calc_total_water_volume(first, flat["cell_areas"]))
        with np.errstate(divide="ignore", invalid="ignore"):
            U = np.stack([hu / first, hv / first], axis=1)
        export_to_vtk_2D(os.path.join(save_path, "forward_simulation_results_%04d.vtk" % index), xyz, cn, cnt,
                         "forward_simulation_saved_index", "integer", index, [first, hu, hv, zb, first + zb],
                         ["h", "hu", "hv", "zb_cell", "WSE"], [U], ["U"], index_base=int(flat.get("index_base", 1)))
    return volumes


def save_sensitivity_results(case_path, pred_array=None, zb_cells=None, wstill=None, hstill=None, sensitivity=None,
                             parameter_name=None, params_vector=None):
    """forward_simulation_results.json (swe_2D_sensitivity.jl:60-72; pred_array [3N, n_saves]) and sensitivity_results.json
    (84-95; sensitivity [3N, n_params]); matrices flattened column-major like the reference's JSON3 writes them."""
    if pred_array is not None:
        write_json_pretty(os.path.join(case_path, "forward_simulation_results.json"),
                          {"forward_simulation_results": _f64(pred_array).ravel(order="F"), "zb_cells": _f64(zb_cells),
                           "wstill": _f64(wstill), "hstill": _f64(hstill)})
    if sensitivity is not None:
        write_json_pretty(os.path.join(case_path, "sensitivity_results.json"),
                          {"params_vector": _f64(params_vector), "parameter_name": str(parameter_name),
                           "sensitivity_results": _f64(sensitivity).ravel(order="F")})



def postprocess_sensitivity_results_swe_2D(flat, sensitivity, params_vector, active_param_name, case_path, write_vtk=True):
    """Per parameter i: sensitivity_results_<name>_<i>.json (dh / dhu / dhv _dparam, parameter_name, parameter_value) and the
    .vtk with the three scalars (process_sensitivity_results_2D.jl:33-78).  sensitivity: [3N, n_params], the reference's layout
    (hg_solve_tsit5_sens returns its transpose)."""
    S = np.asarray(sensitivity, dtype=np.float64)
    N = int(flat["n_cells"])
    p = _f64(params_vector)
    if S.shape != (3 * N, p.size):
        raise ValueError(f"sensitivity has shape {S.shape}, expected {(3 * N, p.size)}")
    for i in range(1, p.size + 1):
        dh, dhu, dhv = (np.ascontiguousarray(S[k * N:(k + 1) * N, i - 1]) for k in range(3))
        if write_vtk:
            ld = int(flat["ld"])
            export_to_vtk_2D(os.path.join(case_path, f"sensitivity_results_{active_param_name}_{i}.vtk"),
                             _f64(flat["node_coords"]).reshape(-1, 3), _i64(flat["cell_nodes"]).reshape(ld, N).T, _i64(flat["cell_nfaces"]),
                             "parameter_number", "integer", i, [dh, dhu, dhv], ["dh_dparam", "dhu_dparam", "dhv_dparam"], [], [],
                             index_base=int(flat.get("index_base", 1)))
        vals = {"dh_dparam": dh, "parameter_name": f"{active_param_name}_{i}", "parameter_value": float(p[i - 1]), "dhv_dparam": dhv,
                "dhu_dparam": dhu}
        write_json_pretty(os.path.join(case_path, f"sensitivity_results_{active_param_name}_{i}.json"),
                          {k: vals[k] for k in SENSITIVITY_PARAM_KEYS})
