"""The reference's sensitivity-analysis driver, one call from the case directory to its result files, on the device.

  run_sensitivity_case(case_path)   solve_swe_2D with bPerform_Sensitivity_Analysis (applications/solve_swe_2D.jl:370-371) ->
                                    swe_2D_sensitivity (applications/sensitivity/swe_2D_sensitivity.jl:2-99): the active parameter
                                    and its values from run_control.json (parameters/process_model_parameters_2D.jl:93-127),
                                    ForwardDiff.jacobian of the adaptive Tsit5 solve = hg_solve_tsit5_sens (values and one partial
                                    per parameter on the device, Dual-aware error norm), then forward_simulation_results.json (the
                                    values at the save times), sensitivity_results.json and the
                                    per-parameter JSON / VTK files (process_sensitivity_results_2D.jl:4-80).

Only the keys this driver reads are interpreted.  Forward mode runs on the plain tables: the context is created with strict = 1.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import results, srh2d
from .api import Context

_VALUES_KEY = {"zb": "sensitivity_zb_values", "ManningN": "sensitivity_ManningN_values", "Q": "sensitivity_inlet_discharge_values"}
_FILE_KEY = {"zb": "zb_cells_param", "ManningN": "ManningN_list_param", "Q": "inlet_discharges_param"}


def _params_vector(opts, active, flat, case_path):
    how = opts["sensitivity_parameter_values_options"]
    if how == "constant":
        v = np.asarray(opts[_VALUES_KEY[active]], dtype=np.float64)
        return np.full(flat["n_cells"], v[0]) if active == "zb" else v.copy()
    if how == "from_file":
        d = json.load(open(os.path.join(case_path, opts["sensitivity_parameter_values_file_name"])))
        return np.asarray(d[_FILE_KEY[active]], dtype=np.float64)
    raise ValueError(f"Invalid parameter values option: {how}. Supported options: constant, from_file.")


def run_sensitivity_case(case_path, out_path=None, device=0, write_vtk=True, controller_pow="fastpow"):
    """Returns dict(Q_T [3N], sensitivity [3N, n_params], params_vector, stats, flat) and writes sensitivity_results.json +
    sensitivity_results_<name>_<i>.json / .vtk into out_path (default: the case directory, like the reference)."""
    out_path = case_path if out_path is None else out_path
    rc = json.load(open(os.path.join(case_path, "run_control.json")))
    if not rc["control_variables"].get("bPerform_Sensitivity_Analysis", False):
        raise ValueError("run_control.json: bPerform_Sensitivity_Analysis is not set")
    ts = rc["time_settings"]
    if ts.get("bUse_srhhydro_time_settings", False):
        raise ValueError("bUse_srhhydro_time_settings = true is not supported")
    opts = rc["sensitivity_analysis_options"]
    names = list(opts["active_param_names"])
    if len(names) != 1 or names[0] not in _VALUES_KEY:
        raise ValueError(f"Invalid active parameter name: {names}. Supported active parameter names: zb, ManningN, Q")
    active = names[0]
    ode = opts["sensitivity_ode_solver_options"]
    if ode["ode_solver"] != "Tsit5()":
        raise ValueError("Not implemented yet")                       # swe_2D_sensitivity.jl:50
    flat = srh2d.process_SRH_2D_input(case_path, rc["control_variables"]["srhhydro_file_name"])
    from .forward import _initial_condition
    ic = {k.replace("sensitivity_", "", 1): v for k, v in opts.items() if k.startswith("sensitivity_forward_simulation_initial_condition")}
    Q0, wstill = _initial_condition(flat, ic, case_path)
    p = _params_vector(opts, active, flat, case_path)
    ctx = Context(flat, device=device, strict=True)
    ctx.set_controller_pow(controller_pow)
    t0, t1 = (float(v) for v in ts["tspan"])
    n_save = int(ode.get("ode_solver_nSave", 0))
    t_save = t0 + (t1 - t0) / n_save * np.arange(n_save + 1) if n_save > 0 else np.zeros(0)      # t_start:dt_save:t_end
    if n_save > 0:
        t_save[-1] = min(t_save[-1], t1)
    QT, S, stats = ctx.solve_tsit5_sens(Q0, p, active, t0, t1, float(ts["dt"]), bool(ode.get("ode_solver_adaptive", True)), 1e-6, 1e-3,
                                        t_save=t_save)
    sens = np.ascontiguousarray(S.T)                                  # [3N, n_params], the reference's Jacobian layout
    pred = stats.pop("saves")
    results.save_sensitivity_results(out_path, pred_array=None if pred is None else pred.T, zb_cells=flat["zb_cells"], wstill=wstill,
                                     hstill=flat["hstill"], sensitivity=sens, parameter_name=active, params_vector=p)
    results.postprocess_sensitivity_results_swe_2D(flat, sens, p, active, out_path, write_vtk=write_vtk)
    return dict(Q_T=QT, sensitivity=sens, params_vector=p, stats=stats, flat=flat, pred=pred, t_save=t_save)
