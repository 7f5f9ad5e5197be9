"""The reference's forward-simulation driver, one call from the case directory to the result files (BASELINE config 1: the
examples/SWE_2D forward cases), with the time integration on the device.

  run_forward_case(case_path)       examples/SWE_2D/forward_simulation/*/run_case.jl -> Hydrograd.solve_swe_2D(...) with
                                    bPerform_Forward_Simulation: reads run_control.json (time_settings,
                                    forward_simulation_options), the SRH-2D files and the initial condition
                                    (applications/solve_swe_2D.jl:46-224, fvm/initial_conditions/process_ICs_2D.jl:48-99),
                                    integrates (applications/forward_simulation/swe_2D_forward_simulation.jl:2-95) and writes
                                    forward_simulation_solution_truth.json, forward_simulation_results_%04d.vtk and
                                    total_water_volume.csv (process_forward_simulation_results_2D.jl:4-86).

Only the keys the forward driver reads are interpreted; everything else in run_control.json (inversion, sensitivity, UDE
options, the .jld2 solution file) belongs to parts of the reference outside this path.  Reader, integrator and writers are the
library's (hg_srh.cpp, the device Tsit5 / Euler / RK4 / AB3 steppers, hg_results.cpp); there is no CPU fallback.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import results, srh2d
from .api import SWE2D_Extra_Parameters, swe_2D_consts


def _initial_condition(flat, fs, case_path):
    opt = fs["forward_simulation_initial_condition_options"]
    if opt == "constant":
        wse, wstill, qx, qy = (float(v) for v in fs["forward_simulation_initial_condition_constant_values"])
        return srh2d.setup_initial_condition(flat, wse, wstill, qx, qy), np.full(flat["n_cells"], wstill)
    if opt == "from_file":
        d = json.load(open(os.path.join(case_path, fs["forward_simulation_initial_condition_file_name"])))
        N = flat["n_cells"]
        for k in ("wse", "wstill", "q_x", "q_y"):
            if len(d[k]) != N:
                raise ValueError(f"The length of the initial condition values of {k} ({len(d[k])}) is not the same as the number of cells ({N}).")
        wstill = np.asarray(d["wstill"], dtype=np.float64)
        return srh2d.setup_initial_condition(flat, np.asarray(d["wse"], dtype=np.float64), wstill, np.asarray(d["q_x"], dtype=np.float64),
                                             np.asarray(d["q_y"], dtype=np.float64)), wstill
    raise ValueError(f"Invalid initial condition option: {opt}. Supported options: constant, from_file.")


def run_forward_case(case_path, out_path=None, device=0, write_vtk=True, controller_pow="fastpow", context_options=None):
    """Runs the forward simulation the case's run_control.json describes and writes its result files into out_path (default:
    the case directory, like the reference).  Returns dict(states [n_saves, 3N], t_save, stats, truth, flat).
    controller_pow: "fastpow" reproduces OrdinaryDiffEq's step sequence of the reference's committed runs."""
    out_path = case_path if out_path is None else out_path
    rc = json.load(open(os.path.join(case_path, "run_control.json")))
    if not rc["control_variables"].get("bPerform_Forward_Simulation", False):
        raise ValueError("run_control.json: bPerform_Forward_Simulation is not set (only the forward driver is mirrored here)")
    ts = rc["time_settings"]
    if ts.get("bUse_srhhydro_time_settings", False):
        raise ValueError("bUse_srhhydro_time_settings = true is not supported (the reference's own unit branch breaks on it, SURVEY 8f-1)")
    fs = rc["forward_simulation_options"]
    flat = srh2d.process_SRH_2D_input(case_path, rc["control_variables"]["srhhydro_file_name"])
    Q0, wstill = _initial_condition(flat, fs, case_path)
    t0, t1 = (float(v) for v in ts["tspan"])
    dt = float(ts["dt"])
    consts = swe_2D_consts(g=flat["g"], k_n=flat["k_n"], h_small=flat["h_small"], dt=dt, tspan=(t0, t1))
    extra = SWE2D_Extra_Parameters(flat, swe_2D_constants=consts, forward_settings=fs,
                                   options=dict(device=device, **(context_options or {})))
    ctx = extra.ctx
    ks_cells = None
    prm = fs.get("forward_simulation_ManningN_function_parameters", {})
    if fs.get("forward_simulation_ManningN_option", "constant") == "variable" and "ks" in prm:
        ks_cells = np.asarray(prm["ks"], dtype=np.float64)[np.asarray(flat["matID_cells"], dtype=np.int64)]
    n_save = int(fs["forward_simulation_nSave"])
    dt_save = (t1 - t0) / n_save
    t_save = t0 + dt_save * np.arange(n_save + 1)            # t_start:dt_save:t_end
    t_save[-1] = min(t_save[-1], t1)
    solver = fs["forward_simulation_solver"]
    stats = {}
    if solver == "SciML":
        ode = fs["forward_simulation_ode_solver"]
        ctx.set_state(Q0)
        if ode in ("Euler()", "RK4()", "AB3()"):
            stepper = {"Euler()": ctx.step_ode_euler, "RK4()": ctx.step_rk4, "AB3()": ctx.step_ab3}[ode]
            per = dt_save / dt
            if abs(per - round(per)) > 1e-9 * per:
                raise ValueError(f"{ode}: the save interval ({dt_save}) must be a multiple of dt ({dt})")
            states = [Q0.copy()]
            for _ in range(n_save):
                stepper(dt, int(round(per)))
                states.append(ctx.get_state())
            states = np.array(states)
            stats = dict(accepted=n_save * int(round(per)), rejected=0)
        elif ode == "Rosenbrock23()":
            raise ValueError("Rosenbrock23() is not available on the device (explicit solvers only: Tsit5(), Euler(), RK4(), AB3())")
        else:                                                 # "Tsit5()" and the reference's fall-through: adaptive Tsit5
            adaptive = bool(fs.get("forward_simulation_adaptive", True)) if ode == "Tsit5()" else True
            ctx.set_controller_pow(controller_pow)
            states, stats = ctx.solve_tsit5(t0, t1, dt, adaptive, 1e-6, 1e-3, t_save=t_save, saveat="interp")
    elif solver == "customized":                              # swe_2D_forward_simulation.jl:71-90: Euler loop, VTK only, no truth file
        sol = ctx.custom_ode_solve(Q0, None, None, t0, t1, dt)                # [3N, n_saves], every step saved
        if write_vtk:
            results.swe_2D_save_results_custom(flat, sol, out_path)
        return dict(states=np.ascontiguousarray(sol.T), t_save=t0 + dt * np.arange(sol.shape[1]), stats=dict(accepted=sol.shape[1] - 1, rejected=0),
                    truth=None, flat=flat)
    else:
        raise ValueError("Wrong solver choice. Supported solvers: SciML, customized. No forward simulation is performed.")
    truth = results.postprocess_forward_simulation_results_swe_2D(
        flat, states[-1], out_path, wstill, flat["ManningN_zone"], flat["inletQ_TotalQ"], forward_settings=fs, ks_cells=ks_cells,
        g=flat["g"], k_n=flat["k_n"], h_small=flat["h_small"],
        save_solution_truth_file_name=fs.get("forward_simulation_save_solution_truth_file_name", "forward_simulation_solution_truth.json"))
    if write_vtk:
        results.swe_2D_save_results_SciML(flat, states, out_path, wstill, truth["friction_x_truth"], truth["friction_y_truth"],
                                          forward_settings=fs, ks_cells=ks_cells, h_small=flat["h_small"])
    return dict(states=states, t_save=t_save, stats=stats, truth=truth, flat=flat)
