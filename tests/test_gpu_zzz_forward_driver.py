"""The reference's forward-simulation driver in one call (hydrograd.jl_b200/forward.py::run_forward_case): from the case
directory -- run_control.json, SRH-2D files, initial condition -- through the device-resident adaptive Tsit5 to the files the
reference writes (forward_simulation_solution_truth.json, VTK, total_water_volume.csv), compared with the reference's own
committed truth files (BASELINE config 1).  (Written after the round's GPU budget was spent: not yet run on a B200.)"""
import json
import os

import numpy as np
import pytest

import _pkg
from tests import cases
from tests.test_results_cpu import _case_dir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    hg = _pkg.load()
    from hydrograd_jl_b200 import forward, results
    return hg, forward, results


@pytest.mark.parametrize("name,truth,tol", [("savannah", "savannah", 1e-5), ("oneD_bump", "oneD_bump", 1e-3)])
def test_forward_case_end_to_end(mods, tmp_path, name, truth, tol):
    # tolerances: the Savannah truth file is reproduced to 1e-9 by the host restatement of this run (tests/test_oracle_golden.py);
    # the channel's file comes from an older version of the reference (it still has S0_faces_truth) and the same restatement is
    # 6e-5 (xi) / 2e-4 (u) away from it -- a soft pin
    hg, forward, results = mods
    d = _case_dir(tmp_path, name, results)
    out = forward.run_forward_case(d, write_vtk=True)
    t = cases.truth(truth)
    got = json.load(open(os.path.join(d, "forward_simulation_solution_truth.json")))
    assert list(got.keys()) == list(results.TRUTH_KEYS)
    N = out["flat"]["n_cells"]
    err = {k: float(np.abs(np.asarray(got[k]) - t[k]).max()) for k in ("xi_truth", "h_truth", "wse_truth", "u_truth", "v_truth")}
    print(name, "forward driver vs the reference's truth file:", {k: "%.1e" % v for k, v in err.items()}, out["stats"])
    assert max(err.values()) <= tol
    for k in ("zb_cell_truth", "S0_cells_truth", "hstill_truth", "wstill_truth", "ManningN_cells_truth", "ManningN_zone_values_truth",
              "inlet_discharges_truth"):
        assert np.array_equal(np.asarray(got[k], dtype=np.float64), t[k]), k          # inputs and geometry: to the bit
    fr = np.abs(np.asarray(got["friction_x_truth"]) - t["friction_x_truth"]).max()
    assert fr <= 10 * tol * max(np.abs(t["friction_x_truth"]).max(), 1e-300) + 1e-12
    n_save = out["states"].shape[0]
    assert n_save == 101 and out["t_save"][0] == 0.0
    vtk = sorted(f for f in os.listdir(d) if f.endswith(".vtk"))
    assert len(vtk) == n_save and vtk[0] == "forward_simulation_results_0001.vtk"
    vol = open(os.path.join(d, "total_water_volume.csv")).read().split("\n")
    assert vol[0] == "total_water_volume" and len(vol) == n_save + 2
    assert np.isfinite(out["states"]).all()


def test_forward_case_with_state_dependent_manning(mods, tmp_path):
    """Savannah_River_ManningN_ks_h_Umag: Cheng's n(h, |U|, ks) inside every RHS, ks per material zone from run_control.json."""
    import shutil
    hg, forward, results = mods
    d = _case_dir(tmp_path, "savannah", results)
    shutil.copyfile(os.path.join(cases.GOLD, "savannah_ks", "run_control.json"), os.path.join(d, "run_control.json"))
    out = forward.run_forward_case(d, write_vtk=False)
    t = cases.truth("savannah_ks")
    got = json.load(open(os.path.join(d, "forward_simulation_solution_truth.json")))
    err = {k: float(np.abs(np.asarray(got[k]) - t[k]).max()) for k in ("xi_truth", "u_truth", "v_truth")}
    print("variable-n forward driver vs truth:", {k: "%.1e" % v for k, v in err.items()}, out["stats"])
    assert max(err.values()) <= 1e-5
    for k, tol in (("ManningN_cells_truth", 1e-4), ("h_ks_cells_truth", 1e-4), ("friction_factor_cells_truth", 1e-3), ("Re_cells_truth", 1e-3)):
        a, b = np.asarray(got[k]), t[k]
        assert np.abs(a - b).max() <= tol * np.abs(b).max(), k


def test_sensitivity_case_end_to_end(mods, tmp_path):
    """sensitivity_analysis/ManningN/Savana_River in one call: run_control.json -> hg_solve_tsit5_sens -> sensitivity_results.json
    and the six per-parameter files, against the reference's committed matrix."""
    import shutil
    hg, forward, results = mods
    from hydrograd_jl_b200 import sensitivity
    d = _case_dir(tmp_path, "savannah", results)
    shutil.copyfile(os.path.join(cases.GOLD, "savannah_sens", "run_control.json"), os.path.join(d, "run_control.json"))
    out = sensitivity.run_sensitivity_case(d, write_vtk=True)
    z = np.load(os.path.join(cases.GOLD, "savannah_sens", "sensitivity.npz"))
    assert np.array_equal(out["params_vector"], z["params_vector"])
    got = json.load(open(os.path.join(d, "sensitivity_results.json")))
    assert list(got.keys()) == list(results.SENSITIVITY_KEYS) and got["parameter_name"] == "ManningN"
    a, b = np.asarray(got["sensitivity_results"], dtype=np.float64), z["sensitivity_results"]
    print("sensitivity driver vs the reference's file: %.1e of %.1f" % (np.abs(a - b).max(), np.abs(b).max()), out["stats"])
    assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    for i in range(1, 7):
        assert os.path.exists(os.path.join(d, f"sensitivity_results_ManningN_{i}.json")) and os.path.exists(os.path.join(d, f"sensitivity_results_ManningN_{i}.vtk"))


@pytest.mark.parametrize("name,early_tol", [("oneD_uniform_sens", (2e-6, 2e-5)), ("oneD_bump_sens", (1e-6, 1e-5))])
def test_channel_sensitivity_case_reproduces_the_saved_trajectory(mods, tmp_path, name, early_tol):
    """The reference's channel sensitivity cases in one call: the VALUES of the Dual solve at the save times
    (forward_simulation_results.json -- the file the hard pins of tests/test_oracle_golden.py are made against) and
    d Q(T) / d ManningN (sensitivity_results.json), from the device-resident solve.  The runs are stability-limited (see
    tests/test_gpu_zzz_reference_replay.py): early saves tight, the final time as loose as the host restatement needs x 10."""
    hg, forward, results = mods
    from hydrograd_jl_b200 import sensitivity
    d = _case_dir(tmp_path, name, results)
    out = sensitivity.run_sensitivity_case(d, write_vtk=False)
    N = out["flat"]["n_cells"]
    tj = np.load(os.path.join(cases.GOLD, name, "trajectory.npz"))
    got = json.load(open(os.path.join(d, "forward_simulation_results.json")))
    assert list(got.keys()) == list(results.FORWARD_RESULTS_KEYS)
    pred = np.asarray(got["forward_simulation_results"], dtype=np.float64).reshape(101, 3 * N)      # column-major 3N x 101
    assert np.array_equal(pred[0], tj["forward_simulation_results"][0])                              # the initial condition
    idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
    err = [max(np.abs(pred[i][:N] - w[:N]).max(), np.abs(pred[i][N:2 * N] - w[N:2 * N]).max()) for i, w in zip(idx[:2], ref[:2])]
    fin = np.abs(pred[100][:2 * N] - tj["forward_simulation_results"][2][:2 * N]).max()
    z = np.load(os.path.join(cases.GOLD, name, "sensitivity.npz"))
    S_ref = z["sensitivity_results"].reshape(z["params_vector"].size, 3 * N)
    es = np.abs(out["sensitivity"].T - S_ref).max() / np.abs(S_ref).max()
    print(name, "early saves", ["%.1e" % e for e in err], "final %.1e" % fin, "sensitivities %.1e" % es, out["stats"])
    for e, tol in zip(err, early_tol):
        assert e <= tol
    assert fin <= 5e-4 and es <= 2e-4
    assert np.array_equal(np.asarray(got["zb_cells"], dtype=np.float64), tj["zb_cells"]) and np.array_equal(np.asarray(got["hstill"]), tj["hstill"])
