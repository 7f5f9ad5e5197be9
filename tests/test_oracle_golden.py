"""CPU tests: pin the oracle against every fixture the reference commits for this path."""
import os

import numpy as np
import pytest

from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "oneD_uniform", "simple"])
def test_bed_elevation_and_slope_bit_exact(name, oracle_lib):
    """zb_cell_truth / S0_cells_truth pin mesh geometry + update_bed_data (a8) to the bit."""
    c, t = cases.load(name), cases.truth(name)
    assert np.array_equal(c.zb_cells, t["zb_cell_truth"])
    if "S0_cells_truth" in t.files:
        assert np.array_equal(c.S0_cells.ravel(order="F"), t["S0_cells_truth"])
        # the C++ restatement of update_bed_data gives the same bits as the numpy one
        zbg, S0 = Oracle(R.flatten(c)).bed(c.zb_cells)
        assert np.array_equal(S0, t["S0_cells_truth"])
        assert np.array_equal(zbg, c.zb_ghost)


def test_mesh_sizes_match_survey(oracle_lib):
    c = cases.load("savannah")
    assert (c.mesh.numOfCells, c.mesh.numOfFaces, c.mesh.numOfAllBounaryFaces) == (1306, 2737, 276)
    assert [len(b["faceIDs"]) for b in c.bc.kinds["inletQ"]] == [10]
    assert [len(b["faceIDs"]) for b in c.bc.kinds["exitH"]] == [10]
    assert sum(len(b["faceIDs"]) for b in c.bc.kinds["wall"]) == 256
    c = cases.load("oneD_bump")
    assert (c.mesh.numOfCells, c.mesh.numOfFaces, c.mesh.numOfAllBounaryFaces) == (200, 601, 402)
    c = cases.load("simple")   # two MONITORING node strings are skipped, default wall appended
    assert c.mesh.numOfCells == 12 and len(c.bc.kinds["wall"]) == 1


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "oneD_uniform"])
def test_friction_truth(name, oracle_lib):
    """friction_x/y_truth pin compute_friction_terms (a6) to <= 6e-16 relative."""
    c, t = cases.load(name), cases.truth(name)
    hs = c.h_small
    h = t["h_truth"]
    qx, qy = t["u_truth"] * (h + hs), t["v_truth"] * (h + hs)   # process_forward_simulation_results_2D.jl:32-33
    fx, fy = Oracle(R.flatten(c)).friction(h, qx, qy, t["ManningN_cells_truth"], c.g, c.k_n, hs)
    assert np.abs(fx - t["friction_x_truth"]).max() <= 6e-16 * np.abs(t["friction_x_truth"]).max()
    assert np.abs(fy - t["friction_y_truth"]).max() <= 6e-16 * max(np.abs(t["friction_y_truth"]).max(), 1e-30) + 1e-30


def test_steady_state_of_reference_trajectory(oracle_lib):
    """The reference's saved Tsit5 trajectory (sensitivity case, ManningN = [0.03,0.02,0.03]) converges to the
    oracle RHS's own steady state: residual at the last save ~2e-5, and explicit Euler integration of the
    oracle RHS from the same IC lands on the saved final state to ~1e-6 (soft pin of a1-a6 together)."""
    c = cases.load("oneD_bump_sens")
    tj = np.load(cases.GOLD + "/oneD_bump_sens/trajectory.npz")
    assert np.array_equal(tj["hstill"], c.hstill) and np.array_equal(tj["zb_cells"], c.zb_cells)
    traj = tj["forward_simulation_results"]
    assert np.array_equal(traj[0], c.Q0)
    o = Oracle(R.flatten(c))
    p = np.array([0.03, 0.02, 0.03])
    r = [np.abs(o.rhs(traj[k], p, 2)).max() for k in range(3)]
    assert r[0] > 1.0 and r[1] < 2e-3 and r[2] < 5e-5
    Q = o.euler(c.Q0, 0.01, 20000, p, 2)                       # t = 200 s
    assert np.abs(Q[:200] - traj[2][:200]).max() < 5e-6        # xi
    assert np.abs(Q[200:400] - traj[2][200:400]).max() < 5e-6  # q_x (|q_x| ~ 0.2)


def test_jvp_matches_finite_differences(oracle_lib):
    c = cases.load("savannah")
    o = Oracle(R.flatten(c))
    rng = np.random.default_rng(0)
    Q = cases.random_state(c, 1, dry_frac=0.0)
    v = rng.standard_normal(Q.size)
    p = c.ManningN_zone.copy()
    vp = rng.standard_normal(p.size) * 0.01
    _, jv = o.jvp(Q, v, p, vp, 2)
    e = 1e-7   # small enough that no wet/dry selector flips between the two evaluations
    fd = (o.rhs(Q + e * v, p + e * vp, 2) - o.rhs(Q - e * v, p - e * vp, 2)) / (2 * e)
    assert np.abs(jv - fd).max() <= 1e-6 * np.abs(fd).max()


def test_vjp_bruteforce_dot_identity(oracle_lib):
    c = cases.load("simple")
    o = Oracle(R.flatten(c))
    rng = np.random.default_rng(3)
    Q = cases.random_state(c, 2)
    lam, v = rng.standard_normal(Q.size), rng.standard_normal(Q.size)
    p = c.zb_cells.copy()
    vp = rng.standard_normal(p.size)
    Qbar, pbar = o.vjp_bruteforce(Q, lam, p, 1)
    _, jv = o.jvp(Q, v, p, vp, 1)
    lhs, rhs = lam @ jv, Qbar @ v + pbar @ vp
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)


@pytest.mark.slow
def test_sensitivity_fixture_pins_the_derivative(oracle_lib):
    """The reference's committed sensitivity_results.json (ForwardDiff through adaptive Tsit5, d(final state)/d(n_zone),
    sensitivity_analysis/ManningN/oneD_channel_with_bump) against forward sensitivities propagated through explicit
    Euler with the oracle's dual-number JVP: at the (near-steady) final time they agree to ~2e-5 relative.  This is the
    one reference fixture that pins the DERIVATIVE of the path (softly)."""
    c = cases.load("oneD_bump_sens")
    o = Oracle(R.flatten(c))
    S = np.load(cases.GOLD + "/oneD_bump_sens/sensitivity.npz")["sensitivity_results"].reshape(3, 600).T
    p = np.array([0.03, 0.02, 0.03])
    dt, nsteps = 0.02, 10000                      # t = 200 s, as in the case's run_control.json
    Q = c.Q0.copy()
    dQ = np.zeros((3, 600))
    eye = np.eye(3)
    for _ in range(nsteps):
        f = o.rhs(Q, p, 2)
        for k in (1, 2):                           # zone 0 (default material) owns no cell: its column is zero
            dQ[k] += dt * o.jvp(Q, dQ[k], p, eye[k], 2)[1]
        Q = Q + dt * f
    assert np.abs(S[:, 0]).max() == 0.0
    for k in (1, 2):
        assert np.abs(dQ[k] - S[:, k]).max() <= 1e-4 * np.abs(S[:, k]).max()


# ---------------------------------------------------------------- variable Manning's n (forward simulation option)
def _closure_inputs(t, h_small=1e-3):
    """The reference post-processes with u = q/(h + h_small) (process_forward_simulation_results_2D.jl:32-38)."""
    return t["h_truth"], np.sqrt(t["u_truth"] ** 2 + t["v_truth"] ** 2)


def test_manning_closure_h_Umag_ks_matches_reference_truth():
    """Cheng (2008) n(h, |U|, ks) (process_ManningN_2D.jl:181-213) against ManningN_cells_truth, Re_cells_truth,
    h_ks_cells_truth and friction_factor_cells_truth of Savannah_River_ManningN_ks_h_Umag."""
    import json
    d = os.path.join(cases.GOLD, "savannah_ks")
    t = np.load(os.path.join(d, "truth.npz"))
    rc = json.load(open(os.path.join(d, "run_control.json")))
    ks_zone = np.array(rc["forward_simulation_options"]["forward_simulation_ManningN_function_parameters"]["ks"])
    c = cases.load("savannah")                     # same mesh files (byte-identical in the reference)
    ks_cells = ks_zone[c.matID]                    # process_SRH_2D_input.jl:159-164, process_ManningN_2D.jl:56-60
    h, U = _closure_inputs(t)
    o = Oracle(R.flatten(c))
    got = o.manning_closure("h_Umag_ks", h, U, ks_cells)
    moving = U > 0                                 # still cells: Re = 0 -> f = Inf -> NaN, written as 0 by replace_nan
    for key, ref in (("n", "ManningN_cells_truth"), ("Re", "Re_cells_truth"), ("h_ks", "h_ks_cells_truth"), ("f", "friction_factor_cells_truth")):
        a, b = got[key][moving], t[ref][moving]
        ok = np.isfinite(a) & np.isfinite(b) & (b != 0)
        assert ok.sum() > 1000
        assert np.abs(a[ok] / b[ok] - 1).max() <= 1e-12, key


def test_manning_closure_sigmoid_matches_reference_truth():
    """sigmoid n(h) (process_ManningN_2D.jl:173-186) against ManningN_cells_truth of oneD_channel_with_bump_ManningN_h."""
    import json
    d = os.path.join(cases.GOLD, "oneD_bump_nh")
    t = np.load(os.path.join(d, "truth.npz"))
    p = json.load(open(os.path.join(d, "run_control.json")))["forward_simulation_options"]["forward_simulation_ManningN_function_parameters"]
    o = Oracle(R.flatten(cases.load("oneD_bump")))
    got = o.manning_closure("sigmoid", t["h_truth"], n_lower=p["n_lower"], n_upper=p["n_upper"], k=p["k"], h_mid=p["h_mid"])["n"]
    assert np.abs(got / t["ManningN_cells_truth"] - 1).max() <= 1e-14


def test_manning_closures_power_law_and_inverse_formulas():
    """n_lower + (n_upper - n_lower) (h + eps)^-k and n_lower + (n_upper - n_lower)/(1 + k h) (:140-168): restated directly."""
    o = Oracle(R.flatten(cases.load("simple")))
    h = np.exp(np.random.default_rng(0).uniform(np.log(1e-3), np.log(10), 500))
    nl, nu, k = 0.02, 0.08, 0.7
    assert np.allclose(o.manning_closure("power_law", h, n_lower=nl, n_upper=nu, k=k)["n"], nl + (nu - nl) * (h + np.finfo(float).eps) ** (-k), rtol=1e-15)
    assert np.allclose(o.manning_closure("inverse", h, n_lower=nl, n_upper=nu, k=k)["n"], nl + (nu - nl) / (1 + k * h), rtol=1e-15)


def test_tsit5_tableau():
    """The restated Tsit5 coefficients: row sums = c, order-5 conditions of the propagating weights, order-4 of the
    embedded ones, error weights summing to zero (all to rounding) -- the tableau is the published one."""
    from tests import tsit5_ref as T
    A = np.zeros((7, 7))
    for i, row in enumerate(T.A):
        A[i, :len(row)] = row
    c, b, bt = np.array(T.C), A[6].copy(), np.array(T.BTILDE)
    assert np.abs(A.sum(1) - c).max() < 1e-15
    for p in range(5):
        assert abs((b * c ** p).sum() - 1 / (p + 1)) < 1e-15
    for val, ref in ((b @ A @ c, 1 / 6), (b @ A @ c ** 2, 1 / 12), (b @ A @ A @ c, 1 / 24), (b @ A @ c ** 3, 1 / 20),
                     (b @ A @ A @ A @ c, 1 / 120), ((b * c) @ A @ c, 1 / 8), ((b * c) @ A @ c ** 2, 1 / 15), (b @ (A @ c) ** 2, 1 / 20)):
        assert abs(val - ref) < 1e-15
    bh = b - bt
    for p in range(4):
        assert abs((bh * c ** p).sum() - 1 / (p + 1)) < 1e-15
    assert abs(bt.sum()) < 1e-15


def test_transient_of_reference_trajectory_with_tsit5(oracle_lib):
    """The reference's saved Tsit5 trajectory (adaptive, abstol 1e-6, reltol 1e-3, dt0 = 0.02, saves 2 s apart) during the
    TRANSIENT (t = 2 ... 60 s, where the state still moves by O(0.1)): the oracle RHS integrated by the restated Tsit5 +
    PI controller stays within the integration tolerance of it.  Soft pin of a1-a6 off the steady state."""
    from tests import tsit5_ref as T
    c = cases.load("oneD_bump_sens")
    tj = np.load(cases.GOLD + "/oneD_bump_sens/trajectory.npz")
    idx = tj["early_index"]
    ref = tj["forward_simulation_results_early"]
    o = Oracle(R.flatten(c))
    p = np.array([0.03, 0.02, 0.03])
    t_save = 2.0 * idx
    N = 200
    moved = np.abs(ref[0][:N] - c.Q0[:N]).max()
    assert moved > 0.02                                          # a real transient: xi moves by ~0.1 m
    # (a) the reference's own settings: both solutions carry the tolerance's error
    _, saves, st = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, 0.0, float(t_save[-1]), 0.02, True, 1e-6, 1e-3, t_save)
    assert st["accepted"] > 100 and st["rejected"] < st["accepted"]
    for k, (got, want) in enumerate(zip(saves, ref)):
        assert np.abs(got[:N] - want[:N]).max() < 1e-3 and np.abs(got[N:2 * N] - want[N:2 * N]).max() < 1e-3, int(idx[k])
    # (b) a tight solve: what is left is the reference's integration error (measured 7e-5 at t = 2 s falling to 5e-7 at 60 s)
    _, saves, st = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, 0.0, float(t_save[-1]), 0.02, True, 1e-9, 1e-7, t_save)
    for k, (got, want) in enumerate(zip(saves, ref)):
        lim = 4e-4 if idx[k] < 20 else 1e-5
        assert np.abs(got[:N] - want[:N]).max() < lim and np.abs(got[N:2 * N] - want[N:2 * N]).max() < lim, int(idx[k])


def test_tsit5_dense_output_conditions():
    """Tsit5's dense output b_i(theta) (OrdinaryDiffEq's Tsit5Interp, restated in tsit5_ref.INTERP): the continuous
    order-4 conditions hold identically in theta (coefficient by coefficient), b_i(1) are the weights of the fifth-order
    solution and b_i(0) = 0 -- the interpolant is the published one."""
    from tests import tsit5_ref as T
    A = np.zeros((7, 7))
    for i, row in enumerate(T.A):
        A[i, :len(row)] = row
    c = np.array(T.C)
    P = np.array(T.INTERP)                                       # b_i(theta) = sum_p P[i, p] theta^(p+1)
    for w, power, val in ((np.ones(7), 1, 1.0), (c, 2, 1 / 2), (c ** 2, 3, 1 / 3), (A @ c, 3, 1 / 6), (c ** 3, 4, 1 / 4),
                          (c * (A @ c), 4, 1 / 8), (A @ c ** 2, 4, 1 / 12), (A @ A @ c, 4, 1 / 24)):
        want = np.zeros(4)
        want[power - 1] = val
        assert np.abs(P.T @ w - want).max() < 5e-14
    assert np.abs(P.sum(1) - A[6]).max() < 5e-15
    assert np.allclose(T.interp_weights(1.0), A[6], atol=5e-15) and T.interp_weights(0.0) == [0.0] * 7
    # on u' = -u the interpolated saves have the dense output's O(h^5) local accuracy, and the step sequence does not
    # depend on the save times (OrdinaryDiffEq's saveat semantics)
    f = lambda u: -u
    _, sv, st = T.solve(f, np.array([1.0]), 0.0, 2.0, 0.1, True, 1e-8, 1e-8, [0.3, 0.77, 1.234, 2.0], saveat="interp")
    _, _, st0 = T.solve(f, np.array([1.0]), 0.0, 2.0, 0.1, True, 1e-8, 1e-8, [])
    assert st == st0
    for s, t in zip(sv, [0.3, 0.77, 1.234, 2.0]):
        assert abs(s[0] - np.exp(-t)) < 2e-7


def test_transient_of_reference_trajectory_with_dense_saveat(oracle_lib):
    """As test_transient_of_reference_trajectory_with_tsit5, with OrdinaryDiffEq's saveat semantics (no stops, dense
    output): this is the reference's own step sequence up to the starting guess of rounding."""
    from tests import tsit5_ref as T
    c = cases.load("oneD_bump_sens")
    tj = np.load(cases.GOLD + "/oneD_bump_sens/trajectory.npz")
    idx = tj["early_index"]
    ref = tj["forward_simulation_results_early"]
    o = Oracle(R.flatten(c))
    p = np.array([0.03, 0.02, 0.03])
    t_save = 2.0 * idx
    N = 200
    _, saves, st = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, 0.0, float(t_save[-1]), 0.02, True, 1e-6, 1e-3, t_save, saveat="interp")
    err = [max(np.abs(g[:N] - w[:N]).max(), np.abs(g[N:2 * N] - w[N:2 * N]).max()) for g, w in zip(saves, ref)]
    print("dense saveat vs reference:", ["%.1e" % e for e in err], st)
    assert len(saves) == len(ref) and max(err) < 1e-3


def _dual_rhs(o, p, K):
    """The RHS on a vector of ForwardDiff.Dual numbers, stored as [1 + K, 3N]: values and the K partials (one dual-number
    pass of the oracle per partial)."""
    def rhs(U):
        out = np.empty_like(U)
        for k in range(K):
            e = np.zeros(K)
            e[k] = 1.0
            f, jv = o.jvp(U[0], U[1 + k], p, e, 2)
            out[1 + k] = jv
        out[0] = f
        return out
    return rhs


@pytest.mark.parametrize("name,p,dt_save,early_tol", [("oneD_bump_sens", [0.03, 0.02, 0.03], 2.0, (1e-8, 3e-9)),
                                                      ("oneD_uniform_sens", [0.03, 0.03], 1.0, (2e-10, 1e-9, 3e-9))])
def test_reference_trajectory_hard_pin(oracle_lib, name, p, dt_save, early_tol):
    """HARD pin of a1-a6 on the reference's own saved transients.  The trajectories in sensitivity_analysis/ManningN/*/
    forward_simulation_results.json are the values of a ForwardDiff.Dual solve (ForwardDiff.jacobian around `solve`,
    swe_2D_sensitivity.jl:34-80), so OrdinaryDiffEq's error estimate -- hence every step size -- includes the partials
    (tsit5_ref.dual_norm).  Integrating the oracle RHS on values AND partials (dual-number JVP), with the restated Tsit5, PI
    controller (DiffEqBase's Float32 `fastpow`) and dense-output saveat, reproduces the reference's saved states to
    1e-11 ... 1e-9 over the first saves (~100 accepted steps x 7 stages of RHS calls on an evolving transient) and to 2e-6
    over the first 12 saves; later the Float32 rounding of the controller's powers (exp2 of this platform vs Julia's) lets
    the step sequences drift apart at the 1e-7 level.  Measured: bump 1.4e-9, 3.8e-10; uniform 1.4e-11, 8.9e-11, 3.0e-10."""
    from tests import tsit5_ref as T
    c = cases.load(name)
    tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
    idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
    assert np.array_equal(tj["forward_simulation_results"][0], c.Q0)             # the reference started from the same state
    o = Oracle(R.flatten(c))
    p = np.array(p)
    K, N = p.size, c.mesh.numOfCells
    U0 = np.zeros((1 + K, 3 * N))
    U0[0] = c.Q0
    n = int(np.searchsorted(idx, 12, side="right"))
    ts = dt_save * idx[:n]
    # (integrate a little past the last save: the solve's end point is a stop that the reference does not have there)
    _, saves, st = T.solve(_dual_rhs(o, p, K), U0, 0.0, float(ts[-1]) + 3 * dt_save, 0.02, True, 1e-6, 1e-3, ts, saveat="interp",
                           norm=T.dual_norm, pow="fastpow")
    err = [max(np.abs(s[0][:N] - w[:N]).max(), np.abs(s[0][N:2 * N] - w[N:2 * N]).max()) for s, w in zip(saves, ref)]
    print(name, "dual-norm Tsit5 vs the reference's saved trajectory:", ["%.1e" % e for e in err], st)
    assert np.abs(ref[0][:N] - c.Q0[:N]).max() > 1e-3                             # a real transient
    for e, tol in zip(err, early_tol):
        assert e <= tol
    assert max(err) <= 2e-6
    # the same solve with a real-valued error norm (what a plain forward run would do) is three orders further away
    _, saves_r, _ = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, 0.0, float(ts[1]) + 3 * dt_save, 0.02, True, 1e-6, 1e-3, ts[:2], saveat="interp",
                            pow="fastpow")
    assert np.abs(saves_r[0][:N] - ref[0][:N]).max() > 100 * err[0]


@pytest.mark.parametrize("name,p,t_end", [("oneD_bump_sens", [0.03, 0.02, 0.03], 200.0), ("oneD_uniform_sens", [0.03, 0.03], 100.0)])
def test_reference_final_state_and_sensitivity_through_the_whole_run(oracle_lib, name, p, t_end):
    """The whole reference run (2100 / 2800 accepted steps): final state within 5e-5 of the saved one and the partials at the
    final time -- d Q(T) / d ManningN, what ForwardDiff.jacobian returns into sensitivity_results.json -- within 2e-5 of the
    largest entry (measured 1e-5 / 3.5e-6): pins the dual-number derivative of the oracle through thousands of RHS calls."""
    from tests import tsit5_ref as T
    c = cases.load(name)
    tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
    S = np.load(cases.GOLD + f"/{name}/sensitivity.npz")["sensitivity_results"]
    o = Oracle(R.flatten(c))
    p = np.array(p)
    K, N = p.size, c.mesh.numOfCells
    S = S.reshape(K, 3 * N)
    U0 = np.zeros((1 + K, 3 * N))
    U0[0] = c.Q0
    U, _, st = T.solve(_dual_rhs(o, p, K), U0, 0.0, t_end, 0.02, True, 1e-6, 1e-3, (), saveat="interp", norm=T.dual_norm, pow="fastpow")
    assert st["accepted"] > 2000 and st["rejected"] < 30
    final = tj["forward_simulation_results"][2]
    assert np.abs(U[0][:2 * N] - final[:2 * N]).max() <= 5e-5
    for k in range(K):
        assert np.abs(U[1 + k] - S[k]).max() <= 2e-5 * max(np.abs(S).max(), 1e-30), k
    assert np.abs(S).max() > 0.5


def reference_step_sequence(name, p, dt_save, n_saves):
    """The step sizes of the reference's sensitivity run, recovered by the Dual solve of test_reference_trajectory_hard_pin."""
    from tests import tsit5_ref as T
    c = cases.load(name)
    o = Oracle(R.flatten(c))
    p = np.array(p)
    K, N = p.size, c.mesh.numOfCells
    U0 = np.zeros((1 + K, 3 * N))
    U0[0] = c.Q0
    steps = []
    T.solve(_dual_rhs(o, p, K), U0, 0.0, dt_save * (n_saves + 3), 0.02, True, 1e-6, 1e-3, (), saveat="interp", norm=T.dual_norm,
            pow="fastpow", record=steps)
    return steps


def test_replaying_the_recovered_step_sequence_with_values_only(oracle_lib):
    """The reference's saved values depend on the partials only through the step sizes: replaying the recovered (t, h)
    sequence with a plain value-only Tsit5 step + dense output gives the same saves.  (This is the logic
    tests/test_gpu_zzz_reference_replay.py runs with the CUDA RHS in place of the oracle.)"""
    from tests import tsit5_ref as T
    name, p, dt_save = "oneD_uniform_sens", np.array([0.03, 0.03]), 1.0
    c = cases.load(name)
    N = c.mesh.numOfCells
    ref = np.load(cases.GOLD + f"/{name}/trajectory.npz")["forward_simulation_results_early"]
    o = Oracle(R.flatten(c))
    steps = reference_step_sequence(name, p, dt_save, 3)
    state = {"u": c.Q0.copy()}

    def step(t0, t1, h, inside):
        u, saves, st = T.solve(lambda v: o.rhs(v, p, 2), state["u"], t0, t1, h, adaptive=False, t_save=inside, saveat="interp")
        assert st["accepted"] == 1
        state["u"] = u
        return saves

    got = T.replay(step, steps, dt_save * np.array([1, 2, 3]))
    assert len(got) == 3
    for g, w, tol in zip(got, ref, (2e-10, 1e-9, 3e-9)):
        assert max(np.abs(g[:N] - w[:N]).max(), np.abs(g[N:2 * N] - w[N:2 * N]).max()) <= tol


def _savannah_ks_cells():
    import json
    rc = json.load(open(os.path.join(cases.GOLD, "savannah_ks", "run_control.json")))
    ks_zone = np.array(rc["forward_simulation_options"]["forward_simulation_ManningN_function_parameters"]["ks"])
    return ks_zone[cases.load("savannah").matID]       # process_SRH_2D_input.jl:159-164, process_ManningN_2D.jl:56-60


def savannah_forward_steps(variable_n):
    """(oracle, recorded (t, h) of the accepted steps, final state) of the reference's Savannah forward simulation:
    solve(prob, Tsit5(), adaptive=true, dt=0.02, saveat=...) over [0, 200] s (forward_simulation/Savannah_River*/run_control.json)
    with a real-valued error norm and DiffEqBase's fastpow in the PI controller."""
    from tests import tsit5_ref as T
    c = cases.load("savannah")
    o = Oracle(R.flatten(c))
    if variable_n:
        o.set_manning_function("h_Umag_ks", ks_cells=_savannah_ks_cells())
    steps = []
    try:
        u, _, st = T.solve(lambda v: o.rhs(v), c.Q0, 0.0, 200.0, 0.02, True, 1e-6, 1e-3, (), saveat="interp", pow="fastpow", record=steps)
    finally:
        o.set_manning_function("constant")
    return steps, u, st


@pytest.mark.parametrize("variable_n", [False, True])
def test_savannah_forward_run_reproduces_the_reference_final_state(oracle_lib, variable_n):
    """HARD pin on the river mesh (1306 mixed triangles / quadrilaterals, six Manning zones, inlet-q / exit-h / walls, real
    bathymetry): the reference's committed final state of its 200 s forward simulations -- xi_truth, u_truth, v_truth of
    forward_simulation/Savannah_River (constant n) and Savannah_River_ManningN_ks_h_Umag (Cheng's n(h, |U|, ks) evaluated
    inside every RHS, semi_discretize_swe_2D.jl:140-149) -- is reproduced by the oracle RHS under the restated adaptive Tsit5
    to 2e-9 / 1e-9 after 202 accepted steps (1219 RHS calls).  With the exact power in the controller: 8e-8 / 1e-7."""
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    t = cases.truth("savannah_ks" if variable_n else "savannah")
    steps, u, st = savannah_forward_steps(variable_n)
    assert st["accepted"] > 150 and st["rejected"] <= 3
    den = u[:N] + flat["hstill"] + flat["h_small"]              # the reference saves u = q / (h + h_small)
    err = (np.abs(u[:N] - t["xi_truth"]).max(), np.abs(u[N:2 * N] / den - t["u_truth"]).max(), np.abs(u[2 * N:] / den - t["v_truth"]).max())
    print("savannah forward run vs truth (xi, u, v):", ["%.1e" % e for e in err], st)
    assert np.abs(u[:N] - c.Q0[:N]).max() > 1e-3                 # the state did move away from the initial condition
    assert max(err) <= 1e-8


def savannah_sensitivity_solve(jvp=None):
    """The reference's Savannah sensitivity run restated: values and six partials through the Dual-norm Tsit5.  Returns
    (U[1 + K, 3N] at T = 200 s, the reference's sensitivity_results as [K, 3N], stats, recorded accepted steps [(t, h)]).
    jvp(Q, V, p, pdot) -> (f, J_Q V + J_p pdot) replaces the oracle's dual pass (the device's or the host build of its source)."""
    from tests import tsit5_ref as T
    c = cases.load("savannah")
    o = Oracle(R.flatten(c))
    z = np.load(os.path.join(cases.GOLD, "savannah_sens", "sensitivity.npz"))
    p = z["params_vector"]
    K, N = p.size, c.mesh.numOfCells
    S = z["sensitivity_results"].reshape(K, 3 * N)
    assert np.array_equal(p, c.ManningN_zone)
    U0 = np.zeros((1 + K, 3 * N))
    U0[0] = c.Q0

    def rhs(U):
        out = np.empty_like(U)
        for k in range(K):
            e = np.zeros(K)
            e[k] = 1.0
            f, jv = o.jvp(U[0], U[1 + k], p, e, 2, nthreads=0) if jvp is None else jvp(U[0], U[1 + k], p, e)
            out[1 + k] = jv
        out[0] = f
        return out

    steps = []
    U, _, st = T.solve(rhs, U0, 0.0, 200.0, 0.02, True, 1e-6, 1e-3, (), saveat="interp", norm=T.dual_norm, pow="fastpow", record=steps)
    return U, S, st, steps


def test_savannah_sensitivity_results_hard_pin(oracle_lib):
    """d Q(T = 200 s) / d ManningN zones on the river mesh, `ForwardDiff.jacobian` of the adaptive solve
    (sensitivity_analysis/ManningN/Savana_River/sensitivity_results.json): values and six partials carried by the oracle's
    dual-number JVP through the restated Dual-norm Tsit5 agree with the reference to 4e-9 of the largest entry (entries up to
    31.6; 202 accepted steps).  Pins the forward-mode derivative of a1-a7 -- friction, the inlet conveyance split through
    n, the zone gather -- that the brute-force J^T lambda of the VJP tests is assembled from."""
    U, S, st, _ = savannah_sensitivity_solve()
    K = S.shape[0]
    err = [np.abs(U[1 + k] - S[k]).max() for k in range(K)]
    print("savannah sensitivities vs reference:", ["%.1e" % e for e in err], "largest entry %.1f" % np.abs(S).max(), st)
    assert np.abs(S).max() > 10.0 and np.abs(S[0]).max() == 0.0        # zone 0 (the default material) owns no cell
    assert max(err) <= 2e-8 * np.abs(S).max()
