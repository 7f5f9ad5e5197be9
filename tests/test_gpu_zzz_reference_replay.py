"""The CUDA path against the reference's OWN saved data.  The trajectories the reference committed for its two channel
sensitivity cases are values of a ForwardDiff.Dual solve; tests/test_oracle_golden.py::test_reference_trajectory_hard_pin
recovers the step sequence of that solve (Dual-aware error norm, DiffEqBase fastpow) by carrying values and partials with
the oracle.  The saved VALUES depend on the partials only through those step sizes, so the device can replay them: one fixed
Tsit5 step per recorded (t, h) on the resident state (hg_solve_tsit5_dense with adaptive = 0, saves inside a step by the
dense output); likewise the 200 s Savannah River forward runs, and the reverse sweep over the Savannah sensitivity run.

What these tests can show is bounded by the runs themselves: all of them are stability-limited (the controller holds the
error estimate at a constant level, the explicit scheme sits at its stability boundary), and in that regime cell-to-cell
rounding differences are amplified along the trajectory.  The oracle reproduces the reference's files to 1e-11 ... 2e-9 because
it evaluates the RHS in the reference's own operation order; the fused kernel differs from it by rounding (<= 1e-12 of the flux
scale per call, typically a few ulp), and on the host, noise of 1e-14 / 1e-13 of the flux scale per RHS call moves the replayed
Savannah final state by 5e-8 / 5e-7 (xi) and the channel saves by 1e-8 ... 2e-7 / 1e-7 ... 2e-6 at the first two saves.  The
tolerances below are set from those figures.  (Written after the round's GPU budget was spent: not yet run on a B200; every
device entry point used here is exercised by tests/test_gpu_tsit5.py and tests/test_gpu_adjoint_time.py.)"""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from tests import cases
from tests import tsit5_ref as T
from tests.test_oracle_golden import reference_step_sequence

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


@pytest.mark.parametrize("variable_n", [False, True])
def test_device_replays_the_savannah_forward_run(hg, variable_n):
    """The reference's 200 s forward simulations on the Savannah River mesh (constant n, and Cheng's n(h, |U|, ks) evaluated
    inside every RHS): the step sequence recovered on the host (tests/test_oracle_golden.py::
    test_savannah_forward_run_reproduces_the_reference_final_state) replayed by the device, 202 fixed Tsit5 steps on the
    resident state; the final state is compared with the reference's committed truth file (oracle: 1.8e-9 / 1.0e-9; see the
    module docstring for what rounding differences of the kernel do to a replay)."""
    from tests.test_oracle_golden import _savannah_ks_cells, savannah_forward_steps
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    t = cases.truth("savannah_ks" if variable_n else "savannah")
    steps, u_host, _ = savannah_forward_steps(variable_n)
    ctx = hg.Context(flat, tile_cells=128)
    if variable_n:
        ctx.set_manning_function("h_Umag_ks", ks_cells=_savannah_ks_cells())
    ctx.set_state(c.Q0)
    for t0, h in steps:
        _, st = ctx.solve_tsit5(t0, t0 + h, h, adaptive=False)
        assert st["accepted"] == 1
    u = ctx.get_state()
    den = u[:N] + flat["hstill"] + flat["h_small"]              # the reference saves u = q / (h + h_small)
    err = (np.abs(u[:N] - t["xi_truth"]).max(), np.abs(u[N:2 * N] / den - t["u_truth"]).max(), np.abs(u[2 * N:] / den - t["v_truth"]).max())
    print("device replay of the savannah forward run vs truth (xi, u, v):", ["%.1e" % e for e in err])
    assert max(err) <= 1e-5
    assert np.abs(u - u_host).max() <= 5e-5


@pytest.mark.parametrize("variable_n", [False, True])
def test_device_adaptive_solve_with_fastpow_lands_on_the_reference_final_state(hg, variable_n):
    """No replay: the device's own adaptive Tsit5 (error norm reduced on the device, PI controller on the host) with
    hg_set_controller_pow(ctx, 1) follows OrdinaryDiffEq's step sequence by itself and lands on the reference's committed
    Savannah final states (host figures: 1e-9 with fastpow, 8e-8 / 1e-7 with the exact power; on the device both are bounded by
    the amplified rounding differences of the kernel, see the module docstring)."""
    from tests.test_oracle_golden import _savannah_ks_cells, savannah_forward_steps
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    t = cases.truth("savannah_ks" if variable_n else "savannah")
    _, _, st_host = savannah_forward_steps(variable_n)
    err = {}
    for mode in ("fastpow", "exact"):
        ctx = hg.Context(flat, tile_cells=128)
        if variable_n:
            ctx.set_manning_function("h_Umag_ks", ks_cells=_savannah_ks_cells())
        ctx.set_controller_pow(mode)
        ctx.set_state(c.Q0)
        _, st = ctx.solve_tsit5(0.0, 200.0, 0.02, True, 1e-6, 1e-3)
        u = ctx.get_state()
        den = u[:N] + flat["hstill"] + flat["h_small"]
        err[mode] = max(np.abs(u[:N] - t["xi_truth"]).max(), np.abs(u[N:2 * N] / den - t["u_truth"]).max(),
                        np.abs(u[2 * N:] / den - t["v_truth"]).max())
        if mode == "fastpow":       # the host run of the same algorithm: 202 accepted, 1 rejected (a borderline accept may flip)
            assert abs(st["accepted"] - st_host["accepted"]) <= 10 and st["rejected"] <= st_host["rejected"] + 5, (st, st_host)
    print("device adaptive solve vs the reference's final state:", {k: "%.1e" % v for k, v in err.items()})
    assert err["fastpow"] <= 1e-5 and err["exact"] <= 1e-5


def test_device_adjoint_through_the_reference_run_matches_its_sensitivities(hg):
    """The device's REVERSE mode against the reference's published sensitivities.  sensitivity_results.json of the Savannah
    case is d Q(200 s) / d ManningN (ForwardDiff through the adaptive solve, step sizes plain Float64: constants of the
    differentiation).  The discrete adjoint of the same step sequence (hg_rk_adjoint_steps: checkpointed reverse sweep, one
    hand-written VJP launch per Tsit5 stage, 202 steps) must return pbar = S lambda for any cotangent lambda of the final state."""
    from tests.test_oracle_golden import savannah_sensitivity_solve
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    U, S, st, steps = savannah_sensitivity_solve()
    hs = np.array([h for _, h in steps])
    assert hs.size == st["accepted"] and abs(hs.sum() - 200.0) < 1e-9
    rng = np.random.default_rng(3)
    lam = rng.standard_normal(3 * N)
    ctx = hg.Context(flat, tile_cells=128)
    QT, Q0bar, pbar = ctx.rk_adjoint_steps("Tsit5", c.Q0, lam, hs, c.ManningN_zone, "ManningN")
    assert np.abs(QT - U[0]).max() <= 5e-5                                # the forward sweep lands on the host's final state
    want = S @ lam
    scale = np.abs(S * lam[None, :]).sum(1)
    print("device adjoint vs reference sensitivities:", ["%.1e" % (abs(a - b) / max(sc, 1e-300)) for a, b, sc in zip(pbar, want, scale)])
    assert pbar[0] == 0.0 and want[0] == 0.0                              # zone 0 (default material) owns no cell
    for k in range(1, S.shape[0]):
        assert abs(pbar[k] - want[k]) <= 1e-4 * scale[k], k
    # and the adaptive device solve hands out its own accepted steps for the same purpose
    ctx.set_controller_pow("fastpow")
    ctx.set_params(c.ManningN_zone, "ManningN")
    ctx.set_state(c.Q0)
    _, st_dev = ctx.solve_tsit5(0.0, 200.0, 0.02, True, 1e-6, 1e-3)
    h_dev = ctx.last_steps()
    assert h_dev.size == st_dev["accepted"] and abs(h_dev.sum() - 200.0) < 1e-9 and h_dev[0] == 0.02


def test_strict_path_replays_the_channel_runs_through_the_host_integrator(hg):
    """The same replay with the STRICT device path (hg_plain.cu: one thread per cell in the reference's operation order, no
    FMA contraction; it has no device-resident integrator, so the restated Tsit5 steps on the host and calls hg_rhs per stage).
    Its rounding differences against the oracle are one to two orders smaller than the fused kernel's (<= 2e-14 of the flux
    scale), and the replay should sit correspondingly closer to the reference's saved states; the assertion only requires
    what the fused path is required to reach, the measured figures are printed."""
    for name, p, dt_save, tols in (("oneD_uniform_sens", [0.03, 0.03], 1.0, (2e-6, 2e-5)), ("oneD_bump_sens", [0.03, 0.02, 0.03], 2.0, (1e-6, 1e-5))):
        c = cases.load(name)
        flat = R.flatten(c)
        N = c.mesh.numOfCells
        tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
        idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
        steps = reference_step_sequence(name, p, dt_save, int(idx[1]))
        ctx = hg.Context(flat, strict=True)
        pv = np.array(p)
        state = {"u": c.Q0.copy()}

        def step(t0, t1, h, inside):
            u, saves, st = T.solve(lambda v: ctx.rhs(v, pv, "ManningN"), state["u"], t0, t1, h, adaptive=False, t_save=inside, saveat="interp")
            state["u"] = u
            return saves

        got = T.replay(step, steps, dt_save * idx[:2])
        err = [max(np.abs(g[:N] - w[:N]).max(), np.abs(g[N:2 * N] - w[N:2 * N]).max()) for g, w in zip(got, ref)]
        print(name, "strict-path replay vs the reference's saved trajectory:", ["%.1e" % e for e in err])
        for e, tol in zip(err, tols):
            assert e <= tol


def test_inversion_gradient_of_the_adaptive_run_against_the_reference_sensitivities(hg):
    """One iteration's work of the reference's Savannah ManningN inversion in its own configuration (adaptive Tsit5 over 200 s):
    inversion.loss_and_gradient(method="Tsit5_adaptive") -- adaptive device solve, discrete adjoint over its accepted steps --
    must return d loss / d ManningN = S lambda + direct terms, with S the reference's committed sensitivities of the same run
    and lambda the loss cotangent at the final state."""
    from hydrograd_jl_b200 import inversion as inv
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    z = np.load(cases.GOLD + "/savannah_sens/sensitivity.npz")
    p = z["params_vector"]
    S = z["sensitivity_results"].reshape(p.size, 3 * N)
    t = cases.truth("savannah")
    rng = np.random.default_rng(8)
    obs = dict(WSE_truth=t["wse_truth"] + 0.05 * rng.standard_normal(N), u_truth=t["u_truth"] * 1.1, v_truth=t["v_truth"] * 0.9,
               zb_cell_truth=t["zb_cell_truth"])
    bound = (np.full(p.size, 0.025), np.full(p.size, 0.06))           # the first zone value (0.02) is out of bounds
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_controller_pow("fastpow")
    loss, parts, grad = inv.loss_and_gradient(ctx, flat, c.Q0, p, "ManningN", obs, 0.02, 10000, method="Tsit5_adaptive", bound=bound)
    QT = ctx.get_state()
    _, _, lam, dp = inv.loss_terms(QT, p, obs, flat, "ManningN", bound=bound)
    want = S @ lam + dp
    scale = np.abs(S * lam[None, :]).sum(1) + np.abs(dp)
    print("inversion gradient vs reference sensitivities:", ["%.1e" % (abs(a - b) / max(s_, 1e-300)) for a, b, s_ in zip(grad, want, scale)],
          "loss %.3e" % loss, parts)
    assert loss > 0 and parts["bound"] > 0 and dp[0] != 0.0
    for k in range(p.size):
        assert abs(grad[k] - want[k]) <= 1e-4 * max(scale[k], 1e-300), k


# last: the stability-limited channel runs amplify rounding differences the most (module docstring)
@pytest.mark.parametrize("name,p,dt_save,early_tol", [("oneD_uniform_sens", [0.03, 0.03], 1.0, (2e-6, 2e-5)),
                                                      ("oneD_bump_sens", [0.03, 0.02, 0.03], 2.0, (1e-6, 1e-5))])
def test_device_replays_the_reference_run(hg, name, p, dt_save, early_tol):
    c = cases.load(name)
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
    idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
    n = 2                                                     # the first two saves (t = 1, 2 s / 2, 4 s): later ones are chaotic
    steps = reference_step_sequence(name, p, dt_save, int(idx[n - 1]))
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_params(np.array(p), "ManningN")
    ctx.set_state(c.Q0)

    def step(t0, t1, h, inside):
        saves, st = ctx.solve_tsit5(t0, t1, h, adaptive=False, t_save=inside, saveat="interp")
        assert st["accepted"] == 1 and st["rejected"] == 0
        return [] if saves is None else list(saves)

    got = T.replay(step, steps, dt_save * idx[:n])
    assert len(got) == n
    err = [max(np.abs(g[:N] - w[:N]).max(), np.abs(g[N:2 * N] - w[N:2 * N]).max()) for g, w in zip(got, ref)]
    print(name, "device replay vs the reference's saved trajectory:", ["%.1e" % e for e in err])
    for e, tol in zip(err, early_tol):
        assert e <= tol
