"""The remaining explicit fixed-step solvers of the control files on the device -- OrdinaryDiffEq's `Euler()` (no dry mask)
and `AB3()` (Ralston start-up, then the three-step Adams-Bashforth formula) -- against the same formulas driven by the oracle
RHS on the host, and the UDE training loss with its gradient through fixed-step Tsit5."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle import ude_ref as U
from oracle.oracle import Oracle
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def ab3_ref(f, u, dt, nsteps):
    """AB3ConstantCache perform_step! of OrdinaryDiffEq (third party, restated): steps 1-2 Ralston, then AB3."""
    k2 = k3 = None
    for step in range(1, nsteps + 1):
        k1 = f(u)
        if step <= 2:
            u = u + dt / 4.0 * (k1 + 3.0 * f(u + 2.0 / 3.0 * dt * k1))
            if step == 1:
                k3 = k1
            else:
                k2 = k1
        else:
            u = u + dt / 12.0 * (23.0 * k1 - 16.0 * k2 + 5.0 * k3)
            k2, k3 = k1, k2
    return u


def _close(a, b, N, tol):
    return np.abs(a[:N] - b[:N]).max() <= tol and np.abs(a[N:] - b[N:]).max() <= tol * max(1.0, np.abs(b[N:]).max())


def test_ab3_and_plain_euler_match_host_formulas(hg):
    c = cases.load("oneD_bump_sens")
    flat = R.flatten(c)
    o = Oracle(flat)
    p = np.array([0.03, 0.02, 0.03])
    f = lambda u: o.rhs(u, p, 2)
    N, dt, n = c.mesh.numOfCells, 0.01, 40
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_params(p, "ManningN")
    # Euler(): u+ = u + dt f(u)
    u = c.Q0.copy()
    for _ in range(n):
        u = u + dt * f(u)
    ctx.set_state(c.Q0)
    ctx.step_ode_euler(dt, n)
    assert _close(ctx.get_state(), u, N, 1e-10)
    # AB3(), in one call and split over three calls (the history stays on the device)
    ref = ab3_ref(f, c.Q0.copy(), dt, n)
    ctx.set_state(c.Q0)
    ctx.step_ab3(dt, n)
    one = ctx.get_state()
    assert _close(one, ref, N, 1e-10)
    ctx.set_state(c.Q0)
    ctx.step_ab3(dt, 1)
    ctx.step_ab3(dt, 6)
    ctx.step_ab3(dt, n - 7)
    assert np.array_equal(ctx.get_state(), one)
    # a restart in the middle is a different (second-order start-up) sequence
    ctx.set_state(c.Q0)
    ctx.step_ab3(dt, 10)
    ctx.step_ab3(dt, n - 10, restart=True)
    mid = ab3_ref(f, ab3_ref(f, c.Q0.copy(), dt, 10), dt, n - 10)
    assert _close(ctx.get_state(), mid, N, 1e-10)
    # another stepper in between resets the history by itself
    ctx.set_state(c.Q0)
    ctx.step_ab3(dt, 10)
    ctx.step_rk4(dt, 0)
    ctx.step_ab3(dt, n - 10)
    assert _close(ctx.get_state(), mid, N, 1e-10)


def test_ude_training_loss_and_gradient_through_tsit5(hg):
    """compute_loss_UDE (swe_2D_UDE.jl:522-599: WSE + velocity mismatch at the final time) and d loss / d theta through the
    fixed-step Tsit5 solve, against central differences of the same loss over the oracle-driven solve."""
    from hydrograd_jl_b200 import inversion as inv
    from hydrograd_jl_b200 import ude as hude
    from tests import tsit5_ref as T
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    cfg = dict(input_dim=1, output_dim=1, hidden_layers=[3, 3], activations=["tanh", "tanh"], h_bounds=[0.1, 0.5], output_bounds=[0.03, 0.06])
    pm = hude.UDEModel("ManningN_h", cfg, layernorm="whole")
    om = U.Model("ManningN_h", [3, 3], ["tanh", "tanh"], "whole", cfg["h_bounds"], cfg["output_bounds"])
    rng = np.random.default_rng(8)
    th = om.init_theta(rng)
    ur = U.UdeRhs(flat, om)
    dt, nsteps = 0.02, 15
    h = c.Q0[:N] + flat["hstill"]
    observed = dict(WSE_truth=h + flat["zb_cells"] + 0.01 * rng.standard_normal(N), u_truth=0.3 + 0.05 * rng.standard_normal(N),
                    v_truth=0.01 * rng.standard_normal(N), zb_cell_truth=np.asarray(flat["zb_cells"]))

    def loss_ref(theta):
        QT, _, _ = T.solve(lambda u: ur.rhs(u, theta), c.Q0, 0.0, dt * nsteps, dt, adaptive=False)
        return inv.loss_terms(QT, theta, observed, flat, "UDE")[0]

    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_ude_model(pm)
    loss, parts, grad = inv.compute_loss_UDE(ctx, flat, c.Q0, th, observed, dt, nsteps, method="Tsit5")
    assert abs(loss - loss_ref(th)) <= 1e-9 * abs(loss) and parts["WSE"] > 0 and parts["uv"] > 0
    w = rng.standard_normal(om.n_params)
    e = 1e-6
    fd = (loss_ref(th + e * w) - loss_ref(th - e * w)) / (2 * e)
    assert abs(fd - grad @ w) <= 1e-5 * np.abs(grad * w).sum()
