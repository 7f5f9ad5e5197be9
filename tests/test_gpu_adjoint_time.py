"""Discrete adjoint through the Euler stepper (hg_euler_adjoint) against forward sensitivities propagated with the
oracle's dual-number JVP: lambda_T . dQ_T/d(Q0,p)[v,w] == Q0bar . v + pbar . w  (gate: 1e-9)."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def oracle_forward_sensitivity(o, flat, Q0, p, code, v, w, dt, nsteps):
    """custom_ODE_update_cells (custom_ODE_solvers.jl:5-33) with its dry mask, state and tangent together."""
    N, hs = flat["n_cells"], flat["h_small"]
    Q, D = Q0.copy(), v.copy()
    for _ in range(nsteps):
        f, jv = o.jvp(Q, D, p, w, code)
        Q = Q + dt * f
        D = D + dt * jv
        m = Q[:N] < hs
        Q[:N][m] = hs; Q[N:2 * N][m] = 0.0; Q[2 * N:][m] = 0.0
        D[:N][m] = 0.0; D[N:2 * N][m] = 0.0; D[2 * N:][m] = 0.0
    return Q, D


@pytest.mark.parametrize("name,mode,nsteps", [("oneD_bump", "ManningN", 300), ("oneD_bump", "zb", 150), ("savannah", "ManningN", 120),
                                              ("savannah", "Q", 120), ("simple", None, 200)])
def test_euler_adjoint_matches_forward_sensitivities(hg, name, mode, nsteps):
    c = cases.load(name)
    flat = R.flatten(c)
    o = Oracle(flat)
    rng = np.random.default_rng(41)
    code = {"ManningN": 2, "zb": 1, "Q": 3, None: 0}[mode]
    p = {"ManningN": c.ManningN_zone.copy(), "zb": c.zb_cells.copy(), "Q": c.bc.inletQ_TotalQ.copy(), None: None}[mode]
    dt = 0.005
    N = flat["n_cells"]
    v = rng.standard_normal(3 * N) * 1e-2
    w = rng.standard_normal(p.size) * (1e-3 if mode != "Q" else 1.0) if p is not None else None
    lamT = rng.standard_normal(3 * N)
    QT_ref, DT = oracle_forward_sensitivity(o, flat, c.Q0, p, code, v, w, dt, nsteps)
    ctx = hg.Context(flat, tile_cells=128)
    QT, Q0bar, pbar = ctx.euler_adjoint(c.Q0, lamT, dt, nsteps, p, mode)
    assert np.abs(QT - QT_ref).max() <= 1e-9 * max(1.0, np.abs(QT_ref).max())
    lhs = lamT @ DT
    rhs = Q0bar @ v + (pbar @ w if p is not None else 0.0)
    assert abs(lhs - rhs) <= 1e-9 * np.abs(lamT * DT).sum(), (name, mode, lhs, rhs)


def test_euler_adjoint_with_dry_mask(hg):
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.dam_break(24, thin_film=True)
    o = Oracle(flat)
    rng = np.random.default_rng(5)
    N = flat["n_cells"]
    v = rng.standard_normal(3 * N) * 1e-3
    lamT = rng.standard_normal(3 * N)
    dt, nsteps = 0.01, 60
    QT_ref, DT = oracle_forward_sensitivity(o, flat, Q0, None, 0, v, None, dt, nsteps)
    assert (QT_ref[:N] == flat["h_small"]).any()          # the mask is exercised
    QT, Q0bar, _ = hg.Context(flat, tile_cells=128).euler_adjoint(Q0, lamT, dt, nsteps)
    assert np.abs(QT - QT_ref).max() <= 1e-9 * max(1.0, np.abs(QT_ref).max())
    lhs, rhs = lamT @ DT, Q0bar @ v
    assert abs(lhs - rhs) <= 1e-9 * np.abs(lamT * DT).sum()


def test_inversion_loss_gradient_savannah(hg):
    """Config C4: Manning's-n inversion step on the Savannah mesh -- loss (swe_2D_inversion.jl:388-467) and its gradient
    from the device forward + adjoint sweeps, against central finite differences of the oracle's Euler run."""
    from hydrograd_jl_b200 import inversion as inv
    c, t = cases.load("savannah"), cases.truth("savannah")
    flat = R.flatten(c)
    o = Oracle(flat)
    observed = dict(WSE_truth=t["wse_truth"], u_truth=t["u_truth"], v_truth=t["v_truth"], zb_cell_truth=t["zb_cell_truth"])
    p = np.full(6, 0.03)                                   # the inversion's initial guess
    dt, nsteps = 0.01, 150
    ctx = hg.Context(flat)
    loss, parts, grad = inv.loss_and_gradient(ctx, flat, c.Q0, p, "ManningN", observed, dt, nsteps, bound=(0.01, 0.06))

    def oracle_loss(pp):
        QT = o.euler(c.Q0, dt, nsteps, pp, 2)
        return inv.loss_terms(QT, pp, observed, flat, "ManningN", bound=(0.01, 0.06))[0]

    assert abs(loss - oracle_loss(p)) <= 1e-10 * loss
    fd = np.zeros(6)
    for k in range(6):
        e = np.zeros(6); e[k] = 1e-6
        fd[k] = (oracle_loss(p + e) - oracle_loss(p - e)) / 2e-6
    assert np.abs(grad - fd).max() <= 1e-5 * np.abs(fd).max(), (grad, fd)


def _rk_tables():
    from tests import tsit5_ref as T
    rk4 = (((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)), (1 / 6, 1 / 3, 1 / 3, 1 / 6))
    tsit = (T.A[:6], T.A[6])
    return {"RK4": rk4, "Tsit5": tsit}


def oracle_rk_forward_sensitivity(o, Q0, p, code, v, w, dt, nsteps, table):
    """State and tangent through nsteps fixed steps of an explicit RK method, the tangent by the dual-number JVP per stage."""
    A, b = table
    Q, D = Q0.copy(), v.copy()
    for _ in range(nsteps):
        k, dk = [], []
        for i in range(len(A)):
            Y, dY = Q.copy(), D.copy()
            for j, a in enumerate(A[i]):
                Y += dt * a * k[j]
                dY += dt * a * dk[j]
            f, jv = o.jvp(Y, dY, p, w, code)
            k.append(f); dk.append(jv)
        for i, bi in enumerate(b):
            Q = Q + dt * bi * k[i]
            D = D + dt * bi * dk[i]
    return Q, D


@pytest.mark.parametrize("method,name,mode,nsteps", [("RK4", "oneD_bump", "ManningN", 60), ("Tsit5", "oneD_bump", "zb", 40),
                                                     ("Tsit5", "savannah", "ManningN", 30), ("RK4", "savannah", "Q", 30),
                                                     ("Tsit5", "simple", None, 50)])
def test_rk_adjoint_matches_forward_sensitivities(hg, method, name, mode, nsteps):
    """hg_rk_adjoint (discrete adjoint of fixed-step RK4 / Tsit5) against forward sensitivities of the same steps propagated
    with the oracle's dual-number JVP:  lambda_T . dQ_T/d(Q0,p)[v,w] == Q0bar . v + pbar . w   (gate 1e-9)."""
    c = cases.load(name)
    flat = R.flatten(c)
    o = Oracle(flat)
    rng = np.random.default_rng(43)
    code = {"ManningN": 2, "zb": 1, "Q": 3, None: 0}[mode]
    p = {"ManningN": c.ManningN_zone * (1 + 0.1 * rng.uniform(-1, 1, c.ManningN_zone.size)), "zb": c.zb_cells.copy(),
         "Q": np.asarray(c.bc.inletQ_TotalQ, dtype=float) * 0.9, None: None}[mode]
    N = c.mesh.numOfCells
    dt = 0.01 if name != "savannah" else 0.02
    v = rng.standard_normal(3 * N) * 1e-2
    w = None if p is None else rng.standard_normal(p.size) * (1e-3 if mode != "Q" else 1.0)
    lam_T = rng.standard_normal(3 * N)
    QT_ref, DT = oracle_rk_forward_sensitivity(o, c.Q0, p, code, v, w, dt, nsteps, _rk_tables()[method])
    ctx = hg.Context(flat, tile_cells=128)
    QT, Q0bar, pbar = ctx.rk_adjoint(method, c.Q0, lam_T, dt, nsteps, p, mode)
    assert np.abs(QT - QT_ref).max() <= 1e-9 * max(1.0, np.abs(QT_ref).max())
    lhs = lam_T @ DT
    rhs = Q0bar @ v + (pbar @ w if p is not None else 0.0)
    scale = np.abs(lam_T * DT).sum()
    assert abs(lhs - rhs) <= 1e-9 * scale, (method, name, mode, lhs, rhs)
    # and the forward part is what the steppers themselves do
    ctx2 = hg.Context(flat, tile_cells=128)
    if p is not None:
        ctx2.set_params(p, mode)
    ctx2.set_state(c.Q0)
    if method == "RK4":
        ctx2.step_rk4(dt, nsteps)
    else:
        ctx2.solve_tsit5(0.0, dt * nsteps, dt, adaptive=False)
    assert np.abs(ctx2.get_state() - QT).max() <= 1e-10 * max(1.0, np.abs(QT).max())


def test_inversion_loss_gradient_savannah_tsit5(hg):
    """The same inversion step with the SciML default integrator (fixed-step Tsit5) instead of the customized Euler loop:
    loss and gradient from hg_solve_tsit5 + hg_rk_adjoint against central finite differences of the host Tsit5 on the oracle."""
    from hydrograd_jl_b200 import inversion as inv
    from tests import tsit5_ref as T
    c, t = cases.load("savannah"), cases.truth("savannah")
    flat = R.flatten(c)
    o = Oracle(flat)
    observed = dict(WSE_truth=t["wse_truth"], u_truth=t["u_truth"], v_truth=t["v_truth"], zb_cell_truth=t["zb_cell_truth"])
    p = np.full(6, 0.03)
    dt, nsteps = 0.05, 40
    ctx = hg.Context(flat)
    loss, parts, grad = inv.loss_and_gradient(ctx, flat, c.Q0, p, "ManningN", observed, dt, nsteps, method="Tsit5", bound=(0.01, 0.06))

    def oracle_loss(pp):
        QT, _, _ = T.solve(lambda u: o.rhs(u, pp, 2), c.Q0, 0.0, dt * nsteps, dt, adaptive=False)
        return inv.loss_terms(QT, pp, observed, flat, "ManningN", bound=(0.01, 0.06))[0]

    assert abs(loss - oracle_loss(p)) <= 1e-10 * loss
    fd = np.zeros(6)
    for k in range(6):
        e = np.zeros(6); e[k] = 1e-6
        fd[k] = (oracle_loss(p + e) - oracle_loss(p - e)) / 2e-6
    assert np.abs(grad - fd).max() <= 1e-5 * np.abs(fd).max(), (grad, fd)
