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
