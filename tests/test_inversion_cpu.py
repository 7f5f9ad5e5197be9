"""Host logic of the inversion loss (hydrograd.jl_b200/inversion.py, mirror of swe_2D_inversion.jl:388-467, 738-748): the
terminal cotangent d loss / d Q(T) and the direct parameter derivative it hands to the device adjoint, against central finite
differences of the loss itself -- no device involved."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from tests import cases

_pkg.load()
from hydrograd_jl_b200 import inversion as inv      # noqa: E402


def _observed(c, t):
    return dict(WSE_truth=t["wse_truth"], u_truth=t["u_truth"], v_truth=t["v_truth"], zb_cell_truth=t["zb_cell_truth"])


@pytest.mark.parametrize("active", ["ManningN", "zb", "Q"])
@pytest.mark.parametrize("terms", [(True, True), (True, False), (False, True)])
def test_loss_cotangents_match_finite_differences(active, terms):
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    t = cases.truth("savannah")
    obs = _observed(c, t)
    rng = np.random.default_rng(2)
    h = t["h_truth"] * (1 + 0.05 * rng.standard_normal(N))
    Q = np.concatenate([h - flat["hstill"], 0.9 * t["u_truth"] * h, 1.1 * t["v_truth"] * h + 0.01])
    if active == "ManningN":
        p, bound = np.array([0.005, 0.04, 0.05, 0.03, 0.2, 0.05]), (np.full(6, 0.01), np.full(6, 0.1))      # two entries out of bounds
    elif active == "zb":
        p, bound = t["zb_cell_truth"] + 0.1 * rng.standard_normal(N), None
    else:
        p, bound = np.array([150.0]), (np.array([160.0]), np.array([300.0]))
    kw = dict(bWSE=terms[0], buv=terms[1], bound=bound)
    loss, parts, lam, dp = inv.loss_terms(Q, p, obs, flat, active, **kw)
    assert loss == pytest.approx(parts["WSE"] + parts["uv"] + parts["bound"]) and loss > 0
    assert lam.shape == (3 * N,) and dp.shape == p.shape

    def f(Qx, px):
        return inv.loss_terms(Qx, px, obs, flat, active, **kw)[0]

    for _ in range(3):                                              # directional derivatives in the state ...
        v = rng.standard_normal(3 * N)
        e = 1e-6
        fd = (f(Q + e * v, p) - f(Q - e * v, p)) / (2 * e)
        assert abs(fd - lam @ v) <= 1e-6 * max(abs(fd), np.abs(lam * v).sum() * 1e-3, 1e-12)
    w = rng.standard_normal(p.size)                                 # ... and in the parameters (direct dependence only)
    e = 1e-7 * max(1.0, np.abs(p).max())
    fd = (f(Q, p + e * w) - f(Q, p - e * w)) / (2 * e)
    assert abs(fd - dp @ w) <= 1e-5 * max(abs(fd), 1e-12) + 1e-12
    if active == "ManningN":
        assert parts["bound"] > 0 and dp[0] < 0 < dp[4] and not dp[1:4].any()


def test_bound_loss_is_zero_inside_the_bounds():
    lo, up = np.array([0.01, 0.01]), np.array([0.1, 0.1])
    val, g = inv.compute_bound_loss(np.array([0.05, 0.1]), lo, up)
    assert val == 0.0 and not g.any()
    val, g = inv.compute_bound_loss(np.array([0.0, 0.19]), lo, up)
    assert val == pytest.approx((0.01 / 0.09) ** 2 + (0.09 / 0.09) ** 2) and g[0] < 0 < g[1]
