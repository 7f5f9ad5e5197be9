"""Device-resident Tsit5 (hg_solve_tsit5) -- the integrator the reference's forward / sensitivity drivers use by default
(swe_2D_forward_simulation.jl:38-41) -- against the same algorithm driven by the oracle RHS on the host (tests/tsit5_ref.py),
and against the reference's own saved trajectory during the transient."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases
from tests import tsit5_ref as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _scaled(a, b, N):
    return max(np.abs(a[:N] - b[:N]).max(), np.abs(a[N:] - b[N:]).max() / max(1.0, np.abs(b[N:]).max()))


def test_fixed_step_tsit5_matches_host_tableau(hg):
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    o = Oracle(flat)
    ts = [0.0, 0.5, 1.0]
    ref_end, ref_saves, st = T.solve(lambda u: o.rhs(u), c.Q0, 0.0, 1.0, 0.01, adaptive=False, t_save=ts)
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_state(c.Q0)
    saves, stats = ctx.solve_tsit5(0.0, 1.0, 0.01, adaptive=False, t_save=ts)
    assert stats["accepted"] == st["accepted"] == 100 and stats["rejected"] == 0 and stats["rhs"] == st["rhs"]
    N = c.mesh.numOfCells
    assert np.array_equal(saves[0], c.Q0)
    for got, want in zip(saves, ref_saves):
        assert _scaled(got, want, N) <= 1e-10
    assert _scaled(ctx.get_state(), ref_end, N) <= 1e-10


@pytest.mark.parametrize("name,mode", [("oneD_bump_sens", "ManningN"), ("savannah", None)])
def test_adaptive_tsit5_matches_host_controller(hg, name, mode):
    """Same tableau, error norm and PI controller on both sides: the step sequences coincide (same accepted / rejected / RHS
    counts).  An adaptive solve amplifies rounding: perturbing the ORACLE RHS by 1e-13 relative moves the host solution of
    this very case by 3e-7 (measured; the fixed-step solve moves by 1e-14), so the states are compared to 2e-6 -- three
    orders below the integration tolerance; the tight check of the stage arithmetic is the fixed-step test above."""
    c = cases.load(name)
    flat = R.flatten(c)
    o = Oracle(flat)
    if mode == "ManningN":
        p, code = np.array([0.03, 0.02, 0.03]), 2
    else:
        p, code = None, 0
    t1 = 12.0 if name == "oneD_bump_sens" else 3.0
    ts = [t1 / 3, 2 * t1 / 3, t1]
    _, ref_saves, st = T.solve(lambda u: o.rhs(u, p, code), c.Q0, 0.0, t1, 0.02, True, 1e-6, 1e-3, ts)
    ctx = hg.Context(flat, tile_cells=128)
    if p is not None:
        ctx.set_params(p, mode)
    ctx.set_state(c.Q0)
    saves, stats = ctx.solve_tsit5(0.0, t1, 0.02, True, 1e-6, 1e-3, ts)
    assert stats == st, (stats, st)
    N = c.mesh.numOfCells
    for got, want in zip(saves, ref_saves):
        assert _scaled(got, want, N) <= 2e-6


def test_adaptive_tsit5_follows_the_reference_transient(hg):
    """The reference's saved trajectory of the sensitivity case (Tsit5, abstol 1e-6, reltol 1e-3, dt0 0.02) at t = 2 ... 24 s,
    where xi is still moving by ~0.1 m: the device solve with the reference's tolerances stays within 1e-3 of it, and a
    tight device solve within the reference's own integration error (4e-4)."""
    c = cases.load("oneD_bump_sens")
    flat = R.flatten(c)
    tj = np.load(cases.GOLD + "/oneD_bump_sens/trajectory.npz")
    idx = tj["early_index"][:6]
    ref = tj["forward_simulation_results_early"][:6]
    ts = list(2.0 * idx)
    N = 200
    for (abstol, reltol), lim in (((1e-6, 1e-3), 1e-3), ((1e-9, 1e-7), 4e-4)):
        ctx = hg.Context(flat, tile_cells=128)
        ctx.set_params(np.array([0.03, 0.02, 0.03]), "ManningN")
        ctx.set_state(c.Q0)
        saves, stats = ctx.solve_tsit5(0.0, ts[-1], 0.02, True, abstol, reltol, ts)
        assert stats["accepted"] > 50
        for got, want in zip(saves, ref):
            assert np.abs(got[:N] - want[:N]).max() < lim and np.abs(got[N:2 * N] - want[N:2 * N]).max() < lim


@pytest.mark.parametrize("adaptive", [False, True])
def test_dense_output_saveat_matches_host(hg, adaptive):
    """hg_solve_tsit5_dense = OrdinaryDiffEq's saveat: save times that are not step ends are interpolated with Tsit5's
    dense output on the device, the step sequence is the one of a solve without save times.  Same algorithm on the host
    (tsit5_ref.solve(saveat="interp")) with the oracle RHS."""
    c = cases.load("oneD_bump_sens")
    flat = R.flatten(c)
    o = Oracle(flat)
    p = np.array([0.03, 0.02, 0.03])
    t1 = 6.0 if adaptive else 1.0
    ts = [0.0, 0.137 * t1, t1 / 2, 0.9 * t1, t1]
    args = (0.0, t1, 0.02 if adaptive else 0.01, adaptive, 1e-6, 1e-3)
    ref_end, ref_saves, st = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, *args, t_save=ts, saveat="interp")
    _, _, st_free = T.solve(lambda u: o.rhs(u, p, 2), c.Q0, *args, t_save=[])
    assert st == st_free
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_params(p, "ManningN")
    ctx.set_state(c.Q0)
    saves, stats = ctx.solve_tsit5(*args, t_save=ts, saveat="interp")
    assert stats == st, (stats, st)
    N = c.mesh.numOfCells
    assert np.array_equal(saves[0], c.Q0)
    tol = 2e-6 if adaptive else 1e-10                # see test_adaptive_tsit5_matches_host_controller for the adaptive bound
    for got, want in zip(saves, ref_saves):
        assert _scaled(got, want, N) <= tol
    assert _scaled(ctx.get_state(), ref_end, N) <= tol
    assert np.array_equal(saves[-1], ctx.get_state())            # a save at the step end is a copy, not an interpolation
