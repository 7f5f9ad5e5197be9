"""The pullback of a forward call (`y, back = Zygote.pullback(swe_2d_rhs, Q, p)`, debug_AD.jl:60,75; the shim's rrule):
hg_rhs uploads the primal's state once, hg_rhs_vjp with Q = NULL differentiates at that resident state and ships only
the cotangent.  Same kernels on the same data as the call that uploads Q again => identical bits; hg_state_generation
tells the caller when the state has moved."""
import numpy as np
import pytest

import _pkg
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def test_pullback_at_the_resident_state_small_mesh(hg):
    """Fixture-sized mesh (one chunk: upload / kernel / download in sequence)."""
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(96, 48)
    N = flat["n_cells"]
    rng = np.random.default_rng(11)
    Q1 = cases.random_state_flat(flat, 3, dry_frac=0.05)
    lam = rng.standard_normal(3 * N)
    p = np.full(flat["n_mat"], 0.03) * (1 + 0.1 * rng.uniform(-1, 1, flat["n_mat"]))
    ctx = hg.Context(flat)
    with pytest.raises(hg.HydrogradError) as e:      # nothing resident yet
        ctx.rhs_vjp(None, lam, p, "ManningN")
    assert "resident" in str(e.value)
    want_Qbar, want_pbar = ctx.rhs_vjp(Q1, lam, p, "ManningN")
    ctx.rhs(Q0, p, "ManningN")                        # another state in between
    dQ = ctx.rhs(Q1, p, "ManningN")
    g = ctx.state_generation()
    Qbar, pbar = ctx.rhs_vjp(None, lam, p, "ManningN")
    assert ctx.state_generation() == g                # a pullback at the resident state does not move it
    assert np.array_equal(Qbar, want_Qbar) and np.array_equal(pbar, want_pbar)
    # everything that moves the state changes the generation
    seen = {g}
    for move in (lambda: ctx.set_state(Q0), lambda: ctx.step_euler(1e-3, 2), lambda: ctx.step_rk4(1e-3, 1),
                 lambda: ctx.rhs(Q0, p, "ManningN"), lambda: ctx.rhs_vjp(Q1, lam, p, "ManningN")):
        move()
        assert ctx.state_generation() not in seen
        seen.add(ctx.state_generation())
    # the Python mirror of the rrule: forward value + closure; falls back to uploading Q once the state has moved
    extra = hg.SWE2D_Extra_Parameters(flat, active_param_name="ManningN")
    y, back = hg.swe_2d_rhs_pullback(Q1, p, 0.0, extra)
    assert np.array_equal(y, dQ)
    b1 = back(lam)
    extra.ctx.set_state(Q0)
    b2 = back(lam)
    for b in (b1, b2):
        assert np.array_equal(b[0], want_Qbar) and np.array_equal(b[1], want_pbar)


@pytest.mark.parametrize("mode", ["ManningN", "Q"])
def test_pullback_at_the_resident_state_pipelined(hg, mode):
    """>= 1M cells: both calls run the three-stream pipeline; the pullback's chunks carry lambda only."""
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(1000, 1050)
    N = flat["n_cells"]
    assert N >= 1 << 20
    rng = np.random.default_rng(7)
    lam = rng.standard_normal(3 * N)
    p = {"ManningN": np.full(flat["n_mat"], 0.03) * (1 + 0.1 * rng.uniform(-1, 1, flat["n_mat"])),
         "Q": np.asarray(flat["inletQ_TotalQ"]) * 0.9}[mode]
    ctx = hg.Context(flat)
    want = ctx.rhs_vjp(Q0, lam, p, mode, want_ncell_bar=True)
    ctx.rhs(Q0 * 0.5, p, mode)
    ctx.rhs(Q0, p, mode)
    got = ctx.rhs_vjp(None, lam, p, mode, want_ncell_bar=True)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
