"""GPU parity of the hand-written VJP kernel against the dual-number oracle (ForwardDiff semantics):
exact J^T lambda assembled from 3N + nP forward-mode passes on the fixture meshes, and the adjoint identity
lambda . (J v) == Qbar . v + pbar . vp on larger synthetic meshes.  Gate (north_star): <= 1e-9 on gradients."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _state(flat, seed, dry_frac):
    Q = cases.random_state_flat(flat, seed, dry_frac=dry_frac)
    return Q


def _params(c, mode, rng):
    if mode == "ManningN":
        return 2, c.ManningN_zone * (1 + 0.2 * rng.uniform(-1, 1, c.ManningN_zone.size))
    if mode == "zb":
        return 1, c.zb_cells + 0.02 * rng.standard_normal(c.zb_cells.size)
    if mode == "Q":
        return 3, c.bc.inletQ_TotalQ * 0.8
    return 0, None


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
@pytest.mark.parametrize("mode", [None, "ManningN", "zb", "Q"])
def test_vjp_matches_bruteforce_oracle(hg, name, mode):
    c = cases.load(name)
    flat = R.flatten(c)
    rng = np.random.default_rng(17)
    code, p = _params(c, mode, rng)
    o = Oracle(flat)
    ctx = hg.Context(flat, tile_cells=128)
    for seed, dry in ((0, 0.0), (1, 0.08)):
        Q = _state(flat, seed, dry)
        lam = rng.standard_normal(Q.size)
        Qbar_ref, pbar_ref = o.vjp_bruteforce(Q, lam, p, code)
        Qbar, pbar = ctx.rhs_vjp(Q, lam, p, mode)
        assert np.abs(Qbar - Qbar_ref).max() <= TOL * np.abs(Qbar_ref).max(), (name, mode, seed)
        if mode:
            assert np.abs(pbar - pbar_ref).max() <= TOL * max(np.abs(pbar_ref).max(), 1e-30), (name, mode, seed)
        Qbar2, _ = ctx.rhs_vjp(Q, lam, p, mode)
        assert np.array_equal(Qbar, Qbar2), "VJP is not bit-reproducible"


@pytest.mark.parametrize("which", ["dam_thin", "river"])
@pytest.mark.parametrize("mode", [None, "ManningN", "zb"])
def test_vjp_adjoint_identity_synthetic(hg, which, mode):
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.dam_break(40, thin_film=True) if which == "dam_thin" else S.river(96, 40)
    if which == "dam_thin" and mode == "ManningN":
        flat = dict(flat)
    rng = np.random.default_rng(23)
    N = flat["n_cells"]
    if mode == "ManningN":
        code, p = 2, np.full(flat["n_mat"], 0.03) * (1 + 0.1 * rng.uniform(-1, 1, flat["n_mat"]))
    elif mode == "zb":
        code, p = 1, flat["zb_cells"] + 0.01 * rng.standard_normal(N)
    else:
        code, p = 0, None
    o = Oracle(flat)
    ctx = hg.Context(flat, tile_cells=256)
    for Q in (Q0, cases.random_state_flat(flat, 4, dry_frac=0.05)):
        lam, v = rng.standard_normal(3 * N), rng.standard_normal(3 * N)
        vp = rng.standard_normal(p.size) * 0.01 if p is not None else None
        _, jv = o.jvp(Q, v, p, vp, code)
        Qbar, pbar = ctx.rhs_vjp(Q, lam, p, mode)
        lhs = lam @ jv
        rhs = Qbar @ v + (pbar @ vp if p is not None else 0.0)
        scale = np.abs(lam * jv).sum()
        assert abs(lhs - rhs) <= 1e-11 * scale, (which, mode)


def test_ncell_bar_hook(hg):
    """ncell_bar = d(lambda . rhs)/d ManningN_cells; summed per zone it is pbar (process_ManningN_2D.jl:88 transposed)."""
    c = cases.load("savannah")
    flat = R.flatten(c)
    rng = np.random.default_rng(3)
    Q = cases.random_state_flat(flat, 9, dry_frac=0.02)
    lam = rng.standard_normal(Q.size)
    ctx = hg.Context(flat, tile_cells=128)
    Qbar, pbar, nbar = ctx.rhs_vjp(Q, lam, c.ManningN_zone, "ManningN", want_ncell_bar=True)
    z = np.array([nbar[c.matID == k].sum() for k in range(c.ManningN_zone.size)])
    assert np.abs(z - pbar).max() <= 1e-12 * np.abs(pbar).max()


def test_vjp_launch_shapes_agree(hg):
    """Every compiled launch shape of the VJP kernel (tile size x threads x one/two faces per trip) on a mesh with
    wet/dry fronts and all boundary types: same result to rounding (the shapes differ only in instruction order)."""
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.river(300, 120)
    Q = cases.random_state_flat(flat, 12, dry_frac=0.06)
    lam = np.random.default_rng(8).standard_normal(Q.size)
    p = np.full(flat["n_mat"], 0.035)
    base = None
    for tile, variant in ((256, 0), (256, 1), (256, 2), (384, 0), (384, 1), (224, 0), (224, 1), (192, 0), (192, 1), (192, 2), (128, 0), (512, 0)):
        Qbar, pbar = hg.Context(flat, tile_cells=tile, vjp_variant=variant).rhs_vjp(Q, lam, p, "ManningN")
        if base is None:
            base = (Qbar, pbar)
            continue
        assert np.abs(Qbar - base[0]).max() <= 1e-13 * np.abs(base[0]).max(), (tile, variant)
        assert np.abs(pbar - base[1]).max() <= 1e-12 * np.abs(base[1]).max(), (tile, variant)


@pytest.mark.parametrize("mode", ["ManningN", "zb", "Q"])
def test_host_buffer_vjp_pipeline_matches_resident(hg, mode):
    """>= 1M cells: hg_rhs_vjp streams Q and lambda in chunks over three streams (tiles run as their chunks land,
    Qbar chunks leave meanwhile; inlet coupling and parameter reductions after the last stage).  Same kernels on the
    same data as the device-resident sequence => identical bits, Qbar and pbar, for every parameter mode."""
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(1000, 1050)
    N = flat["n_cells"]
    assert N >= 1 << 20
    rng = np.random.default_rng(5)
    lam = rng.standard_normal(3 * N)
    p = {"ManningN": np.full(flat["n_mat"], 0.03) * (1 + 0.1 * rng.uniform(-1, 1, flat["n_mat"])),
         "zb": flat["zb_cells"] + 0.01 * rng.standard_normal(N), "Q": np.asarray(flat["inletQ_TotalQ"]) * 0.9}[mode]
    ctx = hg.Context(flat)
    Qbar, pbar, nbar = ctx.rhs_vjp(Q0, lam, p, mode, want_ncell_bar=True)
    ref = hg.Context(flat)
    ref.set_params(p, mode)
    ref.set_state(Q0)
    ref.set_lambda(lam)
    ref.vjp_resident()
    Qbar2, pbar2, nbar2 = ref.get_vjp(p.size, want_ncell_bar=True)
    assert np.array_equal(Qbar, Qbar2)
    assert np.array_equal(pbar, pbar2)
    assert np.array_equal(nbar, nbar2)
