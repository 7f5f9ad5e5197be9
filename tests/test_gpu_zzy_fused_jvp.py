"""The fused forward-mode tile kernel (hg_fjvp.cu: hg_rhs_jvp / hg_rhs_jvp_multi on a fused, non-strict context) against the
oracle's dual-number pass -- ForwardDiff.Dual semantics through swe_2d_rhs (swe_2D_sensitivity.jl:34-80) on the performance
path: each face once on the staged tile, K directions per launch.  Gates: values <= 1e-12 of the flux scale (they are the
fused RHS's), tangents <= 1e-11 relative; also against the strict-path forward mode, the VJP kernel (transpose identity) and
across tile shapes."""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases

pytestmark = pytest.mark.gpu
ACTIVE = {None: 0, "zb": 1, "ManningN": 2, "Q": 3}


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _params(c, flat, mode):
    if mode == "ManningN":
        return np.asarray(c.ManningN_zone, dtype=np.float64).copy()
    if mode == "zb":
        return np.asarray(c.zb_cells, dtype=np.float64).copy()
    if mode == "Q":
        return np.asarray(flat["inletQ_TotalQ"], dtype=np.float64).copy()
    return None


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
@pytest.mark.parametrize("mode", [None, "ManningN", "zb", "Q"])
def test_fused_forward_mode_matches_the_oracle(hg, name, mode):
    c = cases.load(name)
    flat = R.flatten(c)
    if mode == "Q" and flat["n_inletq"] == 0:
        pytest.skip("no inlet-q boundary")
    N = c.mesh.numOfCells
    o = Oracle(flat)
    ctx = hg.Context(flat)
    rng = np.random.default_rng(31)
    p = _params(c, flat, mode)
    for seed in (0, 1, 2):
        Q = cases.random_state(c, seed) if seed else c.Q0
        v = rng.standard_normal(3 * N)
        pdot = rng.standard_normal(p.size) if p is not None else None
        dQ, jv = ctx.rhs_jvp(Q, v, p, mode, pdot)
        ref, ref_jv = o.jvp(Q, v, p, pdot, ACTIVE[mode])
        sc = cases.flat_scale(flat, Q)      # the denominator of the fused-path parity tests (tests/test_gpu_parity.py)
        assert (np.abs(dQ - ref) <= 1e-12 * sc).all(), (name, mode, seed)
        err = np.abs(jv - ref_jv).max() / np.abs(ref_jv).max()
        print(name, mode, seed, "fused forward mode: tangent rel. err %.1e" % err)
        assert err <= 1e-11, (name, mode, seed)
        assert (np.abs(dQ - ctx.rhs(Q, p, mode)) <= 1e-13 * sc).all()     # the values are the fused RHS kernel's, to rounding
        only = ctx.rhs_jvp(Q, v, p, mode, pdot, want_rhs=False)
        assert np.array_equal(only, jv)


@pytest.mark.parametrize("mode", [None, "ManningN", "zb", "Q"])
def test_fused_chunk_of_directions(hg, mode):
    """K directions in ONE launch (blockIdx.y): the same bits as K single calls; linear in the direction."""
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    ctx = hg.Context(flat)
    rng = np.random.default_rng(32)
    p = _params(c, flat, mode)
    Q = cases.random_state(c, 1)
    v = rng.standard_normal(3 * N)
    pdot = rng.standard_normal(p.size) if p is not None else None
    dQ, jv = ctx.rhs_jvp(Q, v, p, mode, pdot)
    V3 = np.stack([v, -2.0 * v, rng.standard_normal(3 * N)])
    P3 = np.stack([pdot, -2.0 * pdot, rng.standard_normal(p.size)]) if p is not None else None
    dQm, JV = ctx.rhs_jvp_multi(Q, V3, p, mode, P3)
    assert np.array_equal(dQm, dQ) and np.array_equal(JV[0], jv)
    assert np.abs(JV[1] + 2.0 * jv).max() <= 1e-12 * np.abs(jv).max()
    assert np.array_equal(JV[2], ctx.rhs_jvp(Q, V3[2], p, mode, None if P3 is None else P3[2], want_rhs=False))


def test_fused_forward_mode_on_synthetic_meshes_with_dry_cells(hg):
    """River (inlet-q / exit-h / walls, six zones) and thin-film dam break with wet/dry fronts: fused vs strict-path forward
    mode, the transpose identity against the VJP kernel, and every tile shape the same to rounding."""
    from hydrograd_jl_b200 import synthetic as S
    rng = np.random.default_rng(5)
    for flat, dry in ((S.river(64, 40)[0], 0.03), (S.dam_break(40, thin_film=True)[0], 0.05)):
        N = flat["n_cells"]
        Q = cases.random_state_flat(flat, 7, dry_frac=dry)
        p = np.linspace(0.02, 0.05, int(flat["n_mat"]))
        v, lam, pdot = rng.standard_normal(3 * N), rng.standard_normal(3 * N), rng.standard_normal(p.size)
        strict = hg.Context(flat, strict=True)
        _, jv_ref = strict.rhs_jvp(Q, v, p, "ManningN", pdot)
        base = None
        for tile in (256, 128, 192, 512):
            fused = hg.Context(flat, tile_cells=tile)
            _, jv = fused.rhs_jvp(Q, v, p, "ManningN", pdot)
            assert np.abs(jv - jv_ref).max() <= 1e-11 * np.abs(jv_ref).max(), tile
            if base is None:
                base = jv
                Qbar, pbar = fused.rhs_vjp(Q, lam, p, "ManningN")[:2]
                lhs, rhs = lam @ jv, Qbar @ v + pbar @ pdot
                assert abs(lhs - rhs) <= 1e-10 * np.abs(lam * jv).sum()
            else:
                assert np.abs(jv - base).max() <= 1e-13 * np.abs(base).max(), tile


def test_fused_forward_mode_million_cells(hg):
    """1M cells: bounded against the strict-path forward mode (which is pinned to the oracle and to the reference's
    sensitivities), and timed next to it."""
    import time
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(1000, 1000)
    N = flat["n_cells"]
    rng = np.random.default_rng(6)
    p = S.RIVER_N_ZONES[:flat["n_mat"]].copy()
    K = 6
    V = rng.standard_normal((K, 3 * N))
    Pd = np.eye(K)[:, :p.size] if p.size == K else rng.standard_normal((K, p.size))
    fused = hg.Context(flat)
    strict = hg.Context(flat, strict=True)
    t0 = time.perf_counter(); _, JV = fused.rhs_jvp_multi(Q0, V, p, "ManningN", Pd); t1 = time.perf_counter()
    _, JVs = strict.rhs_jvp_multi(Q0, V, p, "ManningN", Pd); t2 = time.perf_counter()
    err = np.abs(JV - JVs).max() / np.abs(JVs).max()
    print("1M cells, K = 6: fused %.3f s (incl. PCIe), strict %.3f s, rel. diff %.1e" % (t1 - t0, t2 - t1, err))
    assert err <= 1e-11


def test_fused_sensitivity_solve_reproduces_the_reference(hg):
    """hg_solve_tsit5_sens on a FUSED context: the reference's sensitivity driver (ForwardDiff.jacobian around solve(Tsit5),
    swe_2D_sensitivity.jl:34-80) with the six Manning-zone partials carried by the fused forward-mode kernel, one launch per
    Tsit5 stage -- against the committed sensitivity_results.json of the Savannah case and against the strict-path solve."""
    import os
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    z = np.load(os.path.join(cases.GOLD, "savannah_sens", "sensitivity.npz"))
    p = z["params_vector"]
    S_ref = z["sensitivity_results"].reshape(p.size, 3 * N)
    ctx = hg.Context(flat)
    ctx.set_controller_pow("fastpow")
    QT, S, st = ctx.solve_tsit5_sens(c.Q0, p, "ManningN", 0.0, 200.0, 0.02, True, 1e-6, 1e-3)
    err = [np.abs(S[k] - S_ref[k]).max() for k in range(p.size)]
    print("fused sensitivity solve vs reference:", ["%.1e" % e for e in err], st)
    assert abs(st["accepted"] - 202) <= 10 and st["rejected"] <= 6
    assert max(err) <= 1e-5 * np.abs(S_ref).max()
    strict = hg.Context(flat, strict=True)
    strict.set_controller_pow("fastpow")
    QTs, Ss, sts = strict.solve_tsit5_sens(c.Q0, p, "ManningN", 0.0, 200.0, 0.02, True, 1e-6, 1e-3)
    print("fused vs strict solve: states %.1e, sensitivities %.1e" % (np.abs(QT - QTs).max(), np.abs(S - Ss).max()), st, sts)
    assert np.abs(S - Ss).max() <= 1e-5 * np.abs(Ss).max()
    # dense-output saves and the fixed-step mode through the same code: identical to the strict path to rounding
    ts = np.array([0.0, 0.5, 1.0])
    _, Sf, stf = ctx.solve_tsit5_sens(c.Q0, p, "ManningN", 0.0, 1.0, 0.05, False, t_save=ts)
    _, Sp, stp = strict.solve_tsit5_sens(c.Q0, p, "ManningN", 0.0, 1.0, 0.05, False, t_save=ts)
    assert np.abs(stf["saves"] - stp["saves"]).max() <= 1e-10 and np.abs(Sf - Sp).max() <= 1e-9 * max(np.abs(Sp).max(), 1e-300)
