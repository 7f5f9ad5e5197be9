"""Forward mode on the device (hg_rhs_jvp, hg_jvp.cu) against the oracle's dual-number pass -- the ForwardDiff.Dual semantics the
reference's sensitivity driver and its ForwardDiffSensitivity inversion option rely on.  The arithmetic is the same source as
tests/test_jvp_cpu.py checks on the host; this file covers the kernels' launch structure through the C ABI.  The fused
forward-mode tile kernel (hg_fjvp.cu, non-strict contexts) is covered by tests/test_gpu_zzy_fused_jvp.py."""
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
def test_device_forward_mode_matches_the_oracle(hg, name, mode):
    c = cases.load(name)
    flat = R.flatten(c)
    if mode == "Q" and flat["n_inletq"] == 0:
        pytest.skip("no inlet-q boundary")
    N = c.mesh.numOfCells
    o = Oracle(flat)
    ctx = hg.Context(flat, strict=True)
    rng = np.random.default_rng(21)
    p = _params(c, flat, mode)
    for seed in (0, 1):
        Q = cases.random_state(c, seed) if seed else c.Q0
        v = rng.standard_normal(3 * N)
        pdot = rng.standard_normal(p.size) if p is not None else None
        dQ, jv = ctx.rhs_jvp(Q, v, p, mode, pdot)
        ref, ref_jv = o.jvp(Q, v, p, pdot, ACTIVE[mode])
        sc = cases.flux_scale(c, Q)
        assert (np.abs(dQ - ref) <= 1e-12 * sc).all()
        assert np.abs(jv - ref_jv).max() <= 1e-11 * np.abs(ref_jv).max(), (name, mode, seed)
        strict = ctx.rhs(Q, p, mode)                       # the values are those of the strict path (same operations)
        assert (np.abs(dQ - strict) <= 1e-14 * sc).all()
        print(name, mode, seed, "values equal to the strict path bit for bit:", bool(np.array_equal(dQ, strict)),
              " tangent rel. err %.1e" % (np.abs(jv - ref_jv).max() / np.abs(ref_jv).max()))
        only = ctx.rhs_jvp(Q, v, p, mode, pdot, want_rhs=False)
        assert np.array_equal(only, jv)


@pytest.mark.parametrize("mode", [None, "ManningN", "zb", "Q"])
def test_chunk_of_directions_in_one_call(hg, mode):
    """hg_rhs_jvp_multi (a ForwardDiff chunk): the same sweeps as K single calls, the same bits."""
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    ctx = hg.Context(flat, strict=True)
    rng = np.random.default_rng(22)
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


def test_device_forward_mode_is_the_transpose_of_the_vjp_kernel(hg):
    """<lambda, J v> (forward mode, strict context) = <J^T lambda, v> (hand-written VJP kernel, fused context): the two
    derivative kernels of the library against each other on a synthetic mesh with dry cells."""
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(64, 40)
    N = flat["n_cells"]
    rng = np.random.default_rng(4)
    Q = cases.random_state_flat(flat, 7, dry_frac=0.03)
    p = np.linspace(0.02, 0.05, int(flat["n_mat"]))
    v, lam, pdot = rng.standard_normal(3 * N), rng.standard_normal(3 * N), rng.standard_normal(p.size)
    fwd = hg.Context(flat, strict=True)
    rev = hg.Context(flat)
    _, jv = fwd.rhs_jvp(Q, v, p, "ManningN", pdot)
    Qbar, pbar = rev.rhs_vjp(Q, lam, p, "ManningN")[:2]
    lhs, rhs = lam @ jv, Qbar @ v + pbar @ pdot
    assert abs(lhs - rhs) <= 1e-10 * np.abs(lam * jv).sum()


def test_forward_mode_error_behaviour(hg):
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    ctx = hg.Context(flat, strict=True)
    with pytest.raises(hg.HydrogradError):
        ctx.rhs_jvp(c.Q0, np.ones(3 * N), np.array([0.03]), "ManningN")            # wrong parameter length
    dry = c.Q0.copy()
    dry[:N] = -flat["hstill"] + 1e-4
    with pytest.raises(hg.HydrogradError) as e:
        ctx.rhs_jvp(dry, np.ones(3 * N))
    assert e.value.code == 3                                                          # inlet conveyance assert (bc_2D.jl:678-680)
    dQ, jv = ctx.rhs_jvp(c.Q0, np.zeros(3 * N))                                       # the context is usable afterwards
    assert not jv.any() and np.isfinite(dQ).all()


def test_device_forward_mode_reproduces_the_reference_sensitivities(hg):
    """The reference's Savannah sensitivity run -- ForwardDiff.jacobian of the adaptive Tsit5 solve, d Q(200 s) / d ManningN --
    with every Dual pass done by the device: values and six partials, hg_rhs_jvp per partial and stage, Dual-aware error norm
    and fastpow controller on the host (tests/tsit5_ref.py).  Compared with the reference's committed sensitivity_results.json
    (host build of the same source: 2.5e-9 of the largest entry, tests/test_jvp_cpu.py; tolerance here as for the other replays
    of tests/test_gpu_zzz_reference_replay.py)."""
    from tests.test_oracle_golden import savannah_sensitivity_solve
    c = cases.load("savannah")
    flat = R.flatten(c)
    ctx = hg.Context(flat, strict=True)

    def jvp(Q, V, p, pdot):
        return ctx.rhs_jvp(Q, V, p, "ManningN", pdot)

    U, S, st, _ = savannah_sensitivity_solve(jvp)
    err = [np.abs(U[1 + k] - S[k]).max() for k in range(S.shape[0])]
    print("savannah sensitivities, device forward mode vs reference:", ["%.1e" % e for e in err], st)
    assert max(err) <= 1e-5 * np.abs(S).max()


def test_device_resident_sensitivity_solve_reproduces_the_reference(hg):
    """hg_solve_tsit5_sens: the same run in ONE call -- augmented state [Q; dQ/dp_1..6] resident on the device, K forward-mode
    sweeps per Tsit5 stage, Dual-aware error norm reduced on the device, fastpow PI controller -- against the reference's
    committed sensitivity_results.json and against the host restatement's step count (202 accepted, 1 rejected)."""
    import os
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    z = np.load(os.path.join(cases.GOLD, "savannah_sens", "sensitivity.npz"))
    p = z["params_vector"]
    S_ref = z["sensitivity_results"].reshape(p.size, 3 * N)
    ctx = hg.Context(flat, strict=True)
    ctx.set_controller_pow("fastpow")
    QT, S, st = ctx.solve_tsit5_sens(c.Q0, p, "ManningN", 0.0, 200.0, 0.02, True, 1e-6, 1e-3)
    err = [np.abs(S[k] - S_ref[k]).max() for k in range(p.size)]
    print("device-resident sensitivity solve vs reference:", ["%.1e" % e for e in err], st)
    assert abs(st["accepted"] - 202) <= 10 and st["rejected"] <= 6
    assert max(err) <= 1e-5 * np.abs(S_ref).max()
    t = cases.truth("savannah")                                   # the values ride along: the forward run's final state
    assert np.abs(QT[:N] - t["xi_truth"]).max() <= 5e-3           # (that file comes from the plain forward run, other steps: 7.5e-4 on the host)
    h = ctx.last_steps()
    assert h.size == st["accepted"] and abs(h.sum() - 200.0) < 1e-9
    # fixed-step mode and the other parameters: finite, right shapes; Q: linear dependence -> S p = dQ/dlog-scale check by FD
    pQ = np.asarray(flat["inletQ_TotalQ"], dtype=np.float64)
    Q1, SQ, _ = ctx.solve_tsit5_sens(c.Q0, pQ, "Q", 0.0, 1.0, 0.05, False)
    Q2, _, _ = ctx.solve_tsit5_sens(c.Q0, pQ * (1 + 1e-6), "Q", 0.0, 1.0, 0.05, False)
    fd = (Q2 - Q1) / 1e-6
    assert np.abs(SQ.T @ pQ - fd).max() <= 1e-4 * max(np.abs(fd).max(), 1e-12)
