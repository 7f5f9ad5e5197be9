"""Checks of an RHS / VJP implementation against the independent literal restatement (oracle/rhs_literal.py).  The bodies are
shared: tests/test_gpu_zzzz_literal.py runs them on the device contexts, tests/test_oracle_literal_cpu.py runs the very same
code on a stand-in backed by the C++ oracle, so the harness itself is exercised without a GPU."""
import numpy as np

from oracle import rhs_literal as LIT
from oracle import srh2d_ref as R
from tests import cases

MODES = (("", None), ("ManningN", "n"), ("zb", "z"), ("Q", "q"))


def params_for(c, kind, rng):
    if kind == "n":
        return np.asarray(c.ManningN_zone, dtype=np.float64) * (1 + 0.2 * rng.uniform(-1, 1, c.ManningN_zone.size))
    if kind == "z":
        return np.asarray(c.zb_cells, dtype=np.float64) + 0.02 * rng.standard_normal(c.zb_cells.size)
    if kind == "q":
        return np.asarray(c.bc.inletQ_TotalQ, dtype=np.float64) * 0.8
    return None


def check_rhs(make_ctx, name, tol):
    """ctx.rhs(Q, params, mode) against the literal restatement: relative to the flux scale, all parameter modes, the
    fixture's initial condition and two fuzzed states with dry cells."""
    c = cases.load(name)
    flat = R.flatten(c)
    ctx = make_ctx(flat)
    rng = np.random.default_rng(41)
    worst = 0.0
    for Q in (c.Q0, cases.random_state_flat(flat, 0), cases.random_state_flat(flat, 1, dry_frac=0.08)):   # (seeds of test_gpu_parity's fuzz states)
        sc = cases.flat_scale(flat, Q)
        for mode, kind in MODES:
            p = params_for(c, kind, rng)
            if kind == "q" and p.size == 0:
                continue
            want = LIT.swe_2d_rhs(c, Q, p, mode)
            got = ctx.rhs(Q, p, mode or None)
            err = float((np.abs(got - want) / sc).max())
            worst = max(worst, err)
            assert err <= tol, (name, mode, err)
    return worst


def check_vjp_identity(make_ctx, name, tol):
    """lambda . (J_Q v + J_p pdot) with the AD-free complex-step derivative of the literal restatement on the left, the
    implementation's (Qbar, pbar) = J^T lambda on the right."""
    c = cases.load(name)
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    rng = np.random.default_rng(43)
    e = 1e-30
    worst = 0.0
    Q = cases.random_state_flat(flat, 1, dry_frac=0.08)          # (the state of test_gpu_vjp.py)
    for mode, kind in MODES:
        p = params_for(c, kind, rng)
        if kind == "q" and p.size == 0:
            continue
        lam = rng.standard_normal(3 * N)
        ctx = make_ctx(flat)                                   # one context per parameter mode, like the other VJP tests
        Qbar, pbar = ctx.rhs_vjp(Q, lam, p, mode or None)[:2]
        for _ in range(2):
            v = rng.standard_normal(3 * N)
            pdot = None if p is None else rng.standard_normal(p.size) * (1.0 if kind == "q" else 0.01)
            jv = np.imag(LIT.swe_2d_rhs(c, Q + 1j * e * v, None if p is None else p + 1j * e * pdot, mode)) / e
            lhs = float(lam @ jv)
            rhs = float(Qbar @ v) + (float(np.asarray(pbar) @ pdot) if p is not None else 0.0)
            err = abs(lhs - rhs) / float(np.abs(lam * jv).sum())
            worst = max(worst, err)
            assert err <= tol, (name, mode, err)
    return worst
