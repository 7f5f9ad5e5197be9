"""Parameter ensembles (BASELINE config C5): M members in one launch must equal M separate runs, bit for bit."""
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


@pytest.mark.parametrize("mode", ["ManningN", "Q", None])
def test_ensemble_equals_individual_runs(hg, mode):
    c = cases.load("savannah")
    flat = R.flatten(c)
    rng = np.random.default_rng(11)
    M = 5
    ens = hg.Context(flat, tile_cells=128)
    ens.ensemble_alloc(M, per_member_manning=(mode == "ManningN"))
    one = hg.Context(flat, tile_cells=128)
    o = Oracle(flat)
    states, params = [], []
    for m in range(M):
        Q = cases.random_state_flat(flat, 30 + m, dry_frac=0.03)
        p = None
        if mode == "ManningN":
            p = c.ManningN_zone * (1 + 0.2 * rng.uniform(-1, 1, 6))      # n_zone (1 + 0.2 U(-1,1))
        elif mode == "Q":
            p = c.bc.inletQ_TotalQ * (1 + 0.2 * rng.uniform(-1, 1, 1))
        states.append(Q); params.append(p)
        ens.ensemble_set_member(m, Q, p, mode)
    ens.ensemble_rhs()
    code = {"ManningN": 2, "Q": 3, None: 0}[mode]
    for m in range(M):
        got = ens.ensemble_get_member(m, "rhs")
        ref = one.rhs(states[m], params[m], mode)
        assert np.array_equal(got, ref), (mode, m)
        orc = o.rhs(states[m], params[m], code)
        assert (np.abs(got - orc) / cases.flat_scale(flat, states[m])).max() <= 1e-12
    # a few fused Euler steps of all members at once (from the case's physical initial state) == the oracle's
    # stepper, member by member
    for m in range(M):
        ens.ensemble_set_member(m, c.Q0, params[m], mode)
    ens.ensemble_step_euler(0.01, 20)
    for m in (0, M - 1):
        got = ens.ensemble_get_member(m, "state")
        ref = o.euler(c.Q0, 0.01, 20, params[m], code)
        assert np.isfinite(ref).all()
        assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
