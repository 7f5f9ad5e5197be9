"""Boundary types the reference's fixtures and the synthetic benchmark meshes do not contain -- symmetry, several inlet-q node
strings, corner cells with two boundary types, the default-wall rule -- on random mixed tri / quad meshes written as SRH-2D files:
the whole product path (C++ reader -> hg_create -> fused / strict RHS, hand-written VJP) against the oracle fed by the
reference-shaped builder.  The two sides number the ghost cells differently on purpose, so the test also shows that results
do not depend on the ghost order.  (Written after the round's GPU budget was spent: not yet run on a B200.)"""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases
from tests.test_srh_reader_cpu import _write_random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_symmetry_and_two_inlets_on_random_meshes(hg, tmp_path, seed):
    from hydrograd_jl_b200 import srh2d
    _write_random_case(str(tmp_path), seed)
    c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    ref = R.flatten(c)
    flat = srh2d.process_SRH_2D_input(str(tmp_path), "rnd.srhhydro")
    Q0 = srh2d.setup_initial_condition(flat, 3.0, 2.0, 0.1, 0.0)
    assert flat["n_symm"] == 1 and flat["n_inletq"] == 2 and np.array_equal(Q0, c.Q0)
    N = flat["n_cells"]
    o = Oracle(ref)
    rng = np.random.default_rng(seed)
    pn = np.asarray(c.ManningN_zone, dtype=np.float64) * (1 + 0.1 * rng.uniform(-1, 1, c.ManningN_zone.size))
    pq = np.asarray(ref["inletQ_TotalQ"], dtype=np.float64) * 0.8
    pz = np.asarray(c.zb_cells, dtype=np.float64) + 0.02 * rng.standard_normal(N)
    states = (Q0, cases.random_state_flat(ref, seed + 5, dry_frac=0.08))
    for p, mode, code in ((None, None, 0), (pn, "ManningN", 2), (pq, "Q", 3), (pz, "zb", 1)):
        fused = hg.Context(flat, tile_cells=128)                 # one pair of contexts per parameter mode
        strict = hg.Context(flat, strict=True)
        for Q in states:
            sc = cases.flat_scale(ref, Q)
            want = o.rhs(Q, p, code)
            got = fused.rhs(Q, p, mode)
            assert (np.abs(got - want) <= 1e-12 * sc).all(), (seed, mode)
            assert (np.abs(strict.rhs(Q, p, mode) - want) <= 1e-13 * sc).all(), (seed, mode)
            lam = rng.standard_normal(3 * N)
            Qbar_ref, pbar_ref = o.vjp_bruteforce(Q, lam, p, code)
            Qbar, pbar = fused.rhs_vjp(Q, lam, p, mode)
            assert np.abs(Qbar - Qbar_ref).max() <= 1e-9 * np.abs(Qbar_ref).max(), (seed, mode)
            if mode:
                assert np.abs(pbar - pbar_ref).max() <= 1e-9 * max(np.abs(pbar_ref).max(), 1e-30), (seed, mode)
            v = rng.standard_normal(3 * N)                       # forward mode on the same boundaries
            _, jv = strict.rhs_jvp(Q, v, p, mode)
            assert np.abs(jv - o.jvp(Q, v, p, None, code)[1]).max() <= 1e-11 * np.abs(jv).max(), (seed, mode)
    # Euler stepping with the symmetry boundary in place: 100 steps against the oracle's stepper
    fused = hg.Context(flat, tile_cells=128)
    fused.set_state(Q0)
    fused.step_euler(2e-3, 100)
    assert np.abs(fused.get_state() - o.euler(Q0, 2e-3, 100)).max() <= 1e-9
