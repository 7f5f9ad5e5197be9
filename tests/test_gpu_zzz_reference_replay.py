"""The CUDA path against the reference's OWN saved data.  The trajectories the reference committed for its two channel
sensitivity cases are values of a ForwardDiff.Dual solve; tests/test_oracle_golden.py::test_reference_trajectory_hard_pin
recovers the step sequence of that solve (Dual-aware error norm, DiffEqBase fastpow) by carrying values and partials with
the oracle.  The saved VALUES depend on the partials only through those step sizes, so the device can replay them: one fixed
Tsit5 step per recorded (t, h) on the resident state (hg_solve_tsit5_dense with adaptive = 0, saves inside a step by the
dense output).  The device RHS then has to reproduce the reference's saved states to a few 1e-9 over the first saves and to
1e-5 through the first 12.  Both channel runs are stability-limited (the controller sits at a constant error estimate), so
rounding-level differences in the RHS are amplified along the way: on the host, 1e-12 relative noise on the oracle RHS moves
the replayed saves by 3e-10 ... 3e-9 early and up to 3e-6 later, which is what the tolerances leave room for (the oracle
itself: 2e-11 ... 1e-9 early; CPU check of the replay logic: test_replaying_the_recovered_step_sequence_with_values_only).  (Written after the round's GPU budget was spent: not yet run
on a B200; every device entry point it uses is exercised by tests/test_gpu_tsit5.py.)"""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from tests import cases
from tests import tsit5_ref as T
from tests.test_oracle_golden import reference_step_sequence

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


@pytest.mark.parametrize("name,p,dt_save,early_tol", [("oneD_uniform_sens", [0.03, 0.03], 1.0, (3e-9, 3e-9, 2e-8)),
                                                      ("oneD_bump_sens", [0.03, 0.02, 0.03], 2.0, (1e-8, 5e-9))])
def test_device_replays_the_reference_run(hg, name, p, dt_save, early_tol):
    c = cases.load(name)
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
    idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
    n = int(np.searchsorted(idx, 12, side="right"))
    steps = reference_step_sequence(name, p, dt_save, int(idx[n - 1]))
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_params(np.array(p), "ManningN")
    ctx.set_state(c.Q0)

    def step(t0, t1, h, inside):
        saves, st = ctx.solve_tsit5(t0, t1, h, adaptive=False, t_save=inside, saveat="interp")
        assert st["accepted"] == 1 and st["rejected"] == 0
        return [] if saves is None else list(saves)

    got = T.replay(step, steps, dt_save * idx[:n])
    assert len(got) == n
    err = [max(np.abs(g[:N] - w[:N]).max(), np.abs(g[N:2 * N] - w[N:2 * N]).max()) for g, w in zip(got, ref)]
    print(name, "device replay vs the reference's saved trajectory:", ["%.1e" % e for e in err])
    for e, tol in zip(err, early_tol):
        assert e <= tol
    assert max(err) <= 1e-5
