"""Cross-check of the primary oracle (oracle/swe_oracle.cpp, scalar-expanded C++) against a second restatement written
independently from the Julia source in its own shapes (oracle/rhs_literal.py: 3x3 matrices for the Roe dissipation, per-boundary
vectors + update_1d_array for the ghosts, the comprehension of compute_inviscid_fluxes).  Two routes from the same source that
agree to rounding on fuzzed states pin what the reference's committed trajectories do not reach: every wet/dry branch of the
Roe flux, symmetry boundaries, several inlets, and the zb / ManningN / Q parameter bindings."""
import numpy as np
import pytest

from oracle import rhs_literal as LIT
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases
from tests.test_srh_reader_cpu import _write_random_case

MODES = ((None, "", 0), ("n", "ManningN", 2), ("z", "zb", 1), ("q", "Q", 3))


def _params(c, kind, rng):
    if kind == "n":
        return np.asarray(c.ManningN_zone, dtype=np.float64) * (1 + 0.2 * rng.uniform(-1, 1, c.ManningN_zone.size))
    if kind == "z":
        return np.asarray(c.zb_cells, dtype=np.float64) + 0.02 * rng.standard_normal(c.zb_cells.size)
    if kind == "q":
        return np.asarray(c.bc.inletQ_TotalQ, dtype=np.float64) * 0.8
    return None


def _compare(c, seeds, tol=2e-13):
    flat = R.flatten(c)
    o = Oracle(flat)
    rng = np.random.default_rng(17)
    worst = 0.0
    for Q in [c.Q0] + [cases.random_state_flat(flat, s, dry_frac=0.08) for s in seeds]:
        sc = cases.flat_scale(flat, Q)
        for kind, name, code in MODES:
            p = _params(c, kind, rng)
            if kind == "q" and p.size == 0:
                continue
            want = o.rhs(Q, p, code)
            got = LIT.swe_2d_rhs(c, Q, p, name)
            err = np.abs(got - want) / sc
            worst = max(worst, float(err.max()))
            assert (err <= tol).all(), (name, float(err.max()))
    return worst


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
def test_literal_restatement_agrees_on_the_reference_fixtures(name, oracle_lib):
    c = cases.load(name)
    worst = _compare(c, seeds=(3,) if name == "savannah" else (3, 4))
    print(f"{name}: literal vs C++ oracle, worst error {worst:.2e} of the flux scale")


@pytest.mark.parametrize("seed", [0, 1])
def test_literal_restatement_agrees_with_symmetry_and_two_inlets(tmp_path, seed, oracle_lib):
    _write_random_case(str(tmp_path), seed)
    c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    assert len(c.bc.kinds["symm"]) == 1 and len(c.bc.kinds["inletQ"]) == 2
    _compare(c, seeds=(5, 6, 7))


def test_roe_flux_every_branch(oracle_lib):
    """Single-face fuzz: 4e4 random left / right states, a third of the sides at or below h_small (exact ties included), random
    bed steps so that both 'virtual wall' branches, both one-sided branches, the dry-dry branch and the main branch all occur."""
    c = cases.load("simple")
    o = Oracle(R.flatten(c))
    rng = np.random.default_rng(99)
    hmin, g = 1e-3, 9.81
    seen = {}
    worst = 0.0
    for _ in range(40000):
        th = rng.uniform(0, 2 * np.pi)
        nx, ny = np.cos(th), np.sin(th)
        side = []
        for _s in range(2):
            h = float(np.exp(rng.uniform(np.log(1e-3), np.log(10.0))))
            if rng.random() < 0.33:
                h = float(rng.choice([1e-3, 5e-4, 9.999e-4]))          # the clamp leaves h = h_small; smaller values only occur in ghosts
                h = max(h, 1e-3) if rng.random() < 0.7 else h
            dry = h <= hmin
            sp, a = rng.uniform(0, 3), rng.uniform(0, 2 * np.pi)
            hu, hv = (0.0, 0.0) if dry and rng.random() < 0.8 else (h * sp * np.cos(a), h * sp * np.sin(a))
            zb = rng.uniform(-1, 1) if rng.random() < 0.7 else rng.uniform(-12, 12)
            hstill = rng.uniform(0.0, 5.0)
            side.append((h - hstill, hstill, h, hu, hv, zb))
        (xiL, hsL, hL, huL, hvL, zL), (xiR, hsR, hR, huR, hvR, zR) = side
        if hL <= hmin and hR <= hmin:
            br = "dry-dry"
        elif (hL + zL) < (zR + hmin) and hR <= hmin:
            br = "wall-R"
        elif (hR + zR) < (zL + hmin) and hL <= hmin:
            br = "wall-L"
        elif hL <= hmin:
            br = "dry-L"
        elif hR <= hmin:
            br = "dry-R"
        else:
            br = "wet"
        seen[br] = seen.get(br, 0) + 1
        want = o.roe(list(side[0]) + list(side[1]), nx, ny, g, hmin)
        got = LIT.riemann_2d_roe(*side[0], *side[1], g, (nx, ny), hmin)
        scale = 0.5 * g * (max(hL, hR) ** 2 + xiL ** 2 + xiR ** 2 + 2 * abs(xiL) * hsL + 2 * abs(xiR) * hsR) \
            + (np.hypot(huL, hvL) + np.hypot(huR, hvR)) * (3.0 + np.sqrt(g * max(hL, hR))) + 1e-12
        err = float(np.abs(got - want).max() / scale)
        worst = max(worst, err)
        assert err <= 1e-13, (br, side, got, want)
    assert all(seen.get(b, 0) >= 100 for b in ("dry-dry", "wall-R", "wall-L", "dry-L", "dry-R", "wet")), seen
    print(f"Roe flux, branches seen {seen}, worst error {worst:.2e} of the flux scale")


def test_euler_stepper_with_the_xi_mask_quirk(oracle_lib):
    """custom_ODE_update_cells: 40 steps from a state whose xi dips below h_small in places (the reference masks on xi, not h)."""
    c = cases.load("oneD_bump")
    o = Oracle(R.flatten(c))
    N = c.mesh.numOfCells
    Q = c.Q0.copy()
    Q[:N] = np.where(np.arange(N) % 7 == 0, -0.02, Q[:N])                    # xi below h_small although h = xi + hstill is not
    dt, nsteps = 2e-3, 40
    want = o.euler(Q, dt, nsteps)
    got = Q.copy()
    masked = 0
    for _ in range(nsteps):
        got = LIT.custom_ODE_update_cells(c, got, None, dt)
        masked += int((got[:N] == c.h_small).sum())
    assert masked > 0                                                        # the quirk is exercised
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def _complex_step_check(c, seeds, tol=5e-12):
    """imag(literal(Q + i e v, p + i e pdot)) / e against the oracle's dual-number pass J_Q v + J_p pdot."""
    flat = R.flatten(c)
    o = Oracle(flat)
    rng = np.random.default_rng(23)
    e = 1e-30
    N = c.mesh.numOfCells
    worst = 0.0
    for s in seeds:
        Q = cases.random_state_flat(flat, s, dry_frac=0.08)
        for kind, name, code in MODES:
            p = _params(c, kind, rng)
            if kind == "q" and p.size == 0:
                continue
            v = rng.standard_normal(3 * N)
            pdot = rng.standard_normal(p.size) * (0.01 if kind != "q" else 1.0) if p is not None else None
            want = o.jvp(Q, v, p, pdot, code)[1]
            got = np.imag(LIT.swe_2d_rhs(c, Q + 1j * e * v, None if p is None else p + 1j * e * pdot, name)) / e
            err = float(np.abs(got - want).max() / np.abs(want).max())
            worst = max(worst, err)
            assert err <= tol, (name, err)
    return worst


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
def test_oracle_dual_numbers_against_complex_step_of_the_literal_restatement(name, oracle_lib):
    """The brute-force J^T lambda every VJP test compares with is assembled from the oracle's dual-number passes.  Here those
    passes are checked without any AD: the complex-step derivative of the independently written restatement (exact to
    rounding; predicates and clamps select on the real part, as ForwardDiff treats them) in all four parameter modes."""
    c = cases.load(name)
    worst = _complex_step_check(c, seeds=(11,) if name == "savannah" else (11, 12))
    print(f"{name}: oracle dual pass vs complex step of the literal restatement, worst {worst:.2e} of the largest entry")


def test_oracle_dual_numbers_with_symmetry_and_two_inlets(tmp_path, oracle_lib):
    _write_random_case(str(tmp_path), 2)
    c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    _complex_step_check(c, seeds=(13, 14))


class _OracleBackedContext:
    """Stand-in with the device Context's rhs / rhs_vjp signatures, backed by the C++ oracle: lets the bodies the device tests
    use (tests/literal_checks.py) run on the CPU."""
    CODE = {None: 0, "zb": 1, "ManningN": 2, "Q": 3}

    def __init__(self, flat):
        self.o = Oracle(flat)

    def rhs(self, Q, params=None, active=None):
        return self.o.rhs(Q, params, self.CODE[active])

    def rhs_vjp(self, Q, lam, params=None, active=None):
        return self.o.vjp_bruteforce(Q, lam, params, self.CODE[active])[:2]


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
def test_shared_device_check_bodies_run_on_the_oracle(name, oracle_lib):
    from tests import literal_checks as LC
    assert LC.check_rhs(_OracleBackedContext, name, 2e-13) < 2e-13
    assert LC.check_vjp_identity(_OracleBackedContext, name, 1e-11) < 1e-11


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "oneD_uniform"])
def test_literal_restatement_has_its_own_pins_to_the_reference_files(name):
    """Not only through the C++ oracle: the literal friction terms reproduce the reference's friction_x/y_truth, and its RHS at
    the reference's saved final state of the sensitivity run is at the steady-state residual level."""
    c, t = cases.load(name), cases.truth(name)
    hs = c.h_small
    h = t["h_truth"]
    qx, qy = t["u_truth"] * (h + hs), t["v_truth"] * (h + hs)   # process_forward_simulation_results_2D.jl:32-33
    fx, fy = LIT.compute_friction_terms(h, qx, qy, t["ManningN_cells_truth"], c.g, c.k_n, hs)
    assert np.abs(fx - t["friction_x_truth"]).max() <= 6e-16 * np.abs(t["friction_x_truth"]).max()
    assert np.abs(fy - t["friction_y_truth"]).max() <= 6e-16 * max(np.abs(t["friction_y_truth"]).max(), 1e-30) + 1e-30
    if name == "oneD_bump":
        cs = cases.load("oneD_bump_sens")
        tj = np.load(cases.GOLD + "/oneD_bump_sens/trajectory.npz")["forward_simulation_results"]
        p = np.array([0.03, 0.02, 0.03])
        r = [np.abs(LIT.swe_2d_rhs(cs, tj[k], p, "ManningN")).max() for k in (0, 2)]
        assert r[0] > 1.0 and r[1] < 5e-5


def test_literal_restatement_reproduces_the_reference_transient_on_its_own():
    """The hard pin of tests/test_oracle_golden.py::test_reference_trajectory_hard_pin without the C++ oracle anywhere: the literal
    restatement (values and, by complex step, the partials with respect to the two Manning zones) integrated with the restated
    Tsit5 the way the reference ran (Dual-aware error norm, fastpow, dense-output saves) lands on the reference's saved state of
    the uniform-flow channel at t = 1 s -- about 25 accepted steps x 7 stages on a moving transient."""
    from tests import tsit5_ref as T
    name, p = "oneD_uniform_sens", np.array([0.03, 0.03])
    c = cases.load(name)
    tj = np.load(cases.GOLD + f"/{name}/trajectory.npz")
    idx, ref = tj["early_index"], tj["forward_simulation_results_early"]
    assert idx[0] == 1 and np.array_equal(tj["forward_simulation_results"][0], c.Q0)
    K, N = p.size, c.mesh.numOfCells
    e = 1e-30

    def rhs(U):
        out = np.empty_like(U)
        for k in range(K):
            ek = np.zeros(K)
            ek[k] = 1.0
            z = LIT.swe_2d_rhs(c, U[0] + 1j * e * U[1 + k], p + 1j * e * ek, "ManningN")
            out[0], out[1 + k] = z.real, z.imag / e
        return out

    U0 = np.zeros((1 + K, 3 * N))
    U0[0] = c.Q0
    _, saves, st = T.solve(rhs, U0, 0.0, 1.3, 0.02, True, 1e-6, 1e-3, np.array([1.0]), saveat="interp", norm=T.dual_norm, pow="fastpow")
    err = max(np.abs(saves[0][0][:N] - ref[0][:N]).max(), np.abs(saves[0][0][N:2 * N] - ref[0][N:2 * N]).max())
    print(f"literal restatement + Tsit5 vs the reference's saved state at t = 1 s: {err:.1e}  {st}")
    assert np.abs(ref[0][:N] - c.Q0[:N]).max() > 1e-3          # a real transient
    assert err <= 2e-9
