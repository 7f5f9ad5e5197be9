"""GPU parity tests: the CUDA paths (through the C ABI) against the oracle on the same inputs.

Tolerances (BASELINE.md "parity gates", north_star): <= 1e-12 relative per RHS call, where "relative" is
to the un-cancelled magnitude sum|flux L|/A + |source| of the cell (cases.flat_scale) -- the residual itself
cancels to ~0 near steady state; <= 1e-9 on h after N explicit steps.  The plain/strict path follows the
reference's operation order with FMA contraction off and is held to 2e-14 (libm pow / device sqrt ulps).
"""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases

pytestmark = pytest.mark.gpu

TOL_FUSED = 1e-12
TOL_PLAIN = 2e-14


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


_flat_cache = {}


def fixture_flat(name):
    if name not in _flat_cache:
        _flat_cache[name] = R.flatten(cases.load(name))
    return _flat_cache[name]


def synth(name):
    if name not in _flat_cache:
        from hydrograd_jl_b200 import synthetic as S
        _flat_cache[name] = {"dam": lambda: S.dam_break(48), "dam_thin": lambda: S.dam_break(48, thin_film=True),
                             "river": lambda: S.river(120, 40)}[name]()
    return _flat_cache[name]


def rel_err(flat, Q, got, ref):
    return (np.abs(got - ref) / cases.flat_scale(flat, Q)).max()


def states_for(name, flat):
    c = cases.load(name)
    out = [("ic", c.Q0)]
    out += [(f"fuzz{s}", cases.random_state_flat(flat, s)) for s in range(4)]
    return out


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "simple", "oneD_uniform"])
@pytest.mark.parametrize("path", ["plain", "fused"])
def test_rhs_fixture_meshes(hg, name, path):
    flat = fixture_flat(name)
    o = Oracle(flat)
    ctx = hg.Context(flat, strict=(path == "plain"), tile_cells=128)
    tol = TOL_PLAIN if path == "plain" else TOL_FUSED
    for label, Q in states_for(name, flat):
        ref = o.rhs(Q)
        got = ctx.rhs(Q)
        assert rel_err(flat, Q, got, ref) <= tol, (name, path, label)
        assert np.array_equal(got, ctx.rhs(Q)), "not bit-reproducible run to run"


def test_rhs_at_reference_final_state(hg):
    """The committed truth state of the Savannah forward run (u,v,h -> q) through both paths."""
    c, t = cases.load("savannah"), cases.truth("savannah")
    flat = fixture_flat("savannah")
    h = t["h_truth"]
    Q = np.concatenate([t["xi_truth"], t["u_truth"] * (h + c.h_small), t["v_truth"] * (h + c.h_small)])
    ref = Oracle(flat).rhs(Q)
    for strict in (True, False):
        got = hg.Context(flat, strict=strict).rhs(Q)
        assert rel_err(flat, Q, got, ref) <= (TOL_PLAIN if strict else TOL_FUSED)


@pytest.mark.parametrize("name", ["dam", "dam_thin", "river"])
def test_rhs_synthetic_meshes(hg, name):
    flat, Q0 = synth(name)
    o = Oracle(flat)
    for strict in (True, False):
        ctx = hg.Context(flat, strict=strict, tile_cells=256)
        for Q in [Q0] + [cases.random_state_flat(flat, s) for s in (5, 6)]:
            ref = o.rhs(Q)
            assert rel_err(flat, Q, ctx.rhs(Q), ref) <= (TOL_PLAIN if strict else TOL_FUSED), (name, strict)


def test_tiling_and_ordering_do_not_change_bits(hg):
    """Face-once evaluation with a canonical orientation + per-cell CSR order => the result does not depend
    on how cells are renumbered or tiled."""
    flat, Q0 = synth("river")
    Q = cases.random_state_flat(flat, 11)
    base = hg.Context(flat, tile_cells=512, reorder=True).rhs(Q)
    for tile, reorder, threads in ((128, True, 0), (256, True, 128), (256, True, 256), (512, False, 384), (128, False, 0),
                                  (256, True, 1010), (256, True, 1011), (512, True, 256), (224, True, 0), (224, True, 160), (384, True, 0)):   # two-faces-per-trip configurations
        got = hg.Context(flat, tile_cells=tile, reorder=reorder, threads=threads).rhs(Q)
        assert np.array_equal(base, got), (tile, reorder, threads)


@pytest.mark.parametrize("strict", [True, False])
def test_parameter_binding(hg, strict):
    """params_vector as zb / ManningN / Q (semi_discretize_swe_2D.jl:114-126, 153-161, 190-199)."""
    c = cases.load("savannah")
    flat = fixture_flat("savannah")
    o = Oracle(flat)
    ctx = hg.Context(flat, strict=strict, tile_cells=128)
    tol = TOL_PLAIN if strict else TOL_FUSED
    rng = np.random.default_rng(5)
    Q = cases.random_state_flat(flat, 21, dry_frac=0.02)
    Q[:c.mesh.numOfCells] = np.maximum(Q[:c.mesh.numOfCells], 0.05 - c.hstill)   # keep the inlet wet
    sets = [("ManningN", 2, c.ManningN_zone * (1 + 0.2 * rng.uniform(-1, 1, 6))),
            ("zb", 1, c.zb_cells + 0.05 * rng.standard_normal(c.zb_cells.size)),
            ("Q", 3, np.array([150.0])),
            (None, 0, None),
            ("ManningN", 2, c.ManningN_zone)]
    for name, code, p in sets:
        ref = o.rhs(Q, p, code)
        got = ctx.rhs(Q, p, name)
        assert rel_err(flat, Q, got, ref) <= tol, name


def test_euler_steps_match_oracle(hg):
    """custom_ODE_update_cells incl. the xi-mask quirk; <= 1e-9 on h after N steps."""
    for name, dt, n in (("oneD_bump", 0.005, 400), ("savannah", 0.01, 200)):
        c = cases.load(name)
        flat = fixture_flat(name)
        ctx = hg.Context(flat, tile_cells=128)
        ctx.set_state(c.Q0)
        ctx.step_euler(dt, n)
        got = ctx.get_state()
        ref = Oracle(flat).euler(c.Q0, dt, n)
        N = c.mesh.numOfCells
        assert np.abs(got[:N] - ref[:N]).max() <= 1e-9, name
        assert np.abs(got[N:] - ref[N:]).max() <= 1e-9 * max(1.0, np.abs(ref[N:]).max()), name


def test_euler_dry_mask_quirk(hg):
    flat, Q0 = synth("dam_thin")
    ctx = hg.Context(flat, tile_cells=256)
    ctx.set_state(Q0)
    ctx.step_euler(0.01, 50)
    got = ctx.get_state()
    ref = Oracle(flat).euler(Q0, 0.01, 50)
    N = flat["n_cells"]
    assert np.abs(got[:N] - ref[:N]).max() <= 1e-9
    assert ((got[:N] == flat["h_small"]) == (ref[:N] == flat["h_small"])).all()   # same masked cells


def test_custom_ode_solve_shape_and_values(hg):
    c = cases.load("simple")
    flat = fixture_flat("simple")
    p = hg.SWE2D_Extra_Parameters(flat, swe_2D_constants=hg.swe_2D_consts(dt=0.01, tspan=(0.0, 0.1)))
    sol = hg.custom_ODE_solve(None, c.Q0, None, p)
    assert sol.shape == (3 * 12, 11)                         # length(0:0.01:0.1) = 11 saved states
    ref = Oracle(flat).euler(c.Q0, 0.01, 11)
    assert np.abs(sol[:, -1] - ref).max() <= 1e-11
    out = np.zeros(36)
    p2 = hg.SWE2D_Extra_Parameters(flat, bInPlaceODE=True)
    r = hg.swe_2d_rhs(out, c.Q0, None, 0.0, p2)
    assert r is out and np.abs(out - Oracle(flat).rhs(c.Q0)).max() <= 1e-12


def test_error_behaviour(hg):
    flat = dict(fixture_flat("savannah"))
    ctx = hg.Context(flat)
    N = flat["n_cells"]
    with pytest.raises(hg.HydrogradError) as e:                 # wrong array length
        ctx.rhs(np.zeros(3 * N + 1))
    assert e.value.code == 1
    with pytest.raises(hg.HydrogradError) as e:                 # wrong params length
        ctx.rhs(np.zeros(3 * N), np.ones(3), "ManningN")
    assert e.value.code == 1
    dry = np.concatenate([-flat["hstill"], np.zeros(2 * N)])    # every inlet cell dry -> conveyance assert
    with pytest.raises(hg.HydrogradError) as e:
        ctx.rhs(dry)
    assert e.value.code == 3
    with pytest.raises(RuntimeError):
        Oracle(flat).rhs(dry)                                   # the reference asserts too (bc_2D.jl:678-680)
    assert np.isfinite(ctx.rhs(cases.load("savannah").Q0)).all()  # the context survives the error
    bad = dict(flat, riemann_solver="HLL")
    with pytest.raises(hg.HydrogradError) as e:
        hg.Context(bad)
    assert e.value.code == 4 and "not implemented" in str(e.value)


def test_million_cell_properties(hg):
    """Full-size C2 mesh (1M cells): properties that need no oracle pass -- lake at rest stays at rest,
    mass is conserved by the interior fluxes with all-wall boundaries, strict and fused paths agree."""
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.dam_break(1000)
    N = flat["n_cells"]
    ctx = hg.Context(flat)
    rest = np.concatenate([np.full(N, 0.7), np.zeros(2 * N)])
    assert np.abs(ctx.rhs(rest)).max() < 1e-12
    Q = cases.random_state_flat(flat, 3, dry_frac=0.0)
    dQ = ctx.rhs(Q)
    mass_rate = (dQ[:N] * flat["cell_areas"]).sum()
    assert abs(mass_rate) <= 1e-9 * np.abs(dQ[:N] * flat["cell_areas"]).sum()
    strict = hg.Context(flat, strict=True).rhs(Q)
    assert rel_err(flat, Q, dQ, strict) <= TOL_FUSED
    ref = Oracle(flat).rhs(Q, nthreads=0)
    assert rel_err(flat, Q, dQ, ref) <= TOL_FUSED


def test_product_reader_end_to_end(hg):
    """Savannah through the PRODUCT's own SRH-2D reader (C++), initial condition from the committed JSON, fused RHS,
    checked against the oracle built by the independent numpy reader."""
    import os
    from hydrograd_jl_b200 import srh2d
    flat = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, "savannah"), "savana_SI.srhhydro")
    ic = np.load(os.path.join(cases.GOLD, "savannah", "ic.npz"))
    Q0 = srh2d.setup_initial_condition(flat, ic["wse"], ic["wstill"], ic["q_x"], ic["q_y"])
    c = cases.load("savannah")
    assert np.array_equal(Q0, c.Q0)
    got = hg.Context(flat).rhs(Q0)
    ref = Oracle(fixture_flat("savannah")).rhs(c.Q0)
    assert rel_err(fixture_flat("savannah"), Q0, got, ref) <= TOL_FUSED


def test_rk4_stepper(hg):
    """Device-resident classical RK4 == the same tableau driven by the oracle RHS on the host."""
    c = cases.load("oneD_bump")
    flat = fixture_flat("oneD_bump")
    o = Oracle(flat)
    dt, n = 0.01, 50
    Q = c.Q0.copy()
    for _ in range(n):
        k1 = o.rhs(Q); k2 = o.rhs(Q + 0.5 * dt * k1); k3 = o.rhs(Q + 0.5 * dt * k2); k4 = o.rhs(Q + dt * k3)
        Q = Q + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_state(c.Q0)
    ctx.step_rk4(dt, n)
    got = ctx.get_state()
    assert np.abs(got - Q).max() <= 1e-9 * max(1.0, np.abs(Q).max())


@pytest.mark.parametrize("tile", [128, 256])
def test_l2_prefetch_does_not_change_bits(hg, tile):
    """The L2 prefetch (hg_options.reserved[3]) only moves data earlier: identical bits with it off, at the default
    distance and at an odd distance, for the RHS, fused Euler steps and the VJP (mesh with more tiles than resident CTAs)."""
    from hydrograd_jl_b200 import synthetic as S
    key = "river_big"
    if key not in _flat_cache:
        _flat_cache[key] = S.river(700, 260)
    flat, Q0 = _flat_cache[key]
    Q = cases.random_state_flat(flat, 3, dry_frac=0.03)
    lam = np.random.default_rng(4).standard_normal(Q.size)
    a = hg.Context(flat, tile_cells=tile, prefetch=-1)
    others = [hg.Context(flat, tile_cells=tile, prefetch=0), hg.Context(flat, tile_cells=tile, prefetch=37)]
    for q in (Q0, Q):
        ref = a.rhs(q)
        ref_v = a.rhs_vjp(q, lam)
        for b in others:
            assert np.array_equal(ref, b.rhs(q)), tile
            got_v = b.rhs_vjp(q, lam)
            for x, y in zip(ref_v, got_v):
                if x is not None:
                    assert np.array_equal(x, y)
    a.set_state(Q0)
    a.step_euler(1e-3, 7)
    for b in others:
        b.set_state(Q0)
        b.step_euler(1e-3, 7)
        assert np.array_equal(a.get_state(), b.get_state())


def test_fast_math(hg):
    """The kernels' branch-free helpers (hg_device.cuh: MUFU seed + ONE third-order step) against numpy on 2e5 positive
    normals spanning the ranges the path produces (depths 1e-6 .. 1e3, eps-sized squares, g h + eps): <= 2 ulp for 1/x,
    x^-1/2, sqrt, smooth abs; <= 16 ulp for x^(-7/3) (seven factors of a 1-ulp cube root).  The parity budget of the RHS is 1e-12."""
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.dam_break(8)
    ctx = hg.Context(flat)
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-12, 6, 100000), rng.uniform(0.5, 2.0, 50000), 10.0 ** rng.uniform(-300, 300, 50000)])
    ulp = lambda got, ref: float((np.abs(got - ref) / np.spacing(np.abs(ref))).max())
    xl = x.astype(np.longdouble)
    xs = np.concatenate([x[:150000], -x[:1000], [0.0, 1e-9, -1e-8]])      # smooth abs takes any sign, incl. 0 -> sqrt(eps)
    xp = x[:150000]
    err = {
        "rcp": ulp(ctx.debug_math(0, x), (1 / xl).astype(np.float64)),
        "rsqrt": ulp(ctx.debug_math(1, x), (1 / np.sqrt(xl)).astype(np.float64)),
        "sqrt": ulp(ctx.debug_math(2, x), np.sqrt(xl).astype(np.float64)),
        "smooth_abs": ulp(ctx.debug_math(3, xs), np.sqrt(xs.astype(np.longdouble) ** 2 + np.longdouble(2.220446049250313e-16)).astype(np.float64)),
        "pow_m73": ulp(ctx.debug_math(4, xp), (xp.astype(np.longdouble) ** (np.longdouble(-7) / 3)).astype(np.float64)),
    }
    print("fast math max ulp:", err)
    assert err["rcp"] <= 2 and err["rsqrt"] <= 2 and err["sqrt"] <= 1 and err["smooth_abs"] <= 2 and err["pow_m73"] <= 16, err
