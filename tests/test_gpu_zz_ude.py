"""GPU parity of the UDE closure (hg_set_ude_model + active parameter "UDE"): Manning's n of every cell from the neural network
of the state, evaluated and differentiated on the device (csrc/hg_ude.cu), against the oracle composition -- the C++ oracle RHS
with one Manning value per cell and the numpy restatement of update_ManningN_UDE / the Lux chain (oracle/ude_ref.py).
RHS <= 1e-12 relative, gradients <= 1e-9 (north_star gates).  (File name: these tests were written after the round's GPU
budget was spent and sort last on purpose.)"""
import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from oracle import ude_ref as U
from tests import cases

pytestmark = pytest.mark.gpu

CFG = dict(h_bounds=[0.05, 3.0], Umag_bounds=[0.0, 1.5], ks_bounds=[0.02, 0.3], output_bounds=[0.02, 0.06])
MODELS = [
    ("ManningN_h", [3, 3], ["tanh", "tanh"], "whole"),                   # examples/SWE_2D/UDE/ManningN/oneD_channel_with_bump
    ("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "whole"),           # examples/SWE_2D/UDE/ManningN/Savannah_River_ks_h_Umag
    ("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "cell"),
    ("ManningN_h_Umag_ks", [8, 5, 2], ["softplus", "sigmoid", "leakyrelu"], "whole"),
    ("ManningN_h", [4], ["relu"], "none"),
]


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _models(choice, hidden, acts, ln, seed=0):
    from hydrograd_jl_b200 import ude as hude
    cfg = dict(CFG, input_dim=1 if choice == "ManningN_h" else 3, output_dim=1, hidden_layers=hidden, activations=acts)
    pm = hude.UDEModel(choice, cfg, layernorm=ln)
    om = U.Model(choice, hidden, acts, ln, CFG["h_bounds"], CFG["output_bounds"], CFG["Umag_bounds"], CFG["ks_bounds"])
    assert pm.n_params == om.n_params
    return pm, om, om.init_theta(np.random.default_rng(100 + seed))


def _state(flat, seed, dry_frac=0.06):
    """Depths and speeds inside the network's input bounds (so that no unit is saturated), a few cells at / below the clamp."""
    rng = np.random.default_rng(seed)
    N = int(flat["n_cells"])
    h = np.exp(rng.uniform(np.log(0.05), np.log(3.0), N))
    k = rng.random(N) < dry_frac
    h[k] = rng.choice([5e-4, 1e-3, 9.999e-4], size=int(k.sum()))
    ni = int(np.asarray(flat["bc_ptr"])[int(flat["n_inletq"])])
    ic = np.asarray(flat["bc_internal_cells"])[:ni] - int(flat["index_base"])
    h[ic] = np.maximum(h[ic], 0.05)                                         # positive inlet conveyance (bc_2D.jl:678-680)
    sp, th = rng.uniform(0.05, 1.5, N), rng.uniform(0, 2 * np.pi, N)
    return np.concatenate([h - flat["hstill"], h * sp * np.cos(th), h * sp * np.sin(th)])


@pytest.mark.parametrize("name", ["oneD_bump", "savannah"])
@pytest.mark.parametrize("choice,hidden,acts,ln", MODELS)
def test_ude_rhs_and_vjp_match_oracle(hg, name, choice, hidden, acts, ln):
    c = cases.load(name)
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    pm, om, th = _models(choice, hidden, acts, ln)
    rng = np.random.default_rng(7)
    ks = rng.uniform(0.02, 0.3, N)
    ur = U.UdeRhs(flat, om, ks)
    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_ude_model(pm, ks)
    for seed in (0, 1):
        Q = _state(flat, seed)
        ref = ur.rhs(Q, th)
        dQ = ctx.rhs(Q, th, "UDE")
        fl = dict(flat, ManningN_cells=ur.manning(Q, th))
        assert (np.abs(dQ - ref) / cases.flat_scale(fl, Q)).max() <= 1e-12, (name, seed)
        lam = rng.standard_normal(3 * N)
        Qbar_ref, tbar_ref, nbar_ref = ur.vjp(Q, th, lam)
        Qbar, tbar, nbar = ctx.rhs_vjp(Q, lam, th, "UDE", want_ncell_bar=True)
        assert np.abs(nbar - nbar_ref).max() <= 1e-9 * np.abs(nbar_ref).max()
        assert np.abs(tbar - tbar_ref).max() <= 1e-9 * np.abs(tbar_ref).max(), (name, seed)
        assert np.abs(Qbar - Qbar_ref).max() <= 1e-9 * np.abs(Qbar_ref).max(), (name, seed)
        Qbar2, tbar2 = ctx.rhs_vjp(Q, lam, th, "UDE")
        assert np.array_equal(Qbar, Qbar2) and np.array_equal(tbar, tbar2), "UDE VJP is not bit-reproducible"
        assert np.array_equal(dQ, ctx.rhs(Q, th, "UDE"))


@pytest.mark.parametrize("ln", ["whole", "cell"])
def test_ude_adjoint_identity_on_a_multi_block_mesh(hg, ln):
    """4.6k cells = 5 blocks of the UDE kernels (partial statistics / partial thetabar sums combined across blocks):
    RHS against the oracle composition, gradients through the identity lam . (J v + J_theta w) = Qbar . v + thetabar . w."""
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.river(96, 48)
    N = int(flat["n_cells"])
    pm, om, th = _models("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], ln, seed=1)
    rng = np.random.default_rng(9)
    ks = rng.uniform(0.02, 0.3, N)
    ur = U.UdeRhs(flat, om, ks)
    ctx = hg.Context(flat, tile_cells=256)
    ctx.set_ude_model(pm, ks)
    Q = _state(flat, 3)
    dQ = ctx.rhs(Q, th, "UDE")
    fl = dict(flat, ManningN_cells=ur.manning(Q, th))
    assert (np.abs(dQ - ur.rhs(Q, th)) / cases.flat_scale(fl, Q)).max() <= 1e-12
    lam, v, w = rng.standard_normal(3 * N), rng.standard_normal(3 * N), rng.standard_normal(om.n_params)
    Qbar, tbar = ctx.rhs_vjp(Q, lam, th, "UDE")
    lhs = lam @ ur.jvp(Q, th, v, w)
    rhs = Qbar @ v + tbar @ w
    scale = np.abs(Qbar * v).sum() + np.abs(tbar * w).sum()
    assert abs(lhs - rhs) <= 1e-9 * scale
    # the parameter path alone
    lhs_t = lam @ ur.jvp(Q, th, np.zeros(3 * N), w)
    assert abs(lhs_t - tbar @ w) <= 1e-9 * np.abs(tbar * w).sum()


def test_ude_time_stepping_and_adjoint_through_time(hg):
    """The network is re-evaluated from every stage state: device RK4 against the same tableau driven by the oracle
    composition, and the discrete adjoint of those steps (hg_rk_adjoint) against central differences of the oracle loop."""
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    pm, om, th = _models("ManningN_h", [3, 3], ["tanh", "tanh"], "whole", seed=2)
    ur = U.UdeRhs(flat, om)
    dt, nsteps = 0.01, 12

    def rk4(Q, th_):
        f = lambda u: ur.rhs(u, th_)
        for _ in range(nsteps):
            k1 = f(Q); k2 = f(Q + 0.5 * dt * k1); k3 = f(Q + 0.5 * dt * k2); k4 = f(Q + dt * k3)
            Q = Q + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        return Q

    ctx = hg.Context(flat, tile_cells=128)
    ctx.set_ude_model(pm)
    ctx.set_params(th, "UDE")
    ctx.set_state(c.Q0)
    ctx.step_rk4(dt, nsteps)
    ref = rk4(c.Q0.copy(), th)
    got = ctx.get_state()
    assert np.abs(got[:N] - ref[:N]).max() <= 1e-9 and np.abs(got[N:] - ref[N:]).max() <= 1e-9
    rng = np.random.default_rng(4)
    lamT = rng.standard_normal(3 * N)
    QT, Q0bar, tbar = ctx.rk_adjoint("RK4", c.Q0, lamT, dt, nsteps, th, "UDE")
    assert np.abs(QT - ref).max() <= 1e-9
    w = rng.standard_normal(om.n_params)
    e = 1e-6
    fd = lamT @ (rk4(c.Q0.copy(), th + e * w) - rk4(c.Q0.copy(), th - e * w)) / (2 * e)
    assert abs(fd - tbar @ w) <= 1e-5 * np.abs(tbar * w).sum()
    v = rng.standard_normal(3 * N) * 1e-2
    fdq = lamT @ (rk4(c.Q0 + e * v, th) - rk4(c.Q0 - e * v, th)) / (2 * e)
    assert abs(fdq - Q0bar @ v) <= 1e-5 * np.abs(Q0bar * v).sum()


def test_ude_error_behaviour_and_clearing(hg):
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    pm, om, th = _models("ManningN_h", [3, 3], ["tanh", "tanh"], "whole")
    ctx = hg.Context(flat, tile_cells=128)
    Q = _state(flat, 0)
    plain = ctx.rhs(Q)
    with pytest.raises(hg.HydrogradError, match="no model"):
        ctx.rhs(Q, th, "UDE")
    ctx.set_ude_model(pm)
    with pytest.raises(hg.HydrogradError, match="expected"):
        ctx.rhs(Q, th[:-1], "UDE")
    with pytest.raises(hg.HydrogradError, match="UDE or NONE"):
        ctx.rhs(Q, c.ManningN_zone, "ManningN")
    with pytest.raises(hg.HydrogradError, match="UDE model is set"):
        ctx.set_manning_function("sigmoid", 0.03, 0.06, 100.0, 0.3)
    pm3 = _models("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "whole")[0]
    with pytest.raises(hg.HydrogradError, match="ks_cells is NULL"):
        ctx.set_ude_model(pm3)
    with_net = ctx.rhs(Q, th, "UDE")
    assert not np.array_equal(with_net, plain)
    assert np.array_equal(ctx.rhs(Q), plain)                                # NONE: the bound ManningN_cells, no network
    ctx.rhs(Q, th, "UDE")
    ctx.set_ude_model(None)
    assert np.array_equal(ctx.rhs(Q), plain)                                # cleared: frozen fields restored
    strict = hg.Context(flat, strict=True)
    with pytest.raises(hg.HydrogradError, match="fused path"):
        strict.set_ude_model(pm)


@pytest.mark.parametrize("choice,ln", [("ManningN_h", "whole"), ("ManningN_h_Umag_ks", "whole"), ("ManningN_h", "cell"), ("ManningN_h_Umag_ks", "cell")])
def test_ude_specialised_kernels_agree_with_the_generic_ones(hg, choice, ln):
    """The shapes the reference ships run kernels with a compile-time network shape (tape in registers); hg_options.reserved[5]
    forces the generic kernels.  Same arithmetic in the same order: agreement to rounding (FMA contraction may differ)."""
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    pm, om, th = _models(choice, [3, 3], ["tanh", "tanh"], ln, seed=5)
    rng = np.random.default_rng(12)
    ks = rng.uniform(0.02, 0.3, N)
    Q = _state(flat, 2)
    lam = rng.standard_normal(3 * N)
    out = []
    for generic in (False, True):
        ctx = hg.Context(flat, tile_cells=128, ude_generic=generic)
        ctx.set_ude_model(pm, ks)
        dQ = ctx.rhs(Q, th, "UDE")
        Qbar, tbar = ctx.rhs_vjp(Q, lam, th, "UDE")
        out.append((dQ, Qbar, tbar))
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
