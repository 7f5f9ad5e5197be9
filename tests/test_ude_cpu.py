"""CPU check of the UDE network arithmetic: the per-cell functions of hydrograd.jl_b200/csrc/hg_ude.h -- the same source the
CUDA kernels of hg_ude.cu are built from -- compiled by g++ (tests/ude_host.cpp) and compared with the numpy restatement
of the reference's closure (oracle/ude_ref.py), values and complex-step derivatives.  The kernels themselves (reductions,
launch structure) are covered by tests/test_gpu_zz_ude.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _pkg
from oracle import ude_ref as U

HERE = os.path.dirname(os.path.abspath(__file__))
hg = _pkg.load()
from hydrograd_jl_b200 import _lib as L            # noqa: E402
from hydrograd_jl_b200 import ude as hude          # noqa: E402

f64p = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    out = os.path.join(HERE, "..", "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libude_host.so")
    src = os.path.join(HERE, "ude_host.cpp")
    hdr = os.path.join(HERE, "..", "hydrograd.jl_b200", "csrc", "hg_ude.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(f64p) if a is not None else None


CONFIGS = [
    ("ManningN_h", [3, 3], ["tanh", "tanh"], "whole"),
    ("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "whole"),           # the two shipped run_control.json
    ("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "cell"),
    ("ManningN_h_Umag_ks", [8, 5, 2], ["softplus", "sigmoid", "leakyrelu"], "whole"),
    ("ManningN_h", [4], ["relu"], "cell"),
    ("ManningN_h_Umag_ks", [6, 6], ["leakyrelu", "softplus"], "none"),
]


def _case(choice, hidden, acts, ln, N=57, seed=3):
    rng = np.random.default_rng(seed)
    cfg = dict(input_dim=1 if choice == "ManningN_h" else 3, output_dim=1, hidden_layers=hidden, activations=acts,
               h_bounds=[0.1, 2.5], Umag_bounds=[0.05, 1.5], ks_bounds=[0.02, 0.3], output_bounds=[0.02, 0.06])
    pm = hude.UDEModel(choice, cfg, layernorm=ln)
    om = U.Model(choice, hidden, acts, ln, cfg["h_bounds"], cfg["output_bounds"], cfg["Umag_bounds"], cfg["ks_bounds"])
    assert pm.n_params == om.n_params
    assert list(pm.desc.off_weight[:len(hidden) + 1]) == om.off_w and list(pm.desc.off_bias[:len(hidden) + 1]) == om.off_b
    if ln != "none":
        assert list(pm.desc.off_ln_scale[:len(hidden)]) == om.off_g and list(pm.desc.off_ln_bias[:len(hidden)]) == om.off_be
    th = om.init_theta(rng)
    hstill = rng.uniform(0.2, 2.0, N)
    xi = rng.uniform(-0.1, 0.4, N)
    dry = rng.random(N) < 0.15
    xi[dry] = -hstill[dry] + rng.uniform(-1e-3, 1e-3, dry.sum())          # some cells below / at the clamp
    Q = np.concatenate([xi, rng.uniform(-1.0, 1.0, N), rng.uniform(-0.5, 0.5, N)])
    ks = rng.uniform(0.02, 0.3, N)
    return pm, om, th, Q, hstill, ks, 1e-3


# which compile-time instantiation (hg_ude.h HG_UDE_SPECS) serves each configuration; 0 = generic
SPEC = [1, 2, 4, 0, 0, 0]


def test_spec_offsets_equal_the_canonical_layout(host):
    assert host.ude_host_spec_check() == 0


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("k", range(len(CONFIGS)))
def test_network_values_match_the_restatement(host, k, generic):
    choice, hidden, acts, ln = CONFIGS[k]
    pm, om, th, Q, hstill, ks, hs = _case(choice, hidden, acts, ln)
    N = hstill.size
    n = np.empty(N)
    spec = C.c_int(-1)
    assert host.ude_host_manning(C.byref(pm.desc), C.c_int64(N), _p(Q), _p(hstill), _p(ks), C.c_double(hs), _p(th), _p(n),
                                 generic, C.byref(spec)) == 0
    assert spec.value == (0 if generic else SPEC[k])
    ref = om.manning(Q, th, hstill, ks, hs)
    assert ref.min() > 0.02 and ref.max() < 0.06 and np.ptp(ref) > 1e-4
    assert np.abs(n - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("choice,hidden,acts,ln", CONFIGS)
def test_network_pullback_matches_complex_step(host, choice, hidden, acts, ln, generic):
    pm, om, th, Q, hstill, ks, hs = _case(choice, hidden, acts, ln)
    N = hstill.size
    nbar = np.random.default_rng(5).normal(size=N)
    Qb, tb = np.zeros(3 * N), np.zeros(pm.n_params)
    assert host.ude_host_pullback(C.byref(pm.desc), C.c_int64(N), _p(Q), _p(hstill), _p(ks), C.c_double(hs), _p(th), _p(nbar),
                                  _p(Qb), _p(tb), generic) == 0
    Qr, tr = om.pullback(Q, th, hstill, ks, hs, nbar)
    assert np.abs(tr).max() > 1e-4 and np.abs(Qr).max() > 1e-5
    assert np.abs(tb - tr).max() <= 1e-11 * np.abs(tr).max()
    assert np.abs(Qb - Qr).max() <= 1e-11 * np.abs(Qr).max()
    wet = (Q[:N] + hstill) > hs
    assert np.all(Qb[:N][~wet] == 0.0) and np.all(Qb[N:2 * N][~wet] == 0.0)       # the clamp passes nothing back


def test_descriptor_validation(host):
    pm = _case("ManningN_h", [3, 3], ["tanh", "tanh"], "whole")[0]
    assert host.ude_host_max_params() == 233
    import copy
    for field, val in (("n_hidden", 0), ("n_hidden", 4), ("choice", 3), ("layernorm", 3), ("n_params", 10)):
        d = L.UdeDesc.from_buffer_copy(pm.desc)
        setattr(d, field, val)
        N = 4
        z = np.zeros(3 * N)
        assert host.ude_host_manning(C.byref(d), C.c_int64(N), _p(z), _p(np.ones(N)), None, C.c_double(1e-3), _p(np.zeros(64)), _p(np.zeros(N)), 0, None) == 1
    with pytest.raises(ValueError, match="Unsupported activation"):
        hude.UDEModel("ManningN_h", dict(hidden_layers=[3], activations=["gelu"], h_bounds=[0, 1], output_bounds=[0, 1]))
    with pytest.raises(ValueError, match="Unknown UDE choice"):
        hude.UDEModel("FlowResistance", dict(hidden_layers=[3], activations=["tanh"], h_bounds=[0, 1], output_bounds=[0, 1]))


def test_composed_ude_rhs_gradient_against_finite_differences():
    """The reference for the GPU tests -- oracle RHS with one Manning value per cell composed with the network restatement --
    checked against central differences of the composed RHS itself (whole-array LayerNorm: n of a cell depends on all cells)."""
    from oracle import srh2d_ref as R
    from tests import cases
    c = cases.load("oneD_bump")
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    rng = np.random.default_rng(11)
    om = U.Model("ManningN_h_Umag_ks", [3, 3], ["tanh", "tanh"], "whole", [0.05, 0.6], [0.02, 0.06], [0.0, 1.5], [0.02, 0.3])
    th = om.init_theta(rng)
    ks = rng.uniform(0.02, 0.3, N)
    ur = U.UdeRhs(flat, om, ks)
    Q = c.Q0 + np.concatenate([0.02 * rng.normal(size=N), 0.05 + 0.02 * rng.random(N), 0.01 * rng.normal(size=N)])
    lam = rng.normal(size=3 * N)
    Qbar, tbar, _ = ur.vjp(Q, th, lam)
    v, w = rng.normal(size=3 * N), rng.normal(size=om.n_params)
    e = 1e-6
    fd = lam @ (ur.rhs(Q + e * v, th + e * w) - ur.rhs(Q - e * v, th - e * w)) / (2 * e)
    an = Qbar @ v + tbar @ w
    assert abs(tbar @ w) > 1e-6 * abs(an)                                   # the parameter path carries weight
    assert abs(fd - an) <= 1e-6 * max(abs(an), np.abs(Qbar * v).sum() * 1e-3)
