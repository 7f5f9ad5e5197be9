"""CPU check of the forward-mode (JVP) arithmetic: hydrograd.jl_b200/csrc/hg_jvp_impl.h -- the very source the CUDA kernels of
hg_jvp.cu are compiled from -- built by g++ (tests/jvp_host.cpp) and compared with the oracle: values against its RHS, tangents
against its dual-number pass, for every active parameter, with dry cells and wet/dry fronts.  The launch structure of the
kernels is covered by tests/test_gpu_zzy_jvp.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))
f64p, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class Args(C.Structure):            # hg::jvp::Args
    _fields_ = ([(n, C.c_int32) for n in ("N", "B", "n_inlet", "active")] + [(n, C.c_double) for n in ("g", "k_n", "h_small")] +
                [(n, i32p) for n in ("cf_ptr", "cf_nb")] + [(n, f64p) for n in ("cf_nx", "cf_ny", "cf_len")] +
                [(n, f64p) for n in ("area", "hstill", "zb", "S0x", "S0y", "mann")] + [("matid", i32p)] +
                [(n, i32p) for n in ("bc_type", "bc_group", "bc_ghost", "bc_cell", "inlet_ptr")] +
                [(n, f64p) for n in ("bc_nx", "bc_ny", "bc_l53", "bc_l23", "hstill_g", "zb_g")] +
                [(n, f64p) for n in ("gh", "gqx", "gqy", "gxi", "gh_d", "gqx_d", "gqy_d", "gxi_d")] +
                [(n, f64p) for n in ("Qin", "wse", "Q", "V", "params", "pdot", "dQ", "dQ_d")] + [("err", i32p)])


@pytest.fixture(scope="module")
def host():
    out = os.path.join(HERE, "..", "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libjvp_host.so")
    src = os.path.join(HERE, "jvp_host.cpp")
    hdr = os.path.join(HERE, "..", "hydrograd.jl_b200", "csrc", "hg_jvp_impl.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.jvp_host.argtypes = [C.POINTER(Args), C.c_int]
    lib.jvp_host.restype = C.c_int
    return lib


def plain_tables(flat):
    """The plain-path tables of hg_host.cpp::build_host (cell-face CSR in reference order, boundary entries in processing
    order), restated with numpy."""
    N, ld, base, B = flat["n_cells"], flat["ld"], flat["index_base"], flat["n_ghost"]
    nf = np.asarray(flat["cell_nfaces"], dtype=np.int64)
    cf = np.asarray(flat["cell_faces"]).reshape(ld, N)
    nb = np.asarray(flat["cell_neighbors"]).reshape(ld, N)
    nrm = np.asarray(flat["cell_normals"]).reshape(2, ld, N)
    t = {}
    t["cf_ptr"] = np.concatenate([[0], np.cumsum(nf)]).astype(np.int32)
    cells = np.repeat(np.arange(N), nf)
    js = np.concatenate([np.arange(k) for k in nf])
    fid = np.abs(cf[js, cells]) - base
    isb = np.asarray(flat["face_is_boundary"])[fid].astype(bool)
    nbv = nb[js, cells] - base
    t["cf_nb"] = np.where(isb, N + nbv, nbv).astype(np.int32)
    t["cf_nx"], t["cf_ny"] = nrm[0, js, cells].copy(), nrm[1, js, cells].copy()
    t["cf_len"] = np.asarray(flat["face_lengths"])[fid].copy()
    ptr = np.asarray(flat["bc_ptr"], dtype=np.int64)
    counts = [flat["n_inletq"], flat["n_exith"], flat["n_wall"], flat["n_symm"]]
    typ, grp = np.zeros(B, np.int32), np.zeros(B, np.int32)
    kb = 0
    inlet_ptr = [0]
    for ty, cnt in enumerate(counts):
        for kk in range(cnt):
            typ[ptr[kb]:ptr[kb + 1]] = ty
            grp[ptr[kb]:ptr[kb + 1]] = kk
            if ty == 0:
                inlet_ptr.append(ptr[kb + 1])
            kb += 1
    t["bc_type"], t["bc_group"] = typ, grp
    t["bc_ghost"] = (np.asarray(flat["bc_ghost_ids"]) - base).astype(np.int32)
    t["bc_cell"] = (np.asarray(flat["bc_internal_cells"]) - base).astype(np.int32)
    t["inlet_ptr"] = np.asarray(inlet_ptr, dtype=np.int32)
    bn = np.asarray(flat["bc_normals"], dtype=np.float64)
    t["bc_nx"], t["bc_ny"] = bn[:B].copy(), bn[B:].copy()
    L = np.asarray(flat["bc_lengths"], dtype=np.float64) if flat["n_inletq"] else np.zeros(B)
    inl = typ == 0
    t["bc_l53"], t["bc_l23"] = np.where(inl, np.abs(L) ** (5.0 / 3.0), 0.0), np.where(inl, np.abs(L) ** (2.0 / 3.0), 0.0)
    return t


class HostJvp:
    """hg::jvp::Args for one mesh, filled once; call(...) runs the g++ build of hg_jvp_impl.h."""

    def __init__(self, host, flat):
        self.host, self.flat = host, flat
        N, B = flat["n_cells"], flat["n_ghost"]
        self.N = N
        t = plain_tables(flat)
        self.keep = []
        a = self.a = Args()
        a.N, a.B, a.n_inlet = N, B, flat["n_inletq"]
        a.g, a.k_n, a.h_small = flat["g"], flat["k_n"], flat["h_small"]
        for k in ("cf_ptr", "cf_nb", "bc_type", "bc_group", "bc_ghost", "bc_cell", "inlet_ptr"):
            setattr(a, k, self.i(t[k]))
        for k in ("cf_nx", "cf_ny", "cf_len", "bc_nx", "bc_ny", "bc_l53", "bc_l23"):
            setattr(a, k, self.d(t[k]))
        S0 = np.asarray(flat["S0_cells"], dtype=np.float64)
        a.area, a.hstill, a.zb, a.S0x, a.S0y, a.mann = (self.d(flat["cell_areas"]), self.d(flat["hstill"]), self.d(flat["zb_cells"]),
                                                         self.d(S0[:N]), self.d(S0[N:]), self.d(flat["ManningN_cells"]))
        a.matid = self.i(np.asarray(flat["matID_cells"]) if "matID_cells" in flat else np.zeros(N))
        a.hstill_g, a.zb_g = self.d(flat["hstill_ghost"]), self.d(flat["zb_ghost"])
        for k in ("gh", "gqx", "gqy", "gxi", "gh_d", "gqx_d", "gqy_d", "gxi_d"):
            setattr(a, k, self.d(np.zeros(max(B, 1))))
        a.Qin = self.d(flat["inletQ_TotalQ"] if flat["n_inletq"] else np.zeros(1))
        a.wse = self.d(flat["exitH_WSE"] if flat["n_exith"] else np.zeros(1))
        self.err = np.zeros(1, dtype=np.int32)
        a.err = self.err.ctypes.data_as(i32p)

    def d(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        self.keep.append(x)
        return x.ctypes.data_as(f64p)

    def i(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.int32))
        self.keep.append(x)
        return x.ctypes.data_as(i32p)

    def __call__(self, Q, V=None, params=None, pdot=None, active=0, dual=True):
        a = self.a
        hold = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (Q, V, params, pdot)]
        a.active = active
        a.Q, a.V, a.params, a.pdot = [x.ctypes.data_as(f64p) if x is not None else None for x in hold]
        out, out_d = np.zeros(3 * self.N), np.zeros(3 * self.N)
        a.dQ, a.dQ_d = out.ctypes.data_as(f64p), out_d.ctypes.data_as(f64p)
        self.err[0] = 0
        rc = self.host.jvp_host(C.byref(a), int(dual))
        return out, out_d, rc


def run_host(host, flat, Q, V=None, params=None, pdot=None, active=0, dual=True):
    return HostJvp(host, flat)(Q, V, params, pdot, active, dual)


def _flat(name):
    c = cases.load(name)
    return c, R.flatten(c)


ACTIVE = {"none": 0, "zb": 1, "ManningN": 2, "Q": 3}


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "simple"])
@pytest.mark.parametrize("mode", ["none", "zb", "ManningN", "Q"])
def test_forward_mode_matches_the_oracle_dual_pass(host, name, mode):
    c, flat = _flat(name)
    N = c.mesh.numOfCells
    o = Oracle(flat)
    rng = np.random.default_rng(11)
    params = {"none": None, "zb": c.zb_cells.copy(), "ManningN": np.asarray(c.ManningN_zone, dtype=np.float64).copy(),
              "Q": np.asarray(flat["inletQ_TotalQ"], dtype=np.float64).copy()}[mode]
    if mode == "Q" and flat["n_inletq"] == 0:
        pytest.skip("no inlet-q boundary")
    for seed in (0, 1):
        Q = cases.random_state(c, seed) if seed else c.Q0
        V = rng.standard_normal(3 * N)
        pdot = rng.standard_normal(params.size) if params is not None else None
        dQ, dQd, rc = run_host(host, flat, Q, V, params, pdot, ACTIVE[mode])
        assert rc == 0
        ref = o.rhs(Q, params, ACTIVE[mode])
        ref_d = o.jvp(Q, V, params, pdot, ACTIVE[mode])[1]
        sc = cases.flux_scale(c, Q)
        assert (np.abs(dQ - ref) <= 1e-13 * sc).all()
        # tangent scale: the same flux scale times the size of the direction (|V|, |pdot| ~ 1; zb / n amplify by 1/h, 1/n)
        den = np.abs(ref_d).max()
        assert np.abs(dQd - ref_d).max() <= 1e-12 * den, (name, mode, seed)
        # the double instantiation computes the same values, and no tangent
        dQ0, dQd0, _ = run_host(host, flat, Q, None, params, None, ACTIVE[mode], dual=False)
        assert np.array_equal(dQ0, dQ) and not dQd0.any()


def test_forward_mode_is_linear_and_transposes_to_the_brute_force_vjp(host):
    """<lambda, J v> = <J^T lambda, v> with the oracle's brute-force J^T lambda: ties the forward mode to what the VJP kernel is
    tested against; and J (a v1 + b v2) = a J v1 + b J v2."""
    c, flat = _flat("savannah")
    N = c.mesh.numOfCells
    o = Oracle(flat)
    rng = np.random.default_rng(5)
    Q = cases.random_state(c, 3)
    p = np.asarray(c.ManningN_zone, dtype=np.float64)
    v1, v2, lam = rng.standard_normal(3 * N), rng.standard_normal(3 * N), rng.standard_normal(3 * N)
    pd1, pd2 = rng.standard_normal(p.size), rng.standard_normal(p.size)
    _, j1, _ = run_host(host, flat, Q, v1, p, pd1, 2)
    _, j2, _ = run_host(host, flat, Q, v2, p, pd2, 2)
    _, j12, _ = run_host(host, flat, Q, 0.5 * v1 - 2.0 * v2, p, 0.5 * pd1 - 2.0 * pd2, 2)
    assert np.abs(j12 - (0.5 * j1 - 2.0 * j2)).max() <= 1e-12 * np.abs(j12).max()
    Qbar, pbar = o.vjp_bruteforce(Q, lam, p, 2)[:2]
    lhs, rhs = lam @ j1, Qbar @ v1 + pbar @ pd1
    assert abs(lhs - rhs) <= 1e-11 * (np.abs(lam * j1).sum())


def test_conveyance_assert_is_reported(host):
    c, flat = _flat("oneD_bump")
    N = c.mesh.numOfCells
    Q = c.Q0.copy()
    Q[:N] = -flat["hstill"] + 1e-4          # everything dry: the inlet's conveyance is 0 (bc_2D.jl:678-680)
    _, _, rc = run_host(host, flat, Q, np.ones(3 * N))
    assert rc == 3


def test_reference_sensitivity_run_through_the_forward_mode_source(host):
    """The reference's Savannah sensitivity run (ForwardDiff.jacobian of the adaptive solve: values and six partials through the
    Dual-norm Tsit5, 202 accepted steps x 7 stages x 6 partials) carried by the g++ build of hg_jvp_impl.h -- the source the
    device kernels are compiled from -- lands on the reference's committed sensitivity_results.json."""
    from tests.test_oracle_golden import savannah_sensitivity_solve
    c, flat = _flat("savannah")
    run = HostJvp(host, flat)

    def jvp(Q, V, p, pdot):
        f, jv, rc = run(Q, V, p, pdot, 2)
        assert rc == 0
        return f, jv

    U, S, st, _ = savannah_sensitivity_solve(jvp)
    err = [np.abs(U[1 + k] - S[k]).max() for k in range(S.shape[0])]
    print("savannah sensitivities through hg_jvp_impl.h vs reference:", ["%.1e" % e for e in err], st)
    assert st["accepted"] > 150 and max(err) <= 2e-8 * np.abs(S).max()


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("mode", ["ManningN", "Q", "zb"])
def test_forward_mode_on_random_meshes_with_symmetry_and_two_inlets(host, tmp_path, seed, mode):
    """Boundary types the reference's fixtures do not contain (symmetry, two inlet-q node strings, corner cells with two boundary
    types, the default-wall rule) on random mixed tri / quad meshes, thin-film states: values and tangents against the oracle."""
    from tests.test_srh_reader_cpu import _write_random_case
    _write_random_case(str(tmp_path), seed)
    c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    flat = R.flatten(c)
    N = c.mesh.numOfCells
    o = Oracle(flat)
    rng = np.random.default_rng(seed)
    Q = cases.random_state_flat(flat, seed + 5, dry_frac=0.08)
    p = {"ManningN": np.asarray(c.ManningN_zone, dtype=np.float64), "Q": np.asarray(flat["inletQ_TotalQ"], dtype=np.float64),
         "zb": np.asarray(c.zb_cells, dtype=np.float64)}[mode].copy()
    V, pdot = rng.standard_normal(3 * N), rng.standard_normal(p.size)
    dQ, dQd, rc = run_host(host, flat, Q, V, p, pdot, ACTIVE[mode])
    assert rc == 0
    ref, ref_d = o.jvp(Q, V, p, pdot, ACTIVE[mode])
    assert (np.abs(dQ - ref) <= 1e-13 * cases.flat_scale(flat, Q)).all()
    assert np.abs(dQd - ref_d).max() <= 1e-12 * np.abs(ref_d).max()
    assert flat["n_symm"] == 1 and flat["n_inletq"] == 2


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "simple", "random_symm"])
def test_product_source_against_the_independent_literal_restatement(host, name, tmp_path):
    """Not twin against twin: hg_jvp_impl.h (the source of the strict path and of the forward-mode kernels) against
    oracle/rhs_literal.py, which was written independently of the C++ oracle in the Julia's own shapes -- values against its
    RHS, tangents against its complex-step derivative (AD-free), all four parameter modes, states with dry cells."""
    from oracle import rhs_literal as LIT
    if name == "random_symm":
        from tests.test_srh_reader_cpu import _write_random_case
        _write_random_case(str(tmp_path), 4)
        c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
        flat = R.flatten(c)
    else:
        c, flat = _flat(name)
    N = c.mesh.numOfCells
    hj = HostJvp(host, flat)
    rng = np.random.default_rng(31)
    e = 1e-30
    Q = cases.random_state_flat(flat, 21, dry_frac=0.08)
    sc = cases.flat_scale(flat, Q)
    for mode, lit_name in (("none", ""), ("zb", "zb"), ("ManningN", "ManningN"), ("Q", "Q")):
        params = {"none": None, "zb": c.zb_cells + 0.02 * rng.standard_normal(N),
                  "ManningN": np.asarray(c.ManningN_zone, dtype=np.float64) * (1 + 0.1 * rng.uniform(-1, 1, c.ManningN_zone.size)),
                  "Q": np.asarray(flat["inletQ_TotalQ"], dtype=np.float64) * 0.9}[mode]
        if mode == "Q" and flat["n_inletq"] == 0:
            continue
        V = rng.standard_normal(3 * N)
        pdot = None if params is None else rng.standard_normal(params.size) * (1.0 if mode == "Q" else 0.01)
        dQ, dQd, rc = hj(Q, V, params, pdot, ACTIVE[mode])
        assert rc == 0
        want = LIT.swe_2d_rhs(c, Q, params, lit_name)
        assert (np.abs(dQ - want) <= 2e-13 * sc).all(), (name, mode)
        want_d = np.imag(LIT.swe_2d_rhs(c, Q + 1j * e * V, None if params is None else params + 1j * e * pdot, lit_name)) / e
        assert np.abs(dQd - want_d).max() <= 5e-12 * np.abs(want_d).max(), (name, mode)
