"""CPU tests of the boundary: the shared library loads, exports every symbol the header declares, and refuses
to compute without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hg():
    mod = _pkg.load()
    if not os.path.exists(mod._lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return mod


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "hydrograd_b200.h")).read()
    return set(re.findall(r"^HG_API\s+[\w\s\*]+?\b(hg_\w+)\s*\(", src, flags=re.M))


def test_library_exports_every_declared_symbol(hg):
    lib = hg._lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    assert declared == set(hg._lib.SYMBOLS), "ctypes table and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", hg._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hg_\w+)", nm))
    assert declared <= exported
    assert lib.hg_abi_version() == 3


def test_no_cpu_fallback(hg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.dam_break(8)
    with pytest.raises(hg.HydrogradError) as e:
        hg.Context(flat)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hydrograd.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_synthetic_meshes_are_consistent():
    from oracle.oracle import Oracle
    hg = _pkg.load()
    from hydrograd_jl_b200 import synthetic as S
    for flat, Q0 in (S.dam_break(24), S.river(48, 20)):
        N, ld = flat["n_cells"], flat["ld"]
        nrm = flat["cell_normals"].reshape(2, ld, N)
        fl = flat["face_lengths"][flat["cell_faces"].reshape(ld, N)]
        valid = np.arange(ld)[:, None] < flat["cell_nfaces"][None, :]
        assert np.abs((nrm * fl[None] * valid[None]).sum(1)).max() < 1e-11      # closed cells
        assert (flat["cell_areas"] > 0).all()
        assert set(np.unique(flat["cell_nfaces"])) == {3, 4}                   # mixed tri/quad
        o = Oracle(flat)
        assert np.isfinite(o.rhs(Q0)).all()
    flat, _ = S.dam_break(24)
    N = flat["n_cells"]
    rest = np.concatenate([np.full(N, 0.7), np.zeros(2 * N)])
    assert np.abs(Oracle(flat).rhs(rest)).max() < 1e-13                          # lake at rest


def test_fastpow_of_the_library_is_the_restated_one(hg):
    """hg_fastpow (the PI controller's power when hg_set_controller_pow(ctx, 1)) against tests/tsit5_ref.fastpow, the
    restatement of DiffEqBase.fastpow that reproduces the reference's saved trajectories: same Float32 operations, so the
    results agree to the last bit wherever glibc's exp2f and numpy's Float32 exp2 do -- they differ by one Float32 ulp in about a
    fifth of the cases (neither is Julia's exp2 either: that last-ulp noise is what lets the restated step sequences drift away
    from the reference's after ~100 steps)."""
    from tests import tsit5_ref as T
    lib = hg._lib.load()
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(np.log(1e-6), np.log(50.0), 4000)), [1e-4, 0.1875, 0.25, 0.5, 1.0, 1.5, 2.0]])
    exact, worst = 0, 0.0
    for x in xs:
        for y in (7.0 / 50.0, 2.0 / 25.0):
            a, b = lib.hg_fastpow(float(x), y), T.fastpow(float(x), y)
            exact += a == b
            worst = max(worst, abs(a - b) / b)
            assert abs(a / x ** y - 1.0) < 2e-4                    # and it IS an approximation of the power
    assert worst <= 1.3e-7 and exact >= 0.6 * 2 * xs.size
    assert lib.hg_fastpow(0.0, 0.14) == 0.0
