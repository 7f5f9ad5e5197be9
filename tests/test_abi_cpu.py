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
