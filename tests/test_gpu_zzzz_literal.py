"""Device RHS (fused and strict paths) and the hand-written VJP against the INDEPENDENT literal restatement of the reference
(oracle/rhs_literal.py: matrix-form Roe flux, per-boundary ghost vectors -- written without reference to the C++ oracle, and
agreeing with it to 1e-15 on the CPU).  The RHS through its values; the VJP through lambda . (J v) with J v from the complex
step of the restatement, i.e. without any AD on the checker's side.  The shared bodies (tests/literal_checks.py) also run on
the CPU against an oracle-backed stand-in (tests/test_oracle_literal_cpu.py)."""
import pytest

import _pkg
from tests import literal_checks as LC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
@pytest.mark.parametrize("strict", [False, True])
def test_device_rhs_against_the_literal_restatement(hg, name, strict):
    worst = LC.check_rhs(lambda flat: hg.Context(flat, strict=strict, tile_cells=128), name, 1e-13 if strict else 1e-12)
    print(f"{name} {'strict' if strict else 'fused'}: worst {worst:.2e} of the flux scale")


@pytest.mark.parametrize("name", ["simple", "oneD_bump", "savannah"])
def test_device_vjp_against_the_complex_step_of_the_literal_restatement(hg, name):
    worst = LC.check_vjp_identity(lambda flat: hg.Context(flat, tile_cells=128), name, 1e-9)
    print(f"{name}: adjoint identity, worst {worst:.2e}")
