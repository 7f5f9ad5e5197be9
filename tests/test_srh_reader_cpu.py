"""CPU test of the product-side C++ SRH-2D reader / mesh builder against the oracle's numpy restatement of the
reference builder and against the reference's committed truth (zb, S0)."""
import os

import numpy as np
import pytest

import _pkg
from oracle import srh2d_ref as R
from tests import cases

hg = _pkg.load()
from hydrograd_jl_b200 import srh2d  # noqa: E402


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "oneD_uniform", "simple", "oneD_bump_sens"])
def test_cpp_reader_matches_reference_builder(name):
    c = cases.load(name)
    ref = R.flatten(c)
    got = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, name), cases._STEMS[name] + ".srhhydro")
    for k in ("n_cells", "n_faces", "n_ghost", "ld", "index_base", "n_inletq", "n_exith", "n_wall", "n_symm", "n_mat"):
        assert got[k] == ref[k], k
    for k in ("cell_nfaces", "cell_faces", "cell_neighbors", "cell_normals", "face_is_boundary", "face_lengths", "cell_areas",
              "cell_centroids", "bc_ptr", "bc_ghost_ids", "bc_internal_cells", "bc_normals", "bc_lengths", "zb_cells",
              "zb_ghost", "S0_cells", "ManningN_cells", "matID_cells", "inletQ_TotalQ", "exitH_WSE"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k          # bit-exact, ids included
    if name != "oneD_bump_sens":
        t = cases.truth(name)
        assert np.array_equal(got["zb_cells"], t["zb_cell_truth"])
        if "S0_cells_truth" in t.files:
            assert np.array_equal(got["S0_cells"], t["S0_cells_truth"])
    # initial condition set-up like process_ICs_2D.jl
    ic = cases._IC.get(name)
    if ic is not None:
        Q0 = srh2d.setup_initial_condition(got, ic[1][0], ic[1][1], ic[1][2], ic[1][3])
        assert np.array_equal(Q0, c.Q0) and np.array_equal(got["hstill"], c.hstill)
        assert np.array_equal(got["hstill_ghost"], c.hstill_ghost)


def test_reader_errors_are_reported():
    with pytest.raises(hg.HydrogradError):
        srh2d.process_SRH_2D_input("/nonexistent", "x.srhhydro")
    with pytest.raises(ValueError):
        flat = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, "simple"), "simple.srhhydro")
        srh2d.setup_initial_condition(flat, -5.0, 0.5)


def test_reader_is_linear_time(tmp_path):
    """A 90k-cell structured case written on the fly loads in well under a second per 100k cells."""
    import time
    ni, nj = 300, 300
    with open(tmp_path / "big.srhgeom", "w") as f:
        f.write('SRHGEOM 30\nName "big"\nGridUnit "Meters"\n')
        nid = lambda i, j: i * (nj + 1) + j + 1
        e = 1
        for i in range(ni):
            for j in range(nj):
                f.write(f"Elem {e} {nid(i, j)} {nid(i + 1, j)} {nid(i + 1, j + 1)} {nid(i, j + 1)}\n"); e += 1
        for i in range(ni + 1):
            for j in range(nj + 1):
                f.write(f"Node {nid(i, j)} {i * 1.0} {j * 1.0} {0.001 * i}\n")
        f.write("NodeString 1 " + " ".join(str(nid(0, j)) for j in range(nj + 1)) + "\n")
        f.write("NodeString 2 " + " ".join(str(nid(ni, j)) for j in range(nj + 1)) + "\n")
    with open(tmp_path / "big.srhmat", "w") as f:
        f.write("SRHMAT 30\nNMaterials 2\nMatName 1 \"a\"\nMaterial 1 " + " ".join(str(k) for k in range(1, ni * nj + 1)) + "\n")
    with open(tmp_path / "big.srhhydro", "w") as f:
        f.write('SRHHYDRO 30\nCase "big"\nGrid "big.srhgeom"\nHydroMat "big.srhmat"\nManningsN 0 0.03\nManningsN 1 0.03\n'
                "BC 1 INLET-Q\nBC 2 EXIT-H\nIQParams 1 10.0 SI CONVEYANCE\nEWSParamsC 2 1.0 SI C\n")
    t = time.time()
    flat = srh2d.process_SRH_2D_input(str(tmp_path), "big.srhhydro")
    dt = time.time() - t
    assert flat["n_cells"] == ni * nj and flat["n_wall"] == 1 and flat["n_inletq"] == 1 and flat["n_exith"] == 1
    assert dt < 5.0, dt
