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


def _write_random_case(d, seed, ni=9, nj=7):
    """A jittered ni x nj mesh of quadrilaterals, a third of them split into triangles, three material zones (one cell left to
    the default zone 0), two inlet-q node strings on the left side, exit-h on the right, symmetry at the bottom, an explicit wall
    on part of the top (the rest falls to the default-wall rule, mesh_2D.jl:172-181) and a MONITORING line inside the domain."""
    rng = np.random.default_rng(seed)
    nid = lambda i, j: i * (nj + 1) + j + 1
    xy = {}
    for i in range(ni + 1):
        for j in range(nj + 1):
            jit = 0.0 if i in (0, ni) or j in (0, nj) else 0.25
            xy[nid(i, j)] = (2.0 * i + jit * rng.uniform(-1, 1), 1.5 * j + jit * rng.uniform(-1, 1), 0.05 * i + 0.3 * rng.random())
    elems = []
    for i in range(ni):
        for j in range(nj):
            a, b, c_, e = nid(i, j), nid(i + 1, j), nid(i + 1, j + 1), nid(i, j + 1)
            if rng.random() < 0.33:
                elems += [(a, b, c_), (a, c_, e)] if rng.random() < 0.5 else [(a, b, e), (b, c_, e)]
            else:
                elems.append((a, b, c_, e))
    with open(os.path.join(d, "rnd.srhgeom"), "w") as f:
        f.write('SRHGEOM 30\nName "random"\n\nGridUnit "Meters" \n')
        for k, el in enumerate(elems, 1):
            f.write(f"Elem {k} " + " ".join(map(str, el)) + "\n")
        for k in sorted(xy):
            f.write("Node %d %.17g %.17g %.17g\n" % ((k,) + xy[k]))
        half = nj // 2
        f.write("NodeString 1 " + " ".join(str(nid(0, j)) for j in range(0, half + 1)) + "\n")
        f.write("NodeString 2 " + " ".join(str(nid(0, j)) for j in range(half, nj + 1)) + "\n")
        f.write("NodeString 3 " + " ".join(str(nid(ni, j)) for j in range(nj + 1)) + "\n")
        f.write("NodeString 4 " + " ".join(str(nid(i, 0)) for i in range(ni + 1)) + "\n")
        f.write("NodeString 5 " + " ".join(str(nid(i, nj)) for i in range(2, ni - 1)) + "\n")
        f.write("NodeString 6 " + " ".join(str(nid(3, j)) for j in range(1, nj)) + "\n")
    ne = len(elems)
    zone = rng.integers(1, 4, ne)
    zone[ne // 2] = 0                                                   # not listed in any Material block -> default zone
    with open(os.path.join(d, "rnd.srhmat"), "w") as f:
        f.write("SRHMAT 30\nNMaterials 4\n")
        for z in (1, 2, 3):
            f.write(f'MatName {z} "zone{z}" \n')
        for z in (1, 2, 3):
            ids = [str(k + 1) for k in range(ne) if zone[k] == z]
            f.write(f"Material {z}  " + " ".join(ids[:10]) + "\n")
            for o in range(10, len(ids), 10):                           # continuation lines, as SMS writes them
                f.write(" " + " ".join(ids[o:o + 10]) + " \n")
    with open(os.path.join(d, "rnd.srhhydro"), "w") as f:
        f.write('SRHHYDRO 30\nCase "rnd"\nDescription "fuzz"\nRunType FLOW\nSimTime 0.0 0.02 0.1\nGrid "rnd.srhgeom"\nHydroMat "rnd.srhmat"\n'
                "ManningsN 0 0.035\nManningsN 1 0.02\nManningsN 2 0.045\nManningsN 3 0.03\n"
                "BC 1 INLET-Q\nBC 2 INLET-Q\nBC 3 EXIT-H\nBC 4 SYMM\nBC 5 WALL\nBC 6 MONITORING\n"
                "IQParams 1 0.7 SI CONVEYANCE\nIQParams 2 1.3 SI CONVEYANCE\nEWSParamsC 3 1.25 SI C\n")
    return ne


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cpp_reader_matches_reference_builder_on_random_meshes(tmp_path, seed):
    """The linear-time C++ reader / mesh / boundary / bed builder against the reference-shaped builder (oracle/srh2d_ref.py, a
    restatement of the reference's O(N B) Julia code) on random mixed meshes with every boundary type, several inlets, the
    default-wall rule, a default material zone and multi-line Material blocks: every table to the bit."""
    ne = _write_random_case(str(tmp_path), seed)
    c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    ref = R.flatten(c)
    got = srh2d.process_SRH_2D_input(str(tmp_path), "rnd.srhhydro")
    assert got["n_cells"] == ne and (got["n_inletq"], got["n_exith"], got["n_symm"]) == (2, 1, 1) and got["n_wall"] >= 2 and got["n_mat"] == 4
    for k in ("n_cells", "n_faces", "n_ghost", "ld", "index_base", "n_inletq", "n_exith", "n_wall", "n_symm", "n_mat"):
        assert got[k] == ref[k], k
    # ghost cells are numbered differently on purpose (ascending boundary-face id here, Dict order in Julia / insertion order in
    # the restatement: hg_srh.cpp header); everything that does not carry a ghost id must agree to the bit
    for k in ("cell_nfaces", "cell_faces", "cell_normals", "face_is_boundary", "face_lengths", "cell_areas", "cell_centroids", "bc_ptr",
              "bc_internal_cells", "bc_normals", "bc_lengths", "zb_cells", "S0_cells", "ManningN_cells", "matID_cells",
              "inletQ_TotalQ", "exitH_WSE"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k
    N, ld, B = got["n_cells"], got["ld"], got["n_ghost"]
    fb = np.asarray(got["face_is_boundary"]).astype(bool)
    faces = np.abs(np.asarray(got["cell_faces"]).reshape(ld, N)) - 1
    valid = np.arange(ld)[:, None] < np.asarray(got["cell_nfaces"])[None, :]
    interior = valid & ~fb[np.where(valid, faces, 0)]
    a, b = np.asarray(got["cell_neighbors"]).reshape(ld, N), np.asarray(ref["cell_neighbors"]).reshape(ld, N)
    assert np.array_equal(a[interior], b[interior])
    # ghost ids: a permutation of 1..B on both sides, each tied to the same internal cell; ghost bed = the internal cell's
    for t in (got, ref):
        g = np.asarray(t["bc_ghost_ids"])
        assert sorted(g) == list(range(1, B + 1))
        assert np.array_equal(np.asarray(t["zb_ghost"])[g - 1], np.asarray(t["zb_cells"])[np.asarray(t["bc_internal_cells"]) - 1])
    Q0 = srh2d.setup_initial_condition(got, 3.0, 2.0, 0.1, 0.0)
    assert np.array_equal(Q0, c.Q0) and np.array_equal(got["hstill"], c.hstill)


def test_case_folded_file_names_resolve():
    """The reference's own control file names `Savana_SI.srhhydro` for the file `savana_SI.srhhydro`
    (examples/SWE_2D/forward_simulation/Savannah_River_ManningN_ks_h_Umag/run_control.json:5): the reader falls back to
    the unique case-folded match of the directory, for the .srhhydro itself and for the files it names."""
    d = os.path.join(cases.GOLD, "savannah")
    a = srh2d.process_SRH_2D_input(d, "savana_SI.srhhydro")
    b = srh2d.process_SRH_2D_input(d, "Savana_SI.SRHhydro")
    for k in ("cell_areas", "zb_cells", "cellFacesList"):
        if k in a and k in b:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]))
    with pytest.raises(Exception):
        srh2d.process_SRH_2D_input(d, "no_such_case.srhhydro")
