"""The tile tables the C++ builder (hg_host.cpp) emits -- what hg_create uploads and k_fused_rhs walks -- checked on the CPU:
structure against the mesh (canonical face orientation, normals taken from the L cell's own table, slots in the reference's
face order, halo lists, boundary entries) and, with a plain-Python model of the kernel's data path over those tables (stage the
local cells, every face once, per-cell gather through the slots with the slot's sign, sources), the RHS against the oracle.
The arithmetic inside the model is the independent literal restatement (oracle/rhs_literal.py), so what is under test is the
builder's output and the data-path design, not the flux formulas."""
import numpy as np
import pytest

import _pkg
from oracle import rhs_literal as LIT
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases
from tests.test_srh_reader_cpu import _write_random_case


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _walk_tables(c, t, Q, check_structure=True):
    """dQ/dt [3N] in reference order from the tile tables `t` (hg.plan_tables) of case `c` at state Q."""
    m = c.mesh
    N, T, NF, nd = t["N"], t["T"], t["NF"], t["n_desc"]
    perm = t["perm"].astype(np.int64)
    g, hs = c.g, c.h_small
    xi_r, qx_r, qy_r = Q[:N], Q[N:2 * N], Q[2 * N:]
    h_r = xi_r + c.hstill
    h_r = np.where(h_r <= hs, hs, h_r)
    qx_r = np.where(h_r <= hs, 0.0, qx_r)
    qy_r = np.where(h_r <= hs, 0.0, qy_r)
    gh, gqx, gqy = LIT.process_all_boundaries_2d(c, h_r, qx_r, qy_r, c.ManningN_cells, c.zb_cells, c.bc.inletQ_TotalQ, c.bc.exitH_WSE)
    gxi = gh - c.hstill_ghost
    out = np.zeros(3 * N)
    seen_cells = np.zeros(N, dtype=int)
    for tile in range(t["n_tiles"]):
        c0, nc, hp, nh, fp, nf, nfp, _, _, nint, bfp = (int(x) for x in t["tile_desc"][tile * nd:tile * nd + 11])
        ncp = (nc + 1) & ~1
        assert c0 == tile * T and 0 < nc <= T and nint <= nf <= nfp and nfp % 4 == 0
        loc = np.full(ncp + nh, -1, dtype=np.int64)                 # local index -> reference cell id
        loc[:nc] = perm[c0:c0 + nc]
        halo_int = t["halo"][hp:hp + nh].astype(np.int64)
        loc[ncp:] = perm[halo_int]
        if check_structure:
            assert (np.diff(halo_int) > 0).all()                     # ascending, distinct
            assert ((halo_int < c0) | (halo_int >= c0 + nc)).all()   # owned by other tiles
        F = np.zeros((nfp + 1, 3))                                   # flux * len per local face; slot nfp = the zero-flux slot
        touching = [[] for _ in range(nc)]
        for f in range(nf):
            lr = int(t["face_lr"][fp + f])
            lL, lR = lr & 0xFFFF, lr >> 16
            nx, ny, ln = t["face_nx"][fp + f], t["face_ny"][fp + f], t["face_len"][fp + f]
            rL = int(loc[lL])
            L = (xi_r[rL], c.hstill[rL], h_r[rL], qx_r[rL], qy_r[rL], c.zb_cells[rL])
            if f < nint:
                rR = int(loc[lR])
                Rs = (xi_r[rR], c.hstill[rR], h_r[rR], qx_r[rR], qy_r[rR], c.zb_cells[rR])
                if check_structure:
                    assert rL < rR                                   # canonical orientation: L = the smaller reference id
                    assert lL < nc or lR < nc                        # the tile owns at least one side
                    jL = [j for j in range(int(m.cellNodesCount[rL])) if int(m.cellNeighbors[rL][j]) - 1 == rR
                          and not m.bFace_is_boundary[int(m.cellFacesList[rL, j]) - 1]]
                    assert len(jL) >= 1
                    assert any((nx, ny) == tuple(m.cell_normals[rL][j]) and ln == m.face_lengths[int(m.cellFacesList[rL, j]) - 1] for j in jL)
            else:
                e = int(t["bface_e"][bfp + f - nint])
                gid = int(t["bc_ghost"][e])
                Rs = (gxi[gid], c.hstill_ghost[gid], gh[gid], gqx[gid], gqy[gid], c.zb_ghost[gid])
                if check_structure:
                    assert lR == 0xFFFF and lL < nc and int(t["bc_cell_ref"][e]) == rL
                    assert t["bc_hstill"][e] == c.hstill_ghost[gid] and t["bc_zb"][e] == c.zb_ghost[gid]
                    j = [j for j in range(int(m.cellNodesCount[rL])) if int(m.cellNeighbors[rL][j]) - 1 == gid
                         and m.bFace_is_boundary[int(m.cellFacesList[rL, j]) - 1]]
                    assert len(j) == 1 and (nx, ny) == tuple(m.cell_normals[rL][j[0]])
            F[f] = LIT.riemann_2d_roe(*L, *Rs, g, (nx, ny), hs) * ln
            if lL < nc:
                touching[lL].append((f, +1))
            if f < nint and lR < nc:
                touching[lR].append((f, -1))
        for l in range(nc):
            r = int(loc[l])
            seen_cells[r] += 1
            s = np.zeros(3)
            slots = t["cf_idx"][(tile * T + l) * NF:(tile * T + l + 1) * NF]
            used = []
            for j in range(NF):
                ix = int(slots[j])
                f, sg = ix & 0x7FFF, (-1.0 if ix & 0x8000 else 1.0)
                s = s + sg * F[f]                                    # left to right, like the reference's flux_sum
                if check_structure:
                    if j < int(m.cellNodesCount[r]):
                        assert (f, int(sg)) in touching[l]
                        used.append(f)
                        # slot j is the reference's j-th face of this cell: same neighbour (cell or ghost) across it
                        lr = int(t["face_lr"][fp + f])
                        if f < nint:
                            other = int(loc[(lr >> 16) if sg > 0 else (lr & 0xFFFF)])
                            assert other == int(m.cellNeighbors[r][j]) - 1 and not m.bFace_is_boundary[int(m.cellFacesList[r, j]) - 1]
                        else:
                            assert int(t["bc_ghost"][int(t["bface_e"][bfp + f - nint])]) == int(m.cellNeighbors[r][j]) - 1
                    else:
                        assert f == nfp                              # unused slots point at the zero-flux slot
            if check_structure:
                assert sorted(used) == sorted(f for f, _ in touching[l])          # every face of the cell, once
            upd = -s / m.cell_areas[r]
            fx, fy = LIT.compute_friction_terms(h_r[r], qx_r[r], qy_r[r], c.ManningN_cells[r], g, c.k_n, hs)
            wet = 1.0 if h_r[r] > hs else 0.0
            out[r] = upd[0]
            out[N + r] = upd[1] + wet * (g * xi_r[r] * c.S0_cells[r, 0] - fx)
            out[2 * N + r] = upd[2] + wet * (g * xi_r[r] * c.S0_cells[r, 1] - fy)
    assert (seen_cells == 1).all()                                   # every cell owned by exactly one tile
    return out


@pytest.mark.parametrize("name,tile", [("simple", 128), ("oneD_bump", 128), ("savannah", 128), ("savannah", 192), ("savannah", 256),
                                       ("savannah", 384), ("savannah", 512), ("random_symm", 128)])
def test_tile_tables_reproduce_the_oracle_rhs(hg, name, tile, tmp_path, oracle_lib):
    if name == "random_symm":
        _write_random_case(str(tmp_path), 3, ni=15, nj=11)          # 165+ cells: two tiles, symmetry, two inlets
        c = R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    else:
        c = cases.load(name)
    flat = R.flatten(c)
    t = hg.plan_tables(flat, tile_cells=tile)
    assert t["N"] == c.mesh.numOfCells and t["n_tiles"] == (t["N"] + t["T"] - 1) // t["T"]
    o = Oracle(flat)
    for k, Q in enumerate((c.Q0, cases.random_state_flat(flat, 7, dry_frac=0.08))):
        got = _walk_tables(c, t, Q, check_structure=(k == 0))
        want = o.rhs(Q)
        err = float((np.abs(got - want) / cases.flat_scale(flat, Q)).max())
        assert err <= 2e-13, (name, tile, k, err)
