"""The tile tables the C++ builder (hg_host.cpp) emits -- what hg_create uploads and k_fused_rhs walks -- checked on the CPU:
structure against the mesh (canonical face orientation, normals taken from the L cell's own table, slots in the reference's
face order, halo lists, boundary entries) and, with a plain-Python model of the kernel's data path over those tables (stage the
local cells, every face once, per-cell gather through the slots with the slot's sign, sources), the RHS against the oracle.
The arithmetic inside the model is the independent literal restatement (oracle/rhs_literal.py), so what is under test is the
builder's output and the data-path design, not the flux formulas.  The same walk over the rank-local tables of a partitioned
mesh (C++ partitioner -> tile builder -> halo faces evaluated with the remote cell's state, flipped where the remote cell is
the canonical L side) gives the single-mesh result BIT FOR BIT: the design's rank-count independence, shown without a device."""
import numpy as np
import pytest

import _pkg
from oracle import rhs_literal as LIT
from oracle import srh2d_ref as R
from oracle.oracle import Oracle
from tests import cases
from tests.test_srh_reader_cpu import _write_random_case

BC_INLETQ, BC_EXITH, BC_WALL, BC_SYMM, BC_HALO = range(5)


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


class Mesh:
    """The flat ABI tables of one (global or rank-local) mesh, 0-based."""

    def __init__(self, flat):
        self.flat = flat
        N, ld, base = int(flat["n_cells"]), int(flat["ld"]), int(flat["index_base"])
        self.N, self.B = N, int(flat["n_ghost"])
        self.nf = np.asarray(flat["cell_nfaces"]).astype(np.int64)
        self.neigh = np.asarray(flat["cell_neighbors"]).reshape(ld, N).T - base       # [N, ld]: cell id, or ghost id on boundary faces
        self.face = np.abs(np.asarray(flat["cell_faces"]).reshape(ld, N).T) - base
        nrm = np.asarray(flat["cell_normals"]).reshape(2, ld, N)
        self.nx, self.ny = nrm[0].T, nrm[1].T
        self.isb = np.asarray(flat["face_is_boundary"]).astype(bool)
        self.flen = np.asarray(flat["face_lengths"], dtype=np.float64)
        self.area = np.asarray(flat["cell_areas"], dtype=np.float64)
        self.hstill, self.zb, self.mann = (np.asarray(flat[k], dtype=np.float64) for k in ("hstill", "zb_cells", "ManningN_cells"))
        S0 = np.asarray(flat["S0_cells"], dtype=np.float64)
        self.S0x, self.S0y = S0[:N], S0[N:]
        self.hstill_g, self.zb_g = np.asarray(flat["hstill_ghost"], dtype=np.float64), np.asarray(flat["zb_ghost"], dtype=np.float64)
        self.g, self.k_n, self.hs = float(flat["g"]), float(flat["k_n"]), float(flat["h_small"])
        # boundary entries in processing order (inlet-q, exit-h, wall, symm, then halo boundaries)
        ptr = np.asarray(flat["bc_ptr"], dtype=np.int64)
        counts = [int(flat["n_inletq"]), int(flat["n_exith"]), int(flat["n_wall"]), int(flat["n_symm"]), int(flat.get("n_halo", 0))]
        self.e_type, self.e_group = np.zeros(self.B, dtype=np.int64), np.zeros(self.B, dtype=np.int64)
        kb = 0
        for ty, cnt in enumerate(counts):
            for k in range(cnt):
                self.e_type[ptr[kb]:ptr[kb + 1]] = ty
                self.e_group[ptr[kb]:ptr[kb + 1]] = k
                kb += 1
        self.ptr, self.counts = ptr, counts
        self.e_ghost = np.asarray(flat["bc_ghost_ids"]).astype(np.int64) - base
        self.e_cell = np.asarray(flat["bc_internal_cells"]).astype(np.int64) - base
        bn = np.asarray(flat["bc_normals"], dtype=np.float64)
        self.e_nx, self.e_ny = bn[:self.B], bn[self.B:]
        self.e_len = np.asarray(flat["bc_lengths"], dtype=np.float64) if counts[0] else np.zeros(self.B)
        self.e_flip = np.asarray(flat["halo_flip"]).astype(bool) if counts[4] else np.zeros(self.B, dtype=bool)
        self.Qin, self.wse = np.asarray(flat["inletQ_TotalQ"], dtype=np.float64), np.asarray(flat["exitH_WSE"], dtype=np.float64)


def clamp(mesh, Q):
    """semi_discretize_swe_2D.jl:101-106."""
    N = mesh.N
    xi, qx, qy = Q[:N], Q[N:2 * N], Q[2 * N:]
    h = xi + mesh.hstill
    h = np.where(h <= mesh.hs, mesh.hs, h)
    return xi, h, np.where(h <= mesh.hs, 0.0, qx), np.where(h <= mesh.hs, 0.0, qy)


def ghost_states(mesh, h, qx, qy, remote=None):
    """Ghost (xi, hstill, h, qx, qy, zb) per ghost id: bc_2D.jl:575-875 entry by entry (the literal restatement's logic on the
    flat tables); halo entries take the clamped state of the remote cell, `remote(e)` -> (xi, qx, qy) of that cell."""
    hs = mesh.hs
    gh, gqx, gqy, gxi = (np.zeros(mesh.B) for _ in range(4))
    kb = 0
    for ty, cnt in enumerate(mesh.counts):
        for k in range(cnt):
            es = np.arange(mesh.ptr[kb], mesh.ptr[kb + 1])
            kb += 1
            ic, gid = mesh.e_cell[es], mesh.e_ghost[es]
            if ty == BC_INLETQ:
                L = mesh.e_len[es]
                wet = (h[ic] > hs).astype(np.float64)
                A = 0.0
                for i in range(len(es)):
                    A = A + L[i] ** (5.0 / 3.0) * h[ic[i]] / mesh.mann[ic[i]] * wet[i]
                assert A > 1e-10
                vn = mesh.Qin[k] / A * L ** (2.0 / 3.0) / mesh.mann[ic]
                gh[gid], gqx[gid], gqy[gid] = h[ic], -h[ic] * vn * mesh.e_nx[es] * wet, -h[ic] * vn * mesh.e_ny[es] * wet
            elif ty == BC_EXITH:
                gh[gid], gqx[gid], gqy[gid] = np.maximum(hs, mesh.wse[k] - mesh.zb[ic]), qx[ic], qy[ic]
            elif ty == BC_WALL:
                gh[gid], gqx[gid], gqy[gid] = h[ic], -qx[ic], -qy[ic]
            elif ty == BC_SYMM:
                vdn = qx[ic] * mesh.e_nx[es] + qy[ic] * mesh.e_ny[es]
                gh[gid], gqx[gid], gqy[gid] = h[ic], qx[ic] - 2.0 * vdn * mesh.e_nx[es], qy[ic] - 2.0 * vdn * mesh.e_ny[es]
            else:
                for e, g_ in zip(es, gid):
                    xr, qxr, qyr = remote(int(e))
                    hr = xr + mesh.hstill_g[g_]
                    dry = hr <= hs
                    gh[g_], gqx[g_], gqy[g_], gxi[g_] = (hs if dry else hr), (0.0 if dry else qxr), (0.0 if dry else qyr), xr
            if ty != BC_HALO:
                gxi[gid] = gh[gid] - mesh.hstill_g[gid]                # semi_discretize_swe_2D.jl:220
    return gxi, gh, gqx, gqy


def walk_tables(mesh, t, Q, check_structure=True, remote=None):
    """dQ/dt [3N] in the mesh's own cell order from its tile tables `t` (hg.plan_tables) at state Q [3N]."""
    N, T, NF, nd = t["N"], t["T"], t["NF"], t["n_desc"]
    assert N == mesh.N
    perm = t["perm"].astype(np.int64)
    g, hs = mesh.g, mesh.hs
    xi_r, h_r, qx_r, qy_r = clamp(mesh, Q)
    gxi, gh, gqx, gqy = ghost_states(mesh, h_r, qx_r, qy_r, remote)
    out = np.zeros(3 * N)
    seen_cells = np.zeros(N, dtype=int)
    for tile in range(t["n_tiles"]):
        c0, nc, hp, nh, fp, nf, nfp, _, _, nint, bfp = (int(x) for x in t["tile_desc"][tile * nd:tile * nd + 11])
        ncp = (nc + 1) & ~1
        assert c0 == tile * T and 0 < nc <= T and nint <= nf <= nfp and nfp % 4 == 0
        loc = np.full(ncp + nh, -1, dtype=np.int64)                 # local index -> cell id of the mesh
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
            nx, ny, ln = float(t["face_nx"][fp + f]), float(t["face_ny"][fp + f]), float(t["face_len"][fp + f])
            rL = int(loc[lL])
            L = (xi_r[rL], mesh.hstill[rL], h_r[rL], qx_r[rL], qy_r[rL], mesh.zb[rL])
            if f < nint:
                rR = int(loc[lR])
                Rs = (xi_r[rR], mesh.hstill[rR], h_r[rR], qx_r[rR], qy_r[rR], mesh.zb[rR])
                if check_structure:
                    assert lL < nc or lR < nc                        # the tile owns at least one side
                    jL = [j for j in range(int(mesh.nf[rL])) if int(mesh.neigh[rL, j]) == rR and not mesh.isb[mesh.face[rL, j]]]
                    assert any((nx, ny, ln) == (mesh.nx[rL, j], mesh.ny[rL, j], mesh.flen[mesh.face[rL, j]]) for j in jL)
            else:
                e = int(t["bface_e"][bfp + f - nint])
                gid = int(t["bc_ghost"][e])
                Rs = (gxi[gid], mesh.hstill_g[gid], gh[gid], gqx[gid], gqy[gid], mesh.zb_g[gid])
                if check_structure:
                    assert lR == 0xFFFF and lL < nc and int(t["bc_cell_ref"][e]) == rL and int(t["bc_type"][e]) == mesh.e_type[e]
                    assert t["bc_hstill"][e] == mesh.hstill_g[gid] and t["bc_zb"][e] == mesh.zb_g[gid] and gid == mesh.e_ghost[e]
                    j = [j for j in range(int(mesh.nf[rL])) if int(mesh.neigh[rL, j]) == gid and mesh.isb[mesh.face[rL, j]]]
                    assert len(j) == 1 and (nx, ny) == (mesh.nx[rL, j[0]], mesh.ny[rL, j[0]])
                if mesh.e_type[e] == BC_HALO and mesh.e_flip[e]:
                    # the remote cell is the canonical L side: evaluate in ITS orientation (its outward normal = -n) and hand
                    # the owned cell the opposite flux -- the same call the owning rank of that cell makes (hg_fused.cu phase 2)
                    L, Rs = Rs, L
                    nx, ny, ln = -nx, -ny, -ln
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
                    if j < int(mesh.nf[r]):
                        assert (f, int(sg)) in touching[l]
                        used.append(f)
                        # slot j is the reference's j-th face of this cell: same neighbour (cell or ghost) across it
                        if f < nint:
                            lr = int(t["face_lr"][fp + f])
                            other = int(loc[(lr >> 16) if sg > 0 else (lr & 0xFFFF)])
                            assert other == int(mesh.neigh[r, j]) and not mesh.isb[mesh.face[r, j]]
                        else:
                            assert int(t["bc_ghost"][int(t["bface_e"][bfp + f - nint])]) == int(mesh.neigh[r, j]) and mesh.isb[mesh.face[r, j]]
                    else:
                        assert f == nfp                              # unused slots point at the zero-flux slot
            if check_structure:
                assert sorted(used) == sorted(f for f, _ in touching[l])          # every face of the cell, once
            upd = -s / mesh.area[r]
            fx, fy = LIT.compute_friction_terms(h_r[r], qx_r[r], qy_r[r], mesh.mann[r], g, mesh.k_n, hs)
            wet = 1.0 if h_r[r] > hs else 0.0
            out[r] = upd[0]
            out[N + r] = upd[1] + wet * (g * xi_r[r] * mesh.S0x[r] - fx)
            out[2 * N + r] = upd[2] + wet * (g * xi_r[r] * mesh.S0y[r] - fy)
    assert (seen_cells == 1).all()                                   # every cell owned by exactly one tile
    return out


def _case(name, tmp_path):
    if name == "random_symm":
        _write_random_case(str(tmp_path), 3, ni=15, nj=11)          # 165+ cells: two tiles, symmetry, two inlets
        return R.load_case(str(tmp_path), "rnd.srhhydro", ("constant", [3.0, 2.0, 0.1, 0.0]))
    return cases.load(name)


@pytest.mark.parametrize("name,tile", [("simple", 128), ("oneD_bump", 128), ("savannah", 128), ("savannah", 192), ("savannah", 256),
                                       ("savannah", 384), ("savannah", 512), ("random_symm", 128)])
def test_tile_tables_reproduce_the_oracle_rhs(hg, name, tile, tmp_path, oracle_lib):
    c = _case(name, tmp_path)
    flat = R.flatten(c)
    mesh = Mesh(flat)
    t = hg.plan_tables(flat, tile_cells=tile)
    assert t["N"] == mesh.N and t["n_tiles"] == (t["N"] + t["T"] - 1) // t["T"]
    if name != "random_symm":
        # canonical orientation of the interior faces: L = the smaller reference id
        nd, perm = t["n_desc"], t["perm"]
        for tile_i in range(t["n_tiles"]):
            c0, nc, hp, nh, fp, nf, nfp, _, _, nint, bfp = (int(x) for x in t["tile_desc"][tile_i * nd:tile_i * nd + 11])
            ncp = (nc + 1) & ~1
            loc = np.concatenate([perm[c0:c0 + nc], np.full(ncp - nc, -1), perm[t["halo"][hp:hp + nh]]])
            lr = t["face_lr"][fp:fp + nint].astype(np.int64)
            assert (loc[lr & 0xFFFF] < loc[lr >> 16]).all()
    o = Oracle(flat)
    for k, Q in enumerate((c.Q0, cases.random_state_flat(flat, 7, dry_frac=0.08))):
        got = walk_tables(mesh, t, Q, check_structure=(k == 0))
        want = o.rhs(Q)
        err = float((np.abs(got - want) / cases.flat_scale(flat, Q)).max())
        assert err <= 2e-13, (name, tile, k, err)


@pytest.mark.parametrize("name,P", [("savannah", 3), ("savannah", 4), ("oneD_bump", 2)])
def test_partitioned_tile_tables_give_the_single_mesh_bits(hg, name, P, oracle_lib):
    """C++ partitioner -> rank-local meshes with halo boundaries -> tile builder -> the walk, every halo face fed with the remote
    cell's state: the assembled result equals the single-mesh walk BIT FOR BIT (cut faces are evaluated on both ranks with the
    single-mesh orientation), and the oracle to rounding."""
    from hydrograd_jl_b200 import parallel as PAR
    c = cases.load(name)
    flat = R.flatten(c)
    N = int(flat["n_cells"])
    cen = np.asarray(flat["cell_centroids"])
    part = PAR.rcb_partition(cen[:N], cen[N:], P, keep_together=PAR.inlet_cell_groups(flat))
    assert len(set(part.tolist())) == P
    Q = cases.random_state_flat(flat, 9, dry_frac=0.08)
    single = walk_tables(Mesh(flat), hg.plan_tables(flat, tile_cells=128), Q, check_structure=False)
    out = np.full(3 * N, np.nan)
    n_halo_faces = 0
    for rank in range(P):
        loc, info = PAR.extract_local(flat, part, rank, Q)
        mesh = Mesh(loc)
        own, rem = info["own"], info["halo_remote"]
        n_phys = mesh.B - rem.size
        n_halo_faces += rem.size
        remote = lambda e: (Q[rem[e - n_phys]], Q[N + rem[e - n_phys]], Q[2 * N + rem[e - n_phys]])
        got = walk_tables(mesh, hg.plan_tables(loc, tile_cells=128), info["Q"], check_structure=True, remote=remote)
        n = own.size
        out[own], out[N + own], out[2 * N + own] = got[:n], got[n:2 * n], got[2 * n:]
    assert n_halo_faces > 0 and not np.isnan(out).any()
    assert np.array_equal(out, single)                                # rank-count independent to the bit
    want = Oracle(flat).rhs(Q)
    assert (np.abs(out - want) <= 2e-13 * cases.flat_scale(flat, Q)).all()


def test_launch_orders_of_a_rank_local_mesh(hg):
    """band_order (two-phase overlap): the tiles without halo faces first, then the band; comm_order (library-owned transport): the
    band in the middle of the launch, interior tiles on both sides; tile_order (host-buffer pipeline): a permutation."""
    from hydrograd_jl_b200 import parallel as PAR
    c = cases.load("savannah")
    flat = R.flatten(c)
    N = int(flat["n_cells"])
    cen = np.asarray(flat["cell_centroids"])
    part = PAR.rcb_partition(cen[:N], cen[N:], 4, keep_together=PAR.inlet_cell_groups(flat))
    bands = 0
    for rank in range(4):
        loc, _ = PAR.extract_local(flat, part, rank)
        t = hg.plan_tables(loc, tile_cells=128)
        nd, nt = t["n_desc"], t["n_tiles"]
        band = set()
        for tile in range(nt):
            nf, nint, bfp = (int(t["tile_desc"][tile * nd + k]) for k in (5, 9, 10))
            if any(int(t["bc_type"][int(e)]) == BC_HALO for e in t["bface_e"][bfp:bfp + nf - nint]):
                band.add(tile)
        bands += len(band)
        ni, b0 = t["n_interior_tiles"], t["comm_band0"]
        assert ni == nt - len(band)
        for order in (t["band_order"], t["comm_order"], t["tile_order"]):
            assert sorted(order.tolist()) == list(range(nt))
        assert set(t["band_order"][ni:].tolist()) == band
        assert set(t["comm_order"][b0:b0 + len(band)].tolist()) == band and b0 == ni // 2
    assert bands > 0


@pytest.mark.parametrize("which", ["river", "dam_thin"])
def test_synthetic_meshes_tiles_and_partitions(hg, which, oracle_lib):
    """The benchmark mesh families at test size (mixed triangles / quadrilaterals; river: inlet-q, exit-h, walls, six Manning zones;
    dam break with a thin film: wet/dry fronts): three tile shapes against the oracle, 2 and 5 ranks bit-identical to one."""
    from hydrograd_jl_b200 import parallel as PAR
    from hydrograd_jl_b200 import synthetic as S
    flat, Q0 = S.river(60, 24) if which == "river" else S.dam_break(36, thin_film=True)
    N = int(flat["n_cells"])
    mesh, o = Mesh(flat), Oracle(flat)
    for tile in (128, 224, 256):
        t = hg.plan_tables(flat, tile_cells=tile)
        for k, Q in enumerate((Q0, cases.random_state_flat(flat, 5, dry_frac=0.1))):
            got = walk_tables(mesh, t, Q, check_structure=(k == 0 and tile == 128))
            assert (np.abs(got - o.rhs(Q)) <= 2e-13 * cases.flat_scale(flat, Q)).all(), (which, tile, k)
    cen = np.asarray(flat["cell_centroids"])
    Q = cases.random_state_flat(flat, 6, dry_frac=0.1)
    single = walk_tables(mesh, hg.plan_tables(flat, tile_cells=128), Q, check_structure=False)
    for P in (2, 5):
        part = PAR.rcb_partition(cen[:N], cen[N:], P, keep_together=PAR.inlet_cell_groups(flat))
        out = np.full(3 * N, np.nan)
        for rank in range(P):
            loc, info = PAR.extract_local(flat, part, rank, Q)
            m = Mesh(loc)
            own, rem = info["own"], info["halo_remote"]
            n_phys = m.B - rem.size
            remote = lambda e: (Q[rem[e - n_phys]], Q[N + rem[e - n_phys]], Q[2 * N + rem[e - n_phys]])
            got = walk_tables(m, hg.plan_tables(loc, tile_cells=128), info["Q"], check_structure=True, remote=remote)
            n = own.size
            out[own], out[N + own], out[2 * N + own] = got[:n], got[n:2 * n], got[2 * n:]
        assert np.array_equal(out, single), (which, P)


@pytest.mark.parametrize("name", ["savannah", "oneD_bump"])
def test_adjoint_side_tables(hg, name):
    """What the VJP's follow-up kernels index with: the boundary entries grouped by their owned cell (deterministic scatter of the
    boundary adjoints) and, per cell-face of the reference-order CSR, the position of the same face in the neighbour's list
    (transposed Green-Gauss for the bed gradient)."""
    flat = R.flatten(cases.load(name))
    mesh = Mesh(flat)
    t = hg.plan_tables(flat, tile_cells=128)
    ref, ptr, ent = t["bcell_ref"], t["bcell_ptr"], t["bcell_ent"]
    assert (np.diff(ref) > 0).all() and ptr[0] == 0 and ptr[-1] == mesh.B and ptr.size == ref.size + 1
    assert sorted(ent.tolist()) == list(range(mesh.B))
    for k in range(ref.size):
        es = ent[ptr[k]:ptr[k + 1]]
        assert (t["bc_cell_ref"][es] == ref[k]).all() and (np.diff(es) > 0).all()
    cf_ptr = np.concatenate([[0], np.cumsum(mesh.nf)])
    rev = t["cf_rev"]
    assert rev.size == cf_ptr[-1]
    for i in range(mesh.N):
        for j in range(int(mesh.nf[i])):
            k = cf_ptr[i] + j
            if mesh.isb[mesh.face[i, j]]:
                assert rev[k] == -1
            else:
                nb = int(mesh.neigh[i, j])
                jj = int(rev[k]) - cf_ptr[nb]
                assert 0 <= jj < mesh.nf[nb] and mesh.neigh[nb, jj] == i and mesh.face[nb, jj] == mesh.face[i, j]
                assert rev[rev[k]] == k


def test_shared_memory_bank_statistics_of_the_tables(hg):
    """A performance property of the tables, computed from the tables: a half-warp's 64-bit shared-memory gather takes one
    wavefront per distinct address in the most loaded of the 16 eight-byte banks.  Phase 2 (16 consecutive faces read their L and
    their R cells): the builder's bank matching keeps it near 1.1; phase 3 (16 consecutive cells read their j-th face): ~1.9,
    the known remaining conflict.  Guards both against regressions of the builder."""
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.river(150, 100)
    t = hg.plan_tables(flat, tile_cells=256)
    nd, T, NF = t["n_desc"], t["T"], t["NF"]

    def wavefronts(idx):
        return np.bincount(np.unique(idx) % 16, minlength=16).max()

    w2 = n2 = w3 = n3 = 0
    for tile in range(t["n_tiles"]):
        c0, nc, hp, nh, fp, nf, nfp, _, _, nint, bfp = (int(x) for x in t["tile_desc"][tile * nd:tile * nd + 11])
        lr = t["face_lr"][fp:fp + nint].astype(np.int64)
        for b in range(0, nint, 16):
            w2 += wavefronts(lr[b:b + 16] & 0xFFFF) + wavefronts(lr[b:b + 16] >> 16)
            n2 += 2
        cf = t["cf_idx"][tile * T * NF:(tile * T + nc) * NF].astype(np.int64).reshape(nc, NF) & 0x7FFF
        for b in range(0, nc, 16):
            for j in range(NF):
                w3 += wavefronts(cf[b:b + 16, j])
                n3 += 1
    print(f"wavefronts per half-warp gather: phase 2 {w2 / n2:.3f}, phase 3 {w3 / n3:.3f}")
    assert w2 / n2 <= 1.2 and w3 / n3 <= 2.1
