"""Stage tables of the host-buffer pipeline (hg_rhs / hg_rhs_vjp with host pointers on >= 1M cells), checked on the CPU against
the mesh adjacency.  Chunk c of a component whose host address needs a shift sh < margin moves the reference rows
[c csz - sh, (c+1) csz - sh) (every copy then starts at a multiple of 256 bytes), so:
  * a tile may run in stage s only if all its cells AND their face neighbours lie below (s+1) csz - margin (or s is the last stage);
  * result chunk c takes rows from c csz - margin on and may leave after stage chunk_done[c] only if every tile owning one of
    those rows has run by then."""
import numpy as np
import pytest

import _pkg


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _needed_stage(rows, K, csz, margin):
    return np.minimum(K - 1, (rows + margin) // csz)


@pytest.mark.parametrize("chunks", [0, 5, 64])
def test_stage_tables_respect_the_shifted_chunk_boundaries(hg, chunks):
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.river(1000, 1050)
    N, ld = int(flat["n_cells"]), int(flat["ld"])
    assert N >= 1 << 20
    plan = hg.plan_pipeline(flat, pipeline_chunks=chunks)
    _, perm = hg.plan_stats(flat, want_perm=True)          # internal -> reference cell id
    K, csz, T, margin = plan["n_chunks"], plan["chunk_cells"], plan["tile_cells"], plan["margin"]
    assert K == (chunks or 8) and margin == 32 and csz % margin == 0 and K * csz >= N and (K - 1) * csz < N
    assert plan["n_tiles"] == (N + T - 1) // T
    tile_of_row = np.empty(N, dtype=np.int64)
    tile_of_row[perm] = np.arange(N) // T                    # tile that owns reference row r
    # rows a tile needs: its own cells and their face neighbours (boundary faces have ghosts, not rows)
    nf = np.asarray(flat["cell_nfaces"])
    neigh = np.asarray(flat["cell_neighbors"]).reshape(ld, N).T - int(flat["index_base"])
    faces = np.asarray(flat["cell_faces"]).reshape(ld, N).T
    isb = np.asarray(flat["face_is_boundary"]).astype(bool)
    need = _needed_stage(np.arange(N), K, csz, margin)       # stage at which a row has landed for every component
    tile_need = np.zeros(plan["n_tiles"], dtype=np.int64)
    np.maximum.at(tile_need, tile_of_row, need)
    for j in range(ld):
        valid = (j < nf)
        fid = np.abs(faces[:, j]) - (1 if int(flat["index_base"]) == 1 else 0)
        interior = valid & ~isb[np.where(valid, fid, 0)]
        rows = np.nonzero(interior)[0]
        np.maximum.at(tile_need, tile_of_row[rows], need[neigh[rows, j]])
    ts = plan["tile_stage"].astype(np.int64)
    assert (ts >= tile_need).all()                            # never before its inputs
    # the only tiles held back are those with inlet-q faces (boundary-wide conveyance sum): they wait for the last chunk
    late = np.nonzero(ts > tile_need)[0]
    ptr = np.asarray(flat["bc_ptr"])
    inlet_cells = np.asarray(flat["bc_internal_cells"])[:ptr[int(flat["n_inletq"])]] - int(flat["index_base"])
    assert set(late) <= set(tile_of_row[inlet_cells]) and (ts[late] == K - 1).all()
    # results: chunk c needs every row in [c csz - margin, (c+1) csz) (to N for the last chunk) computed
    for c in range(K):
        lo, hi = max(0, c * csz - margin), (N if c == K - 1 else min(N, (c + 1) * csz))
        assert plan["chunk_done"][c] == ts[tile_of_row[lo:hi]].max()


def test_chunk_rows_cover_the_vector_at_aligned_host_addresses(hg):
    """hg_debug_chunk_rows = the geometry the copies use: for any host address of a component, the K chunks tile [0, N) without
    gaps, every interior boundary is a multiple of 256 bytes in HOST address, chunk c has landed at least the rows the stage
    tables count on ((c+1) csz - margin) and never reaches beyond the rows result chunk c gathers."""
    import ctypes as C
    lib = hg._lib.load()
    rng = np.random.default_rng(3)
    r0, r1 = C.c_int64(), C.c_int64()
    margin = 32
    for N, K in [(15998186, 30), (1 << 20, 8), (1050000, 8), (1234567, 64), (4096, 4), (100000, 97)]:
        csz = ((N + K - 1) // K + margin - 1) // margin * margin if K > 1 else N
        for addr in [0x7F0000000000, 0x7F0000000008, 0x7F00000000F8] + [int(a) * 8 for a in rng.integers(1 << 20, 1 << 40, 5)]:
            prev = 0
            for c in range(K):
                assert lib.hg_debug_chunk_rows(addr, c, K, csz, N, C.byref(r0), C.byref(r1)) == 0
                a, b = r0.value, r1.value
                assert 0 <= a <= b <= N
                if b > a:
                    assert a == prev
                    prev = b
                    if c > 0:
                        assert (addr + 8 * a) % 256 == 0
                    assert max(0, c * csz - margin) <= a and b <= (N if c == K - 1 else min(N, (c + 1) * csz))
                landed = N if c == K - 1 else max(0, min(N, (c + 1) * csz - margin))
                assert prev >= landed
            assert prev == N
    assert lib.hg_debug_chunk_rows(0, 3, 3, 32, 100, C.byref(r0), C.byref(r1)) != 0      # c out of range


@pytest.mark.parametrize("tile", [128, 256])
def test_tile_builder_accounts_for_every_face_and_halo_cell(hg, tile):
    """hg_plan_stats against the mesh adjacency: with cells renumbered so that tile t owns the internal cells [t T, (t+1) T), every
    interior face is stored once in the tile that owns both its cells and once in EACH of the two tiles it connects otherwise
    (cut faces are evaluated redundantly, no flux exchange); every boundary face once; a tile's halo is the set of distinct
    neighbour cells outside it."""
    from hydrograd_jl_b200 import synthetic as S
    flat, _ = S.river(300, 120)
    N, ld, base = int(flat["n_cells"]), int(flat["ld"]), int(flat["index_base"])
    st, perm = hg.plan_stats(flat, tile_cells=tile, want_perm=True)
    assert sorted(perm.tolist()) == list(range(N))                     # a permutation of the reference ids
    tile_of = np.empty(N, dtype=np.int64)
    tile_of[perm] = np.arange(N) // tile
    assert st["n_tiles"] == (N + tile - 1) // tile and st["max_local"] >= min(tile, N)
    nf = np.asarray(flat["cell_nfaces"])
    neigh = np.asarray(flat["cell_neighbors"]).reshape(ld, N).T - base
    faces = np.abs(np.asarray(flat["cell_faces"]).reshape(ld, N).T) - base
    isb = np.asarray(flat["face_is_boundary"]).astype(bool)
    same = cut = bnd = 0
    halo = set()
    for j in range(ld):
        rows = np.nonzero(j < nf)[0]
        b = isb[faces[rows, j]]
        bnd += int(b.sum())
        r, nb = rows[~b], neigh[rows[~b], j]
        s = tile_of[r] == tile_of[nb]
        same += int(s.sum())                                          # seen from both cells: two per face
        cut += int((~s).sum())                                        # seen from both cells: one per (face, tile)
        halo.update(zip(tile_of[r[~s]].tolist(), nb[~s].tolist()))
    assert st["sum_cell_faces"] == int(nf.sum())
    assert st["interior_tile_faces"] == same // 2 + cut
    assert st["tile_faces"] == same // 2 + cut + bnd
    assert st["halo_cells"] == len(halo)
