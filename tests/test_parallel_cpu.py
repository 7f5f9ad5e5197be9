"""CPU tests of the multi-GPU host logic: RCB partition, rank-local mesh extraction, and the halo exchange over
torch.distributed with the gloo backend (world_size 2).  No CUDA compute is involved: the send blocks are packed
on the host exactly like k_halo_pack does on the device."""
import os

import numpy as np
import pytest

import _pkg

hg = _pkg.load()
from hydrograd_jl_b200 import parallel as P  # noqa: E402
from hydrograd_jl_b200 import synthetic as S  # noqa: E402


def host_pack(loc, info, Q, lam=None):
    """What k_halo_pack writes: block k = [xi | qx | qy | l0 | l1 | l2] of the owned cells on the cut, 6*n_k doubles."""
    n = loc["n_cells"]
    out = np.zeros(6 * sum(info["counts"]))
    o = 0
    e = 0
    for nk in info["counts"]:
        c = info["halo_cells"][e:e + nk]
        for comp in range(3):
            out[o + comp * nk:o + (comp + 1) * nk] = Q[comp * n + c]
            if lam is not None:
                out[o + (3 + comp) * nk:o + (4 + comp) * nk] = lam[comp * n + c]
        o += 6 * nk
        e += nk
    return out


def test_rcb_partition_balanced_and_compact():
    flat, _ = S.dam_break(48)
    N = flat["n_cells"]
    cx, cy = flat["cell_centroids"][:N], flat["cell_centroids"][N:]
    for Pn in (2, 3, 4, 8):
        part = P.rcb_partition(cx, cy, Pn)
        cnt = np.bincount(part, minlength=Pn)
        assert cnt.max() - cnt.min() <= 2
        # compact parts: the cut is O(sqrt(N)) faces, not O(N)
        cut = sum(sum(P.extract_local(flat, part, r)[1]["counts"]) for r in range(Pn))
        assert cut < 16 * np.sqrt(N) * np.log2(Pn + 1)


def test_local_meshes_cover_the_global_one():
    flat, Q0 = S.river(72, 20)
    N = flat["n_cells"]
    part = (np.arange(N) * 3 // N).astype(np.int32)          # slabs along the stream (cells are i-major)
    seen = np.zeros(N, dtype=int)
    nI = nE = 0
    for r in range(3):
        loc, info = P.extract_local(flat, part, r, Q0)
        seen[info["own"]] += 1
        nI += loc["n_inletq"]; nE += loc["n_exith"]
        st = hg.plan_stats(loc, tile_cells=128)                # the C++ builder accepts the local mesh
        assert st["n_tiles"] == -(-loc["n_cells"] // 128)
        assert loc["n_ghost"] == len(loc["bc_internal_cells"]) == loc["bc_ptr"][-1]
        # every cut face appears on both sides with the same canonical order and opposite flip flags
    assert (seen == 1).all() and nI == 1 and nE == 1
    a, ia = P.extract_local(flat, part, 0, Q0)
    b, ib = P.extract_local(flat, part, 1, Q0)
    na = ia["counts"][ia["neighbors"].index(1)]
    assert na == ib["counts"][ib["neighbors"].index(0)]
    fa = a["halo_flip"][-na:] if ia["neighbors"][-1] == 1 else None
    assert fa is not None and (fa == 0).all()                 # rank 0 holds the smaller global ids
    ea = b["halo_flip"][b["bc_ptr"][b["n_exith"] + b["n_wall"] + b["n_inletq"]]:][:na]
    assert (ea == 1).all()
    # the remote cells rank 1 expects are exactly the cells rank 0 packs, in the same order
    assert np.array_equal(ia["own"][ia["halo_cells"][-na:]], ib["halo_remote"][:na])


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat, Q0 = S.dam_break(24)
        N = flat["n_cells"]
        part = P.rcb_partition(flat["cell_centroids"][:N], flat["cell_centroids"][N:], world)
        loc, info = P.extract_local(flat, part, rank, Q0)
        rng = np.random.default_rng(5)
        lam_g = rng.standard_normal(3 * N)
        n = loc["n_cells"]
        lam = np.concatenate([lam_g[c * N + info["own"]] for c in range(3)])
        send = torch.from_numpy(host_pack(loc, info, info["Q"], lam))
        recv = torch.zeros_like(send)
        ex = P.HaloExchanger(send, recv, info["neighbors"], info["counts"])
        ex.exchange(with_lambda=True)
        # what must have arrived: state and cotangent of the remote cells, entry by entry
        got = recv.numpy()
        o = e = 0
        ok = True
        for nk in info["counts"]:
            rc = info["halo_remote"][e:e + nk]
            for comp in range(3):
                ok &= np.array_equal(got[o + comp * nk:o + (comp + 1) * nk], Q0[comp * N + rc])
                ok &= np.array_equal(got[o + (3 + comp) * nk:o + (4 + comp) * nk], lam_g[comp * N + rc])
            o += 6 * nk
            e += nk
        # RHS-only exchange moves only the first half of each block
        recv.zero_()
        ex.exchange(with_lambda=False)
        nk = info["counts"][0]
        ok &= bool((recv.numpy()[3 * nk:6 * nk] == 0).all()) and bool((recv.numpy()[:3 * nk] != 0).any())
        q.put((rank, bool(ok), n))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_partition_keeps_inlets_whole():
    """An RCB cut along the stream splits the river's inlet node-string between ranks; with keep_together the inlet's
    cells move to one rank as a whole (the conveyance sum of bc_2D.jl:665-691 runs over all of its faces), every rank's
    local mesh builds, exactly one rank holds the inlet, and the balance stays within the size of the inlet."""
    flat, Q0 = S.river(40, 48)
    N = flat["n_cells"]
    cx, cy = flat["cell_centroids"][:N], flat["cell_centroids"][N:]
    groups = P.inlet_cell_groups(flat)
    assert len(groups) == 1 and groups[0].size >= 40
    # cut the domain ACROSS the inlet: sort by the cross-stream coordinate of the inlet cells' neighbourhood
    g = groups[0]
    d = np.stack([cx[g] - cx[g].mean(), cy[g] - cy[g].mean()])
    axis = np.linalg.svd(d, full_matrices=False)[0][:, 0]      # direction along the inlet node-string
    key = (cx - cx[g].mean()) * axis[0] + (cy - cy[g].mean()) * axis[1]
    part = (key > np.median(key)).astype(np.int32)
    assert len(np.unique(part[g])) == 2, "the test partition must split the inlet"
    with pytest.raises(NotImplementedError):
        for r in range(2):
            P.extract_local(flat, part, r, Q0)
    fixed = part.copy()
    for grp in groups:
        fixed[grp] = np.bincount(fixed[grp], minlength=2).argmax()
    n_in = 0
    for r in range(2):
        loc, info = P.extract_local(flat, fixed, r, Q0)
        n_in += loc["n_inletq"]
        assert hg.plan_stats(loc, tile_cells=128)["n_tiles"] >= 1
    assert n_in == 1
    # the same fix-up inside rcb_partition
    for Pn in (2, 4, 8):
        pr = P.rcb_partition(cx, cy, Pn, keep_together=groups)
        assert len(np.unique(pr[g])) == 1
        cnt = np.bincount(pr, minlength=Pn)
        assert cnt.max() - cnt.min() <= 2 + g.size
        assert sum(P.extract_local(flat, pr, r, Q0)[0]["n_inletq"] for r in range(Pn)) == 1


def test_cut_slot_is_not_confused_with_a_colliding_ghost_id():
    """Ghost ids (boundary slots) and cell ids (interior slots) of the neighbour table share the range 0..; a boundary
    cell whose ghost id equals the global id of its remote neighbour must still get the halo ghost in the CUT slot and
    keep its physical ghost (round-1 advice: the slot search matched the boundary slot first)."""
    flat, Q0 = S.dam_break(24)
    flat = dict(flat)
    N, ld, base = flat["n_cells"], flat["ld"], flat["index_base"]
    neigh = np.asarray(flat["cell_neighbors"]).copy().reshape(ld, N)
    faces = np.abs(np.asarray(flat["cell_faces"]).reshape(ld, N)) - base
    isb = np.asarray(flat["face_is_boundary"]).astype(bool)
    nf = np.asarray(flat["cell_nfaces"])
    B = flat["n_ghost"]
    # a boundary cell c whose boundary slot comes BEFORE an interior slot holding a neighbour r < B
    pick = None
    for c in range(N):
        js = [j for j in range(nf[c]) if isb[faces[j, c]]]
        ks = [j for j in range(nf[c]) if not isb[faces[j, c]] and neigh[j, c] - base < B]
        if js and ks and js[0] < ks[-1]:
            pick = (c, js[0], ks[-1]); break
    assert pick is not None
    c, jb, ji = pick
    g, r = int(neigh[jb, c]) - base, int(neigh[ji, c]) - base
    # renumber the ghosts: swap ids g and r, so that cell c's ghost id equals its neighbour's cell id
    swap = np.arange(B); swap[g], swap[r] = r, g
    bslots = np.zeros((ld, N), dtype=bool)
    for cc in range(N):
        for j in range(nf[cc]):
            bslots[j, cc] = isb[faces[j, cc]]
    neigh[bslots] = swap[neigh[bslots] - base] + base
    flat["cell_neighbors"] = neigh.ravel()
    flat["bc_ghost_ids"] = swap[np.asarray(flat["bc_ghost_ids"]) - base] + base
    for k in ("hstill_ghost", "zb_ghost"):
        a = np.asarray(flat[k]).copy(); a[[g, r]] = a[[r, g]]; flat[k] = a
    assert neigh[jb, c] == neigh[ji, c]
    part = np.zeros(N, dtype=np.int32); part[r] = 1          # the colliding neighbour lives on the other rank
    loc, info = P.extract_local(flat, part, 0, Q0)
    lc = int(np.nonzero(info["own"] == c)[0][0])
    ln = np.asarray(loc["cell_neighbors"]).reshape(ld, loc["n_cells"])
    n_phys = loc["n_ghost"] - sum(info["counts"])
    assert ln[jb, lc] < n_phys <= ln[ji, lc]                  # physical ghost kept, halo ghost in the cut slot
    assert hg.plan_stats(loc, tile_cells=128)["n_tiles"] >= 1


@pytest.mark.parametrize("case", ["dam", "river_inlet", "river_slab_gid"])
def test_native_partitioner_matches_the_numpy_twin(case):
    """hg_partition_rcb / hg_partition_extract (C++, what a non-Python host calls) against the numpy statements of the same
    rules: identical parts, identical local tables (ids, orders, flip flags, fields), for RCB with kept-together inlets, for
    slabs, and with explicit global ids."""
    if case == "dam":
        flat, Q0 = S.dam_break(40); Pn = 4
    else:
        flat, Q0 = S.river(60, 48); Pn = 3 if case == "river_slab_gid" else 4
    N = flat["n_cells"]
    cx, cy = flat["cell_centroids"][:N], flat["cell_centroids"][N:]
    groups = P.inlet_cell_groups(flat)
    gid = None
    if case == "river_slab_gid":
        part = (np.arange(N) * Pn // N).astype(np.int32)
        gid = np.random.default_rng(1).permutation(N).astype(np.int64) + 5     # any injective map is a valid set of global ids
    else:
        part = P.rcb_partition(cx, cy, Pn, keep_together=groups)
        ref = P.rcb_partition_reference(cx, cy, Pn, keep_together=groups)
        assert np.array_equal(part, ref)
        for Pk in (2, 3, 5, 8):
            assert np.array_equal(P.rcb_partition(cx, cy, Pk), P.rcb_partition_reference(cx, cy, Pk)), Pk
    for r in range(Pn):
        a, ia = P.extract_local(flat, part, r, Q0, gid=gid)
        b, ib = P.extract_local_reference(flat, part, r, Q0, gid=gid)
        for k in b:
            if b[k] is None:
                assert a.get(k) is None or np.size(a[k]) == 0, k
            elif isinstance(b[k], np.ndarray):
                assert np.array_equal(np.asarray(a[k]), b[k]), (k, r)
            else:
                assert a[k] == b[k], (k, r)
        for k in ("own", "halo_remote", "halo_cells", "Q"):
            assert np.array_equal(ia[k], ib[k]), (k, r)
        assert ia["neighbors"] == ib["neighbors"] and ia["counts"] == ib["counts"]
