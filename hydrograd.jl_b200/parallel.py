"""Multi-GPU domain decomposition: one process / one context per GPU (the reference itself is a single
serial process, so nothing here has a counterpart in Hydrograd.jl).

  rcb_partition   recursive coordinate bisection of the cell centroids into P parts (north_star)
  extract_local   the rank-local mesh in the flat ABI layout: owned cells + one layer of remote cells that
                  appear as the ghost cells of "halo boundaries" (one per neighbouring rank).  Cut faces are
                  evaluated redundantly on both ranks with the SAME canonical orientation as a single-GPU run
                  (L = smaller global id), so a P-rank result is bit-identical to the 1-rank result.
  HaloExchanger   per RHS: pack -> grouped send/recv with each neighbour (torch.distributed: NCCL over NVLink on
                  GPUs, gloo in the CPU tests) -> the kernel reads the received block.  Payloads are 24 B
                  (RHS) or 48 B (VJP: + cotangent) per cut face.
"""
from __future__ import annotations

import numpy as np


def inlet_cell_groups(flat):
    """Internal cells of every inlet-q boundary (0-based ids), one array per inlet: the groups a partition must keep on
    one rank, because the conveyance-weighted discharge split sums over ALL faces of the inlet (bc_2D.jl:665-691)."""
    base = int(flat["index_base"])
    bc_ptr = np.asarray(flat["bc_ptr"], dtype=np.int64)
    b_ic = np.asarray(flat["bc_internal_cells"], dtype=np.int64) - base
    return [np.unique(b_ic[bc_ptr[k]:bc_ptr[k + 1]]) for k in range(int(flat["n_inletq"]))]


def rcb_partition(cx, cy, P, keep_together=None):
    """Recursive coordinate bisection (hg_partition_rcb, csrc/hg_partition.cpp): split the longer extent at the weighted
    median until P parts exist; `keep_together` = list of cell-id arrays (e.g. inlet_cell_groups(flat)) that are moved as a
    whole to the rank owning most of them."""
    import ctypes as C
    from . import _lib as L
    lib = L.load()
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    cy = np.ascontiguousarray(cy, dtype=np.float64)
    groups = [np.asarray(g, dtype=np.int64) for g in (keep_together or [])]
    gptr = np.concatenate([[0], np.cumsum([g.size for g in groups])]).astype(np.int64)
    gcells = np.ascontiguousarray(np.concatenate(groups) if groups else np.zeros(1, dtype=np.int64), dtype=np.int64)
    part = np.zeros(cx.size, dtype=np.int32)
    rc = lib.hg_partition_rcb(cx.size, cx.ctypes.data_as(L.c_f64p), cy.ctypes.data_as(L.c_f64p), int(P), len(groups),
                              gptr.ctypes.data_as(L.c_i64p), gcells.ctypes.data_as(L.c_i64p), part.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise ValueError(f"hg_partition_rcb failed ({rc})")
    return part


def rcb_partition_reference(cx, cy, P, keep_together=None):
    """numpy twin of hg_partition_rcb (tests compare the two): split the longer extent at the weighted median until P parts exist.

    `keep_together`: list of cell-id arrays (e.g. inlet_cell_groups(flat)); after the bisection every group is moved as a
    whole to the rank that already owns most of it (ties: the lowest rank).  An inlet is a few hundred cells, so the load
    balance is unaffected, and because the per-rank results do not depend on the partition (redundant cut faces, canonical
    orientation) neither are the bits.  This is how split inlet-q boundaries are avoided instead of all-reduced."""
    N = cx.size
    part = np.zeros(N, dtype=np.int32)

    def rec(idx, p0, p):
        if p == 1:
            part[idx] = p0
            return
        x, y = cx[idx], cy[idx]
        key = x if (x.max() - x.min()) >= (y.max() - y.min()) else y
        pl = p // 2
        k = int(round(idx.size * pl / p))
        order = np.argsort(key, kind="stable")
        rec(idx[order[:k]], p0, pl)
        rec(idx[order[k:]], p0 + pl, p - pl)

    rec(np.arange(N), 0, P)
    for g in keep_together or []:
        g = np.asarray(g, dtype=np.int64)
        if g.size:
            part[g] = np.bincount(part[g], minlength=P).argmax()
    return part


def _tables(flat):
    N, ld, base = int(flat["n_cells"]), int(flat["ld"]), int(flat["index_base"])
    nf = np.asarray(flat["cell_nfaces"], dtype=np.int64)
    faces = np.abs(np.asarray(flat["cell_faces"], dtype=np.int64).reshape(ld, N).T) - base
    neigh = np.asarray(flat["cell_neighbors"], dtype=np.int64).reshape(ld, N).T - base
    normals = np.asarray(flat["cell_normals"]).reshape(2, ld, N).transpose(2, 1, 0)
    valid = np.arange(ld)[None, :] < nf[:, None]
    return N, ld, base, nf, faces, neigh, normals, valid


def extract_local(flat, part, rank, Q=None, gid=None):
    """Rank-local flat mesh (index_base 0) through hg_partition_extract (csrc/hg_partition.cpp).  `gid` are the global cell ids
    of `flat`'s cells (default: identity); they only define the canonical face orientation / ordering and must be consistent
    across ranks.  Returns (local_flat, info); info = dict(own=global-row indices of the owned cells, neighbors=[rank...],
    counts=[entries per neighbour], halo_cells, halo_remote, Q=local state)."""
    import ctypes as C
    from . import _lib as L
    from .api import _descs
    lib = L.load()
    mesh, bc, fields, keep = _descs(flat)
    part32 = np.ascontiguousarray(part, dtype=np.int32)
    g = None if gid is None else np.ascontiguousarray(gid, dtype=np.int64)
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.hg_partition_extract(C.byref(h), C.byref(mesh), C.byref(bc), C.byref(fields), part32.ctypes.data_as(C.POINTER(C.c_int32)),
                                  int(rank), None if g is None else g.ctypes.data_as(L.c_i64p), err, 512)
    if rc:
        msg = err.value.decode()
        raise (NotImplementedError if "split across ranks" in msg else ValueError)(msg)
    try:
        dims = np.zeros(16, dtype=np.int64)
        lib.hg_case_dims(h, dims.ctypes.data_as(L.c_i64p))
        dts = {0: (np.float64, C.c_double), 1: (np.int64, C.c_int64), 2: (np.uint8, C.c_uint8)}

        def arr(name, default_dtype=np.float64):
            ptr, cnt, dt = C.c_void_p(), C.c_int64(0), C.c_int32(0)
            if lib.hg_case_array(h, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(dt)):
                return None
            npt, ct = dts[dt.value]
            if cnt.value == 0:
                return np.zeros(0, dtype=npt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(cnt.value,)).astype(npt, copy=True)

        loc = dict(n_cells=int(dims[0]), n_faces=int(dims[1]), n_ghost=int(dims[2]), ld=int(dims[3]), index_base=0,
                   n_inletq=int(dims[5]), n_exith=int(dims[6]), n_wall=int(dims[7]), n_symm=int(dims[8]), n_mat=int(dims[9]),
                   n_halo=int(dims[10]), g=flat["g"], k_n=flat["k_n"], h_small=flat["h_small"])
        for name in ("cell_nfaces", "cell_faces", "cell_neighbors", "cell_normals", "face_is_boundary", "face_lengths", "cell_areas",
                     "cell_centroids", "bc_ptr", "bc_ghost_ids", "bc_internal_cells", "bc_normals", "bc_lengths", "halo_flip",
                     "halo_area", "hstill", "hstill_ghost", "zb_cells", "zb_ghost", "S0_cells", "ManningN_cells", "matID_cells",
                     "inletQ_TotalQ", "exitH_WSE"):
            loc[name] = arr(name)
        for k in ("bc_lengths", "halo_flip", "halo_area", "hstill_ghost", "zb_ghost", "inletQ_TotalQ", "exitH_WSE"):
            if loc[k] is None:
                loc[k] = np.zeros(0, dtype=np.uint8 if k == "halo_flip" else np.float64)
        info = dict(own=arr("own"), neighbors=[int(r) for r in arr("neighbors")], counts=[int(c) for c in arr("counts")],
                    halo_remote=arr("halo_remote") if arr("halo_remote") is not None else np.zeros(0, dtype=np.int64),
                    halo_cells=arr("halo_cells") if arr("halo_cells") is not None else np.zeros(0, dtype=np.int64))
    finally:
        lib.hg_case_free(h)
    del keep
    if Q is not None:
        Q = np.asarray(Q)
        N, own = int(flat["n_cells"]), info["own"]
        info["Q"] = np.concatenate([Q[:N][own], Q[N:2 * N][own], Q[2 * N:][own]])
    return loc, info


def extract_local_reference(flat, part, rank, Q=None, gid=None):
    """numpy twin of hg_partition_extract (tests compare the two).  Rank-local flat mesh (index_base 0).  `gid` are the global cell ids of `flat`'s cells (default: identity);
    they only define the canonical face orientation / ordering and must be consistent across ranks.

    Returns (local_flat, info); info = dict(own=global-row indices of the owned cells, neighbors=[rank...],
    counts=[entries per neighbour], Q=local state)."""
    N, ld, base, nf, faces, neigh, normals, valid = _tables(flat)
    part = np.asarray(part)
    gid = np.arange(N, dtype=np.int64) if gid is None else np.asarray(gid, dtype=np.int64)
    isb = np.asarray(flat["face_is_boundary"]).astype(bool)
    own = np.nonzero(part == rank)[0]
    n_own = own.size
    g2l = np.full(N, -1, dtype=np.int64)
    g2l[own] = np.arange(n_own)
    o_faces, o_neigh, o_valid, o_norm = faces[own], neigh[own], valid[own], normals[own]
    o_isb = isb[np.where(o_valid, o_faces, 0)] & o_valid
    interior = o_valid & ~o_isb
    nb_part = np.where(interior, part[np.where(interior, o_neigh, 0)], rank)
    cut = interior & (nb_part != rank)

    # ---- physical boundaries restricted to the owned cells (processing order kept; empty ones dropped)
    B = int(flat["n_ghost"])
    bc_ptr = np.asarray(flat["bc_ptr"], dtype=np.int64)
    b_gh = np.asarray(flat["bc_ghost_ids"], dtype=np.int64) - base
    b_ic = np.asarray(flat["bc_internal_cells"], dtype=np.int64) - base
    b_n = np.asarray(flat["bc_normals"]).reshape(2, B).T if B else np.zeros((0, 2))
    b_len = np.asarray(flat["bc_lengths"])
    counts_in = [int(flat["n_inletq"]), int(flat["n_exith"]), int(flat["n_wall"]), int(flat["n_symm"])]
    new_counts = [0, 0, 0, 0]
    ptr = [0]
    e_gh, e_ic, e_n, e_len, keep_Q, keep_W = [], [], [], [], [], []
    kb = 0
    for t in range(4):
        for kk in range(counts_in[t]):
            sl = slice(bc_ptr[kb], bc_ptr[kb + 1])
            m = g2l[b_ic[sl]] >= 0
            if m.any():
                if t == 0 and not m.all():
                    raise NotImplementedError("an inlet-q boundary is split across ranks: its conveyance sum runs over all of its "
                                              "faces -- partition with rcb_partition(..., keep_together=inlet_cell_groups(flat))")
                new_counts[t] += 1
                e_gh.append(b_gh[sl][m]); e_ic.append(g2l[b_ic[sl][m]]); e_n.append(b_n[sl][m]); e_len.append(b_len[sl][m])
                ptr.append(ptr[-1] + int(m.sum()))
                if t == 0:
                    keep_Q.append(kk)
                if t == 1:
                    keep_W.append(kk)
            kb += 1
    # ---- halo boundaries: one per neighbouring rank, entries sorted by (min gid, max gid) of the two cells
    ci, cj = np.nonzero(cut)
    nb = o_neigh[ci, cj]
    nbr_rank = part[nb]
    neighbors = sorted(set(int(r) for r in nbr_rank))
    h_ic, h_j, h_n, h_len, h_flip, h_area, h_nb, h_counts = [], [], [], [], [], [], [], []
    flen = np.asarray(flat["face_lengths"])
    area = np.asarray(flat["cell_areas"])
    for q in neighbors:
        m = nbr_rank == q
        c, j, r = ci[m], cj[m], nb[m]
        ga, gb = gid[own[c]], gid[r]
        order = np.lexsort((np.maximum(ga, gb), np.minimum(ga, gb)))
        c, j, r, ga, gb = c[order], j[order], r[order], ga[order], gb[order]
        h_ic.append(c); h_j.append(j); h_n.append(o_norm[c, j]); h_len.append(flen[o_faces[c, j]])
        h_flip.append((gb < ga).astype(np.uint8)); h_area.append(area[r]); h_nb.append(r)
        h_counts.append(c.size)
        ptr.append(ptr[-1] + c.size)
    # ---- local ghost ids: physical entries first, then halo entries, in entry order
    n_phys = sum(a.size for a in e_gh)
    n_halo_e = sum(h_counts)
    Bl = n_phys + n_halo_e
    phys_old_gh = np.concatenate(e_gh) if e_gh else np.zeros(0, dtype=np.int64)
    old2new = np.full(max(B, 1), -1, dtype=np.int64)
    old2new[phys_old_gh] = np.arange(n_phys)
    # ---- local neighbour table
    l_neigh = np.zeros_like(o_neigh)
    l_neigh[interior & ~cut] = g2l[o_neigh[interior & ~cut]]
    l_neigh[o_isb] = old2new[o_neigh[o_isb]]
    off = n_phys
    for k, q in enumerate(neighbors):
        l_neigh[h_ic[k], h_j[k]] = off + np.arange(h_counts[k])    # the cut slot itself, never a boundary slot whose ghost id collides
        off += h_counts[k]
    # ---- local faces
    used = np.unique(o_faces[o_valid])
    f2l = np.full(int(flat["n_faces"]), -1, dtype=np.int64)
    f2l[used] = np.arange(used.size)
    l_faces = np.where(o_valid, f2l[np.where(o_valid, o_faces, 0)], 0)
    l_isb = np.zeros(used.size, dtype=np.uint8)
    l_isb[f2l[o_faces[o_isb]]] = 1
    l_isb[f2l[o_faces[cut]]] = 1
    # ---- fields
    hst, zb = np.asarray(flat["hstill"]), np.asarray(flat["zb_cells"])
    hst_g, zb_g = np.asarray(flat["hstill_ghost"]), np.asarray(flat["zb_ghost"])
    halo_nb = np.concatenate(h_nb) if h_nb else np.zeros(0, dtype=np.int64)
    S0 = np.asarray(flat["S0_cells"])
    ic_all = np.concatenate(e_ic + h_ic) if (e_ic or h_ic) else np.zeros(0, dtype=np.int64)
    nrm_all = np.concatenate(e_n + h_n) if (e_n or h_n) else np.zeros((0, 2))
    loc = dict(
        n_cells=n_own, n_faces=int(used.size), n_ghost=Bl, ld=ld, index_base=0,
        cell_nfaces=nf[own], cell_faces=np.asfortranarray(l_faces).ravel(order="F"),
        cell_neighbors=np.asfortranarray(l_neigh).ravel(order="F"),
        cell_normals=np.asfortranarray(o_norm).ravel(order="F"),
        face_is_boundary=l_isb, face_lengths=flen[used], cell_areas=area[own],
        cell_centroids=(np.concatenate([np.asarray(flat["cell_centroids"])[:N][own], np.asarray(flat["cell_centroids"])[N:][own]])
                        if flat.get("cell_centroids") is not None else None),
        n_inletq=new_counts[0], n_exith=new_counts[1], n_wall=new_counts[2], n_symm=new_counts[3], n_halo=len(neighbors),
        bc_ptr=np.array(ptr, dtype=np.int64), bc_ghost_ids=np.arange(Bl, dtype=np.int64), bc_internal_cells=ic_all,
        bc_normals=np.concatenate([nrm_all[:, 0], nrm_all[:, 1]]),
        bc_lengths=np.concatenate(e_len + h_len) if (e_len or h_len) else np.zeros(0),
        halo_flip=np.concatenate([np.zeros(n_phys, dtype=np.uint8)] + h_flip) if Bl else np.zeros(0, dtype=np.uint8),
        halo_area=np.concatenate([np.ones(n_phys)] + h_area) if Bl else np.zeros(0),
        g=flat["g"], k_n=flat["k_n"], h_small=flat["h_small"],
        hstill=hst[own], hstill_ghost=np.concatenate([hst_g[phys_old_gh], hst[halo_nb]]),
        zb_cells=zb[own], zb_ghost=np.concatenate([zb_g[phys_old_gh], zb[halo_nb]]),
        S0_cells=np.concatenate([S0[:N][own], S0[N:][own]]), ManningN_cells=np.asarray(flat["ManningN_cells"])[own],
        matID_cells=(np.asarray(flat["matID_cells"])[own] if flat.get("matID_cells") is not None else None),
        n_mat=int(flat.get("n_mat", 0)),
        inletQ_TotalQ=np.asarray(flat["inletQ_TotalQ"])[keep_Q], exitH_WSE=np.asarray(flat["exitH_WSE"])[keep_W])
    info = dict(own=own, neighbors=neighbors, counts=h_counts, halo_remote=halo_nb,
                halo_cells=(np.concatenate(h_ic) if h_ic else np.zeros(0, dtype=np.int64)))
    if Q is not None:
        Q = np.asarray(Q)
        info["Q"] = np.concatenate([Q[:N][own], Q[N:2 * N][own], Q[2 * N:][own]])
    return loc, info


class HaloExchanger:
    """Grouped send/recv of the halo blocks with every neighbouring rank.

    `send` / `recv` are 1-D float64 torch tensors laid out like the context's halo buffers (block k = 6*n_k
    doubles); `ranks[k]` is the peer of block k.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""

    def __init__(self, send, recv, ranks, counts, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.send, self.recv, self.ranks, self.counts = send, recv, list(ranks), list(counts)
        self.offsets = np.concatenate([[0], np.cumsum([6 * c for c in counts])]).astype(np.int64)

    def exchange(self, with_lambda=False):
        if not self.ranks:
            return
        dist = self.dist
        ops = []
        per = 6 if with_lambda else 3
        for k, (peer, n) in enumerate(zip(self.ranks, self.counts)):
            o = int(self.offsets[k])
            ops.append(dist.P2POp(dist.isend, self.send[o:o + per * n], peer, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv[o:o + per * n], peer, self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _DevArray:
    """Zero-copy view of context-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def attach_exchanger(ctx, ranks, group=None):
    """Bind a context's halo buffers to a HaloExchanger on torch's current stream (NCCL)."""
    import torch
    counts = [int(c) for c in ctx.halo_info()]
    sp, rp, n = ctx.halo_buffers()
    if n == 0:
        return HaloExchanger(None, None, [], [])
    send = torch.as_tensor(_DevArray(sp, n), device="cuda")
    recv = torch.as_tensor(_DevArray(rp, n), device="cuda")
    if torch.cuda.current_stream().cuda_stream != 0:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    return HaloExchanger(send, recv, ranks, counts, group)


# ---------------------------------------------------------------- library-owned exchange (hg_comm.cu): wiring only
def _peer_tables(rank, neighbors, tables):
    """For each neighbour q of `rank`: (entries before rank's block in q's boundary list, index of rank in q's list).
    tables[q] = (neighbors_q, counts_q)."""
    off, idx = [], []
    for q in neighbors:
        nb_q, cnt_q = tables[q]
        j = list(nb_q).index(rank)
        off.append(int(sum(cnt_q[:j])))
        idx.append(j)
    return off, idx


def connect_contexts(ctxs, infos):
    """Several rank contexts living in THIS process (one host process driving several GPUs, or emulated ranks on one
    device): exchange the handles directly.  infos[r] = the info dict of extract_local for rank r."""
    handles = [c.comm_export() if info["neighbors"] else None for c, info in zip(ctxs, infos)]
    tables = {r: (info["neighbors"], info["counts"]) for r, info in enumerate(infos)}
    for r, (c, info) in enumerate(zip(ctxs, infos)):
        if not info["neighbors"]:
            continue
        off, idx = _peer_tables(r, info["neighbors"], tables)
        c.comm_connect([handles[q] for q in info["neighbors"]], off, idx)


def connect_ranks(ctx, info, group=None):
    """One process per GPU (torchrun): all-gather the handles and neighbour tables over torch.distributed (plumbing; the
    data path afterwards is peer stores over NVLink, no NCCL)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    mine = (ctx.comm_export() if info["neighbors"] else None, list(info["neighbors"]), [int(c) for c in info["counts"]])
    everyone = [None] * dist.get_world_size(group)
    dist.all_gather_object(everyone, mine, group=group)
    if info["neighbors"]:
        tables = {q: (e[1], e[2]) for q, e in enumerate(everyone)}
        off, idx = _peer_tables(rank, info["neighbors"], tables)
        ctx.comm_connect([everyone[q][0] for q in info["neighbors"]], off, idx)
    dist.barrier(group)


def attach_allreduce(ctx, group=None):
    """Sum over ranks for the global scalars of adaptive solves (hg_comm_set_allreduce) through torch.distributed."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"

    def allreduce(buf):
        t = torch.tensor(buf, dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=group)
        buf[:] = t.cpu().numpy()

    ctx.comm_set_allreduce(allreduce)
