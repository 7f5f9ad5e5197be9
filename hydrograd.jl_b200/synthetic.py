"""Synthetic unstructured meshes for the throughput configurations of BASELINE.json:

  C2  `dam_break`  ~1M-cell jittered quad/tri mesh on [0,1000]^2, all-wall, dam-break initial state
  C3  `river`      ~16M-cell meandering river, parabolic bed + valley slope, 6 Manning zones,
                   inlet-Q upstream / exit-H downstream / walls on the banks

Both are a logical ni x nj block of quadrilaterals whose nodes are displaced (jitter / meander map) and a
seeded ~10 % of the quads split into two triangles, so the mesh is genuinely mixed tri/quad and its data
go through exactly the same flat tables (include/hydrograd_b200.h) as an SRH-2D mesh would.  Everything is
vectorised numpy (O(N)); nothing here is on the timed path.  Geometry formulas follow the reference's
mesh builder (meshes/mesh_2D.jl:550-652: shoelace area, outward edge normals, centroid) and bed
preprocessing (parameters/process_bed_2D.jl:46-66) so that a generated case is a valid reference input.

Cell numbering: quad (i, j) -> running id in (i-major, j-minor) order; a split quad contributes two
consecutive ids (triangle A = nodes 0,1,2 then triangle B = nodes 0,2,3).
"""
from __future__ import annotations

import numpy as np


def _build(X, Y, Z, split, west, east, ni0=0, ni_total=None):
    """X, Y, Z: node arrays [ni+1, nj+1]; split: bool [ni, nj]; west/east: 'wall' | 'inletq' | 'exith' | 'halo'.

    Returns the flat dict (index_base 0, ld 4) plus per-cell helper arrays."""
    ni, nj = split.shape
    nq = ni * nj
    sp = split.ravel()
    base = np.arange(nq, dtype=np.int64) + np.concatenate(([0], np.cumsum(sp)[:-1])).astype(np.int64)
    idA = base.reshape(ni, nj)
    idB = (base + sp).reshape(ni, nj)
    N = int(nq + sp.sum())

    # ---- face ids
    n_ie = ni * (nj + 1)                       # i-edges  (i,j)->(i+1,j)
    n_je = (ni + 1) * nj                       # j-edges  (i,j)->(i,j+1)
    Ei = np.arange(n_ie, dtype=np.int64).reshape(ni, nj + 1)
    Ej = n_ie + np.arange(n_je, dtype=np.int64).reshape(ni + 1, nj)
    diag = np.full(nq, -1, dtype=np.int64)
    diag[sp] = n_ie + n_je + np.arange(int(sp.sum()), dtype=np.int64)
    diag = diag.reshape(ni, nj)
    F = int(n_ie + n_je + sp.sum())

    # ---- ghost ids: south (j=0) [ni], north (j=nj) [ni], west (i=0) [nj], east (i=ni) [nj]
    B = 2 * ni + 2 * nj
    g_s = np.arange(ni, dtype=np.int64)
    g_n = ni + np.arange(ni, dtype=np.int64)
    g_w = 2 * ni + np.arange(nj, dtype=np.int64)
    g_e = 2 * ni + nj + np.arange(nj, dtype=np.int64)

    # ---- neighbours through the four quad edges (ghost id on the boundary)
    nb_s = np.empty((ni, nj), np.int64); nb_s[:, 1:] = idB[:, :-1]; nb_s[:, 0] = g_s
    nb_e = np.empty((ni, nj), np.int64); nb_e[:-1, :] = idB[1:, :]; nb_e[-1, :] = g_e
    nb_n = np.empty((ni, nj), np.int64); nb_n[:, :-1] = idA[:, 1:]; nb_n[:, -1] = g_n
    nb_w = np.empty((ni, nj), np.int64); nb_w[1:, :] = idA[:-1, :]; nb_w[0, :] = g_w

    # ---- node coordinates of each quad's corners, CCW: (i,j) (i+1,j) (i+1,j+1) (i,j+1)
    def corners(A):
        return A[:-1, :-1], A[1:, :-1], A[1:, 1:], A[:-1, 1:]
    x0, x1, x2, x3 = corners(X)
    y0, y1, y2, y3 = corners(Y)
    z0, z1, z2, z3 = corners(Z)

    def edge(xa, ya, xb, yb):                  # outward normal of edge a->b of a CCW polygon + length
        nx, ny = yb - ya, -(xb - xa)
        ln = np.sqrt(nx * nx + ny * ny)
        return nx / ln, ny / ln, ln

    e_s = edge(x0, y0, x1, y1); e_e = edge(x1, y1, x2, y2); e_n = edge(x2, y2, x3, y3); e_w = edge(x3, y3, x0, y0)
    e_d_A = edge(x2, y2, x0, y0)               # diagonal as seen from triangle A (n2 -> n0)
    e_d_B = edge(x0, y0, x2, y2)               # ... from triangle B (n0 -> n2)

    ld = 4
    nf = np.full(N, 4, dtype=np.int64)
    faces = np.zeros((N, ld), dtype=np.int64)
    neigh = np.zeros((N, ld), dtype=np.int64)
    normals = np.zeros((N, ld, 2))
    flen = np.zeros(F)
    area = np.zeros(N)
    cx = np.zeros(N); cy = np.zeros(N); zb = np.zeros(N)

    q = ~split
    a = idA[q]                                  # ---- unsplit quads
    for k, (fid, nb, e) in enumerate(((Ei[:, :-1], nb_s, e_s), (Ej[1:, :], nb_e, e_e), (Ei[:, 1:], nb_n, e_n), (Ej[:-1, :], nb_w, e_w))):
        faces[a, k] = fid[q]; neigh[a, k] = nb[q]; normals[a, k, 0] = e[0][q]; normals[a, k, 1] = e[1][q]

    def poly(xs, ys, m):
        ar = np.zeros(int(m.sum())); sx = np.zeros_like(ar); sy = np.zeros_like(ar)
        n = len(xs)
        for k in range(n):
            xa, ya, xb, yb = xs[k][m], ys[k][m], xs[(k + 1) % n][m], ys[(k + 1) % n][m]
            cr = xa * yb - xb * ya
            ar += cr; sx += (xa + xb) * cr; sy += (ya + yb) * cr
        A = np.abs(ar) / 2
        return A, sx / (6 * A), sy / (6 * A)

    area[a], cx[a], cy[a] = poly((x0, x1, x2, x3), (y0, y1, y2, y3), q)
    zb[a] = (((z0[q] + z1[q]) + z2[q]) + z3[q]) / 4
    s = split                                   # ---- split quads: triangle A (0,1,2), triangle B (0,2,3)
    if s.any():
        ta, tb = idA[s], idB[s]
        nf[ta] = 3; nf[tb] = 3
        for k, (fid, nb, e) in enumerate(((Ei[:, :-1], nb_s, e_s), (Ej[1:, :], nb_e, e_e), (diag, idB, e_d_A))):
            faces[ta, k] = fid[s]; neigh[ta, k] = nb[s]; normals[ta, k, 0] = e[0][s]; normals[ta, k, 1] = e[1][s]
        for k, (fid, nb, e) in enumerate(((diag, idA, e_d_B), (Ei[:, 1:], nb_n, e_n), (Ej[:-1, :], nb_w, e_w))):
            faces[tb, k] = fid[s]; neigh[tb, k] = nb[s]; normals[tb, k, 0] = e[0][s]; normals[tb, k, 1] = e[1][s]
        area[ta], cx[ta], cy[ta] = poly((x0, x1, x2), (y0, y1, y2), s)
        area[tb], cx[tb], cy[tb] = poly((x0, x2, x3), (y0, y2, y3), s)
        zb[ta] = ((z0[s] + z1[s]) + z2[s]) / 3
        zb[tb] = ((z0[s] + z2[s]) + z3[s]) / 3
        flen[diag[s]] = e_d_A[2][s]
    flen[Ei[:, :-1].ravel()] = e_s[2].ravel(); flen[Ei[:, -1]] = e_n[2][:, -1]
    flen[Ej[:-1, :].ravel()] = e_w[2].ravel(); flen[Ej[-1, :]] = e_e[2][-1, :]
    isb = np.zeros(F, dtype=np.uint8)
    isb[Ei[:, 0]] = 1; isb[Ei[:, -1]] = 1; isb[Ej[0, :]] = 1; isb[Ej[-1, :]] = 1

    # ---- boundary entries in processing order inlet-q, exit-h, wall, symm
    south = dict(g=g_s, c=idA[:, 0], n=(e_s[0][:, 0], e_s[1][:, 0]), L=e_s[2][:, 0])
    north = dict(g=g_n, c=idB[:, -1], n=(e_n[0][:, -1], e_n[1][:, -1]), L=e_n[2][:, -1])
    westd = dict(g=g_w, c=idB[0, :], n=(e_w[0][0, :], e_w[1][0, :]), L=e_w[2][0, :])
    eastd = dict(g=g_e, c=idA[-1, :], n=(e_e[0][-1, :], e_e[1][-1, :]), L=e_e[2][-1, :])
    groups = {"inletq": [], "exith": [], "wall": [south, north], "symm": []}
    groups[west if west != "halo" else "wall"].append(westd)
    groups[east if east != "halo" else "wall"].append(eastd)
    ptr, gh, ic, nxs, nys, Ls = [0], [], [], [], [], []
    for kind in ("inletq", "exith", "wall", "symm"):
        for b in groups[kind]:
            gh.append(b["g"]); ic.append(b["c"]); nxs.append(b["n"][0]); nys.append(b["n"][1]); Ls.append(b["L"])
            ptr.append(ptr[-1] + len(b["g"]))
    gh, ic = np.concatenate(gh), np.concatenate(ic)
    bcn = np.concatenate([np.concatenate(nxs), np.concatenate(nys)])
    ghost_cell = np.empty(B, dtype=np.int64); ghost_cell[gh] = ic

    # ---- bed slope: update_bed_data (Green-Gauss over face-averaged zb), reference accumulation order
    zf = np.empty((N, ld))
    gx = np.zeros(N); gy = np.zeros(N)
    fl = flen[faces]
    for k in range(ld):
        valid = k < nf
        interior = valid & (isb[faces[:, k]] == 0)
        nbz = np.where(interior, zb[np.where(interior, neigh[:, k], 0)], zb)
        zfk = np.where(interior, (zb + nbz) / 2.0, zb)
        gx = np.where(valid, gx + normals[:, k, 0] * zfk * fl[:, k], gx)
        gy = np.where(valid, gy + normals[:, k, 1] * zfk * fl[:, k], gy)
    S0 = np.concatenate([-1.0 * (gx / area), -1.0 * (gy / area)])

    flat = dict(
        n_cells=N, n_faces=F, n_ghost=B, ld=ld, index_base=0,
        cell_nfaces=nf, cell_faces=np.asfortranarray(faces).ravel(order="F"),
        cell_neighbors=np.asfortranarray(neigh).ravel(order="F"),
        cell_normals=np.asfortranarray(normals).ravel(order="F"),
        face_is_boundary=isb, face_lengths=flen, cell_areas=area, cell_centroids=np.concatenate([cx, cy]),
        n_inletq=len(groups["inletq"]), n_exith=len(groups["exith"]), n_wall=len(groups["wall"]), n_symm=0,
        bc_ptr=np.array(ptr, dtype=np.int64), bc_ghost_ids=gh, bc_internal_cells=ic, bc_normals=bcn,
        bc_lengths=np.concatenate(Ls), zb_cells=zb, zb_ghost=zb[ghost_cell], S0_cells=S0,
        g=9.81, k_n=1.0, h_small=1.0e-3)
    aux = dict(ni=ni, nj=nj, idA=idA, idB=idB, ghost_cell=ghost_cell, cx=cx, cy=cy)
    return flat, aux


def _split_mask(ni, nj, seed, frac=0.1):
    rng = np.random.default_rng(seed)
    return rng.random((ni, nj)) < frac


def dam_break(n=1000, seed=1234, thin_film=False):
    """C2: n x n logical quads on [0,1000]^2 m, node jitter U(-0.25,0.25)*dx (interior nodes), ~10 % of the quads
    split into triangles, flat bed, n = 0.03, wstill = 1 m, h = 10 m for x < 500 else 1 m (2e-3 m for the
    thin-film variant, which exercises the wet/dry branches), q = 0, all-wall boundary."""
    rng = np.random.default_rng(seed)
    dx = 1000.0 / n
    g = np.linspace(0.0, 1000.0, n + 1)
    X, Y = np.meshgrid(g, g, indexing="ij")
    jx = rng.uniform(-0.25, 0.25, X.shape) * dx
    jy = rng.uniform(-0.25, 0.25, X.shape) * dx
    jx[0, :] = jx[-1, :] = 0.0; jy[:, 0] = jy[:, -1] = 0.0     # keep the outer box straight
    X = X + jx; Y = Y + jy
    flat, aux = _build(X, Y, np.zeros_like(X), _split_mask(n, n, seed), "wall", "wall")
    N = flat["n_cells"]
    wstill = 1.0
    h = np.where(aux["cx"] < 500.0, 10.0, 2.0e-3 if thin_film else 1.0)
    hstill = wstill - flat["zb_cells"]
    flat.update(hstill=hstill, hstill_ghost=hstill[aux["ghost_cell"]], ManningN_cells=np.full(N, 0.03),
                matID_cells=np.zeros(N, dtype=np.int64), n_mat=1, inletQ_TotalQ=np.zeros(0), exitH_WSE=np.zeros(0))
    Q0 = np.concatenate([h - hstill, np.zeros(N), np.zeros(N)])
    return flat, Q0


RIVER_N_ZONES = np.array([0.02, 0.04, 0.05, 0.03, 0.045, 0.05])   # the Savannah zone values (savana_SI.srhhydro)


def river(ni=16000, nj=1000, seed=1234, i0=0, ni_total=None, perturb=0.05):
    """C3: ni x nj logical quads (1 m x 1 m nominal) mapped onto a sinusoidal meander (amplitude 2 km,
    wavelength 20 km, width nj metres); parabolic cross-section bed (2 m bank rise) + 1e-4 valley slope; six
    cross-stream Manning bands; inlet-Q at the upstream end (Q = 187.4 * width/100), exit-H downstream, walls
    on the banks.  State: water surface 3 m above the thalweg, q from Manning normal flow, plus a seeded
    relative perturbation so no two cells carry the same numbers.

    `i0`/`ni_total` generate the slab [i0, i0+ni) of a longer river (weak-scaling runs: each rank builds only
    its own slab plus one halo column; the cut ends are reported as 'halo' boundaries)."""
    ni_total = ni_total or ni
    W = float(nj)
    amp, lam, slope = 2000.0, 20000.0, 1.0e-4
    s = (i0 + np.arange(ni + 1, dtype=np.float64))[:, None]
    t = (np.arange(nj + 1, dtype=np.float64) - nj / 2.0)[None, :]
    kx = 2 * np.pi / lam
    yc = amp * np.sin(kx * s)
    phi = np.arctan(amp * kx * np.cos(kx * s))
    X = s - t * np.sin(phi)
    Y = yc + t * np.cos(phi)
    Z = 2.0 * (2.0 * t / W) ** 2 - slope * s + 10.0 + np.zeros_like(X)
    west = "inletq" if i0 == 0 else "halo"
    east = "exith" if i0 + ni == ni_total else "halo"
    # the split pattern must not depend on how the river is cut into slabs
    split = np.zeros((ni, nj), dtype=bool)
    for b0 in range(i0 - i0 % 1024, i0 + ni, 1024):
        blk = _split_mask(1024, nj, seed + b0 // 1024)
        lo, hi = max(b0, i0), min(b0 + 1024, i0 + ni)
        split[lo - i0:hi - i0] = blk[lo - b0:hi - b0]
    flat, aux = _build(X, Y, Z, split, west, east)
    N = flat["n_cells"]
    # local stream coordinates of every cell from its owning quad
    own_i = np.empty(N, dtype=np.int64); own_j = np.empty(N, dtype=np.int64)
    II, JJ = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
    own_i[aux["idA"].ravel()] = II.ravel(); own_i[aux["idB"].ravel()] = II.ravel()
    own_j[aux["idA"].ravel()] = JJ.ravel(); own_j[aux["idB"].ravel()] = JJ.ravel()
    sc = i0 + own_i + 0.5
    zone = np.minimum((own_j * 6) // nj, 5).astype(np.int64)
    mann = RIVER_N_ZONES[zone]
    wse = (10.0 - slope * sc) + 3.0
    wstill = np.full(N, 13.0)      # constant still-water level, as in every reference case
    zb = flat["zb_cells"]
    hstill = wstill - zb
    h = wse - zb
    rng = np.random.default_rng(seed + 7919 * (i0 + 1))
    h = h * (1.0 + perturb * rng.uniform(-1, 1, N))
    u = h ** (2.0 / 3.0) * np.sqrt(slope) / mann * (1.0 + perturb * rng.uniform(-1, 1, N))
    ph = np.arctan(amp * kx * np.cos(kx * sc)) + perturb * rng.uniform(-1, 1, N)
    flat.update(hstill=hstill, hstill_ghost=hstill[aux["ghost_cell"]], ManningN_cells=mann, matID_cells=zone,
                n_mat=6, inletQ_TotalQ=np.full(flat["n_inletq"], 187.4 * W / 100.0),
                exitH_WSE=np.full(flat["n_exith"], 10.0 - slope * ni_total + 3.0))
    Q0 = np.concatenate([h - hstill, h * u * np.cos(ph), h * u * np.sin(ph)])
    return flat, Q0


def cell_columns(flat, ni, nj):
    """Stream-wise quad column (0..ni-1) of every cell of a `river`/`dam_break` mesh (cells are numbered i-major,
    a split quad contributes two consecutive cells)."""
    N = flat["n_cells"]
    nf = np.asarray(flat["cell_nfaces"])
    # walk the cells in order: a quad closes one logical cell, two consecutive triangles close one
    w = np.where(nf == 4, 1.0, 0.5)
    q = np.floor(np.cumsum(w) - w + 1e-9).astype(np.int64)     # logical quad index of every cell
    return np.minimum(q // nj, ni - 1)
