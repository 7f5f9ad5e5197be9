"""ORACLE (test infrastructure, NOT product code) -- setup side.

CPU restatement of the reference's setup path: SRH-2D readers, mesh_2D builder,
boundary-condition tables, bed data, Manning table and initial conditions.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Everything here keeps the reference's 1-based ids (cells 1..N, faces 1..F, ghosts 1..B,
boundaries 1..nB) so that the code reads like the files it follows; `flatten()` emits the
flat arrays of the C-ABI (include/hydrograd_b200.h) with index_base = 1, i.e. exactly what
the Julia shim would pass.

Reference files followed (paths relative to /root/reference/src):
  utilities/SRH_2D/SRH_2D_SRHGeom.jl   150-219 (parse), 251-382 (edges / boundary edges)
  utilities/SRH_2D/SRH_2D_SRHHydro.jl  25-93
  utilities/SRH_2D/SRH_2D_SRHMat.jl    30-108
  utilities/process_SRH_2D_input.jl    136-153 (matID_cells)
  meshes/mesh_2D.jl                    75-453 (topology), 456-652 (geometry)
  fvm/boundary_conditions/bc_2D.jl     152-304 (indices), 307-570 (static tables)
  parameters/process_bed_2D.jl         9-107 ; fvm/discretization/fvm_schemes_2D.jl 3-167
  parameters/process_ManningN_2D.jl    3-47
  fvm/initial_conditions/process_ICs_2D.jl 48-87, 145-155

Known, deliberate deviation: the reference numbers ghost cells in Julia `Dict` iteration
order (SRH_2D_SRHGeom.jl:290); results do not depend on it, here ghosts are numbered by
ascending boundary-face id.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

G_MAX_NODES_PER_ELEMENT = 8  # SRH_2D_SRHGeom.jl:2


# --------------------------------------------------------------------------- readers
def read_srhhydro(path):
    """SRH_2D_SRHHydro.jl:25-93 (only the keys the RHS path needs)."""
    res = {"ManningsN": {}, "BC": {}, "MONITORING": {}, "IQParams": {}, "EWSParamsC": {}}
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) <= 1:
                continue
            k = parts[0]
            if k == "ManningsN":
                res["ManningsN"][int(parts[1])] = float(parts[2])
            elif k == "BC":
                if parts[2] == "MONITORING":
                    res["MONITORING"][int(parts[1])] = parts[2]
                else:
                    res["BC"][int(parts[1])] = parts[2]
            elif k == "IQParams":
                res["IQParams"][int(parts[1])] = parts[2:5]
            elif k == "EWSParamsC":
                res["EWSParamsC"][int(parts[1])] = parts[2:5]
            elif k == "SimTime":
                res[k] = [float(x) for x in parts[1:4]]
            else:
                res[k] = parts[1]
    return res


def read_srhmat(path):
    """SRH_2D_SRHMat.jl:30-108 -> (numOfMaterials, {matID: [cell ids]})."""
    n_mat = -1
    zones = {}
    cur_id, cur = 0, []
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if not parts or parts[0] == "SRHMAT":
                continue
            if parts[0] == "NMaterials":
                n_mat = int(parts[1])
            elif parts[0] == "MatName":
                pass
            elif parts[0] == "Material":
                if cur_id != 0 and cur:
                    zones[cur_id] = list(cur)
                    cur = []
                cur_id = int(parts[1])
                cur += [int(p) for p in parts[2:]]
            else:
                cur += [int(p) for p in parts]
    if cur_id != 0:
        zones[cur_id] = list(cur)
    zones[0] = []
    return n_mat, zones


@dataclass
class SRHGeom:
    numOfElements: int = 0
    numOfNodes: int = 0
    elementNodesList: np.ndarray = None      # [N, 8] 1-based node ids, 0 padded
    elementNodesCount: np.ndarray = None     # [N]
    elementEdgesList: np.ndarray = None      # [N, 8] signed 1-based edge ids
    nodeCoordinates: np.ndarray = None       # [nNodes, 3]
    nodeStringsDict: dict = field(default_factory=dict)
    edges: dict = field(default_factory=dict)       # (n1,n2) sorted -> edge id
    edges_r: dict = field(default_factory=dict)     # edge id -> (n1,n2)
    edgeElements: dict = field(default_factory=dict)  # edge id -> [cells]
    boundaryEdges: dict = field(default_factory=dict)  # boundary id -> signed edge ids
    allBoundaryEdgeIDs: list = field(default_factory=list)


def read_srhgeom(path, bcDict):
    """SRH_2D_SRHGeom.jl:150-219 + 251-382."""
    elems, nodes, nstr = {}, {}, {}
    cur_ns = -1
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if not parts:
                continue
            k = parts[0]
            if k == "Elem":
                elems[int(parts[1])] = [int(p) for p in parts[2:]]
            elif k == "Node":
                nodes[int(parts[1])] = [float(p) for p in parts[2:]]
            elif k == "NodeString":
                cur_ns = int(parts[1])
                nstr[cur_ns] = [int(p) for p in parts[2:]]
            elif k.lower() in ("name", "gridunit"):
                pass
            elif "srhgeom" not in k.lower():
                if cur_ns > 0:
                    nstr[cur_ns] += [int(p) for p in parts]
    g = SRHGeom()
    g.numOfElements, g.numOfNodes = len(elems), len(nodes)
    N = g.numOfElements
    g.elementNodesList = np.zeros((N, G_MAX_NODES_PER_ELEMENT), dtype=np.int64)
    g.elementNodesCount = np.zeros(N, dtype=np.int64)
    g.elementEdgesList = np.zeros((N, G_MAX_NODES_PER_ELEMENT), dtype=np.int64)
    g.nodeCoordinates = np.zeros((g.numOfNodes, 3))
    for eid, nl in elems.items():
        g.elementNodesList[eid - 1, : len(nl)] = nl
        g.elementNodesCount[eid - 1] = len(nl)
    for nid, c in nodes.items():
        g.nodeCoordinates[nid - 1, :] = c
    g.nodeStringsDict = dict(sorted(nstr.items()))

    # edges: id assigned at first appearance while sweeping cells then local edges (251-285)
    cur = 1
    for c in range(1, N + 1):
        cnt = g.elementNodesCount[c - 1]
        for i in range(cnt):
            n1 = int(g.elementNodesList[c - 1, i])
            n2 = int(g.elementNodesList[c - 1, (i + 1) % cnt])
            key = (min(n1, n2), max(n1, n2))
            if key not in g.edges:
                g.edges[key] = cur
                g.edges_r[cur] = key
                g.edgeElements[cur] = [c]
                cur += 1
            else:
                e = g.edges[key]
                assert len(g.edgeElements[e]) == 1
                g.edgeElements[e].append(c)
    # deviation: ascending edge id instead of Dict order (see module docstring)
    g.allBoundaryEdgeIDs = sorted(e for e, el in g.edgeElements.items() if len(el) == 1)
    used = {e: False for e in g.allBoundaryEdgeIDs}
    for ns, nl in g.nodeStringsDict.items():
        if ns not in bcDict or "WEIR" in bcDict[ns] or "PRESSURE" in bcDict[ns]:
            continue
        lst = []
        for i in range(len(nl) - 1):
            fwd, rev = (nl[i], nl[i + 1]), (nl[i + 1], nl[i])
            if fwd in g.edges:
                lst.append(g.edges[fwd]); used[g.edges[fwd]] = True
            elif rev in g.edges:
                lst.append(-g.edges[rev]); used[g.edges[rev]] = True
            else:
                raise ValueError(f"Boundary edge {fwd} in NodeString {ns} not found")
        g.boundaryEdges[ns] = lst
    unused = [e for e in g.allBoundaryEdgeIDs if not used[e]]
    if unused:
        g.boundaryEdges[len(g.boundaryEdges) + 1] = unused
    for c in range(1, N + 1):
        cnt = g.elementNodesCount[c - 1]
        for i in range(cnt):
            n1 = int(g.elementNodesList[c - 1, i])
            n2 = int(g.elementNodesList[c - 1, (i + 1) % cnt])
            if (n1, n2) in g.edges:
                g.elementEdgesList[c - 1, i] = g.edges[(n1, n2)]
            else:
                g.elementEdgesList[c - 1, i] = -g.edges[(n2, n1)]
    return g


# --------------------------------------------------------------------------- mesh_2D
@dataclass
class Mesh2D:
    """Fields of `mesh_2D` (meshes/mesh_2D.jl:2-71) that the RHS path touches."""
    numOfCells: int
    numOfFaces: int
    numOfAllBounaryFaces: int
    cellNodesList: np.ndarray
    cellNodesCount: np.ndarray
    cellFacesList: np.ndarray            # [N, 8] abs face ids
    cellNeighbors: list                  # list (per cell) of neighbour ids / ghost ids
    faceCells: dict
    faceNodes_r: dict
    bFace_is_boundary: np.ndarray        # [F] bool (index f-1)
    boundaryFaces: dict                  # boundary id -> [face ids]
    boundaryFaces_direction: dict        # boundary id -> [+1/-1]
    allBoundaryFacesIDs_List: list
    boundaryFaceID_to_ghostCellID: dict
    boundaryFaceID_to_internalCellID: dict
    cell_areas: np.ndarray
    cell_centroids: np.ndarray
    cell_normals: list                   # [N][nF] -> (nx, ny)
    cell_distances_to_neighbors: list
    face_normals: np.ndarray             # [F, 2]
    face_lengths: np.ndarray             # [F]
    numOfBoundaries: int = 0


def _polygon_area(v):
    """mesh_2D.jl:600-616 (shoelace, sequential accumulation)."""
    a = 0.0
    n = len(v)
    for i in range(n):
        x1, y1 = v[i]
        x2, y2 = v[(i + 1) % n]
        a += x1 * y2 - x2 * y1
    if a < 0:
        raise ValueError("polygon is not counter-clockwise")
    return abs(a) / 2


def _polygon_centroid(v):
    """mesh_2D.jl:619-634."""
    cx = cy = s = 0.0
    n = len(v)
    for i in range(n):
        x1, y1 = v[i]
        x2, y2 = v[(i + 1) % n]
        cr = x1 * y2 - x2 * y1
        cx += (x1 + x2) * cr
        cy += (y1 + y2) * cr
        s += cr
    a = abs(s) / 2
    return cx / (6 * a), cy / (6 * a)


def initialize_mesh_2D(geom: SRHGeom, srhhydro_BC: dict) -> Mesh2D:
    """meshes/mesh_2D.jl:75-453 and 456-652. Mutates srhhydro_BC like the reference (172-181)."""
    N = geom.numOfElements
    cnt = geom.elementNodesCount
    faceCells = geom.edgeElements
    F = len(faceCells)
    signed = geom.elementEdgesList
    signed_set = set(int(x) for x in signed.ravel() if x != 0)
    boundaryFaces = {k: [abs(e) for e in v] for k, v in geom.boundaryEdges.items()}
    direction = {}
    for bid, fl in boundaryFaces.items():
        d = []
        for fid in fl:
            if fid in signed_set and -fid in signed_set:
                raise ValueError("boundary face ID is both positive and negative")
            d.append(-1 if -fid in signed_set else 1)
        direction[bid] = d
    cellFacesList = np.abs(signed)
    if len(srhhydro_BC) < len(boundaryFaces):
        assert len(srhhydro_BC) == len(boundaryFaces) - 1
        for k in boundaryFaces:
            if k not in srhhydro_BC:
                srhhydro_BC[k] = "wall"
    allB = list(geom.allBoundaryEdgeIDs)
    B = len(allB)
    f2g = {allB[i]: i + 1 for i in range(B)}
    f2c = {}
    for fid in allB:
        assert len(faceCells[fid]) == 1
        f2c[fid] = faceCells[fid][0]
    neigh = []
    for c in range(1, N + 1):
        lst = []
        for j in range(cnt[c - 1]):
            fid = int(cellFacesList[c - 1, j])
            fc = faceCells[fid]
            if len(fc) == 2:
                lst.append(fc[1] if fc[0] == c else fc[0])
            else:
                lst.append(f2g[fid])
        neigh.append(lst)
    isb = np.zeros(F, dtype=bool)
    for fid, fc in faceCells.items():
        isb[fid - 1] = len(fc) != 2

    # ---- geometry (456-652)
    xy = geom.nodeCoordinates
    areas = np.zeros(N)
    cent = np.zeros((N, 2))
    cnormals, cdist = [], []
    for c in range(N):
        nl = geom.elementNodesList[c, : cnt[c]]
        v = [(float(xy[n - 1, 0]), float(xy[n - 1, 1])) for n in nl]
        areas[c] = _polygon_area(v)
        cent[c, :] = _polygon_centroid(v)
        nn = []
        for i in range(len(v)):                       # 550-575
            x1, y1 = v[i]
            x2, y2 = v[(i + 1) % len(v)]
            fx, fy = x2 - x1, y2 - y1
            n0, n1 = fy, -fx
            ln = math.sqrt(n0 * n0 + n1 * n1)
            nn.append((n0 / ln, n1 / ln))
        cnormals.append(nn)
    for c in range(N):                                # 494-530
        dl = []
        for j in range(cnt[c]):
            fid = int(cellFacesList[c, j])
            if isb[fid - 1]:
                n1, n2 = geom.edges_r[fid]
                fc = (xy[n1 - 1, :2] + xy[n2 - 1, :2]) / 2
                dl.append(float(np.linalg.norm(cent[c] - fc)))
            else:
                dl.append(float(np.linalg.norm(cent[c] - cent[neigh[c][j] - 1])))
        cdist.append(dl)
    fnorm = np.zeros((F, 2))
    flen = np.zeros(F)
    for fid in range(1, F + 1):                       # 637-652
        n1, n2 = geom.edges_r[fid]
        x1, y1 = float(xy[n1 - 1, 0]), float(xy[n1 - 1, 1])
        x2, y2 = float(xy[n2 - 1, 0]), float(xy[n2 - 1, 1])
        nx, ny = y2 - y1, -(x2 - x1)
        ln = math.sqrt(nx * nx + ny * ny)
        fnorm[fid - 1] = (nx / ln, ny / ln)
        flen[fid - 1] = ln
    return Mesh2D(
        numOfCells=N, numOfFaces=F, numOfAllBounaryFaces=B,
        cellNodesList=geom.elementNodesList, cellNodesCount=cnt, cellFacesList=cellFacesList,
        cellNeighbors=neigh, faceCells=faceCells, faceNodes_r=geom.edges_r, bFace_is_boundary=isb,
        boundaryFaces=boundaryFaces, boundaryFaces_direction=direction,
        allBoundaryFacesIDs_List=allB, boundaryFaceID_to_ghostCellID=f2g,
        boundaryFaceID_to_internalCellID=f2c, cell_areas=areas, cell_centroids=cent,
        cell_normals=cnormals, cell_distances_to_neighbors=cdist, face_normals=fnorm,
        face_lengths=flen, numOfBoundaries=len(srhhydro_BC))


# --------------------------------------------------------------------------- boundary tables
@dataclass
class BoundaryConditions2D:
    """Static part of fvm/boundary_conditions/bc_2D.jl:3-47, built as in 50-570."""
    kinds: dict      # "inletQ"/"exitH"/"wall"/"symm" -> list of dict(faceIDs, ghostCellIDs, internalCellIDs, normals[n,2], lengths[n])
    inletQ_TotalQ: np.ndarray
    exitH_WSE: np.ndarray
    all_boundary_ghost_ids: list
    all_boundary_ghost_indices: list


def initialize_boundary_conditions_2D(mesh: Mesh2D, hydro: dict) -> BoundaryConditions2D:
    bc = hydro["BC"]
    idx = {"inletQ": [], "exitH": [], "wall": [], "symm": []}
    for ib in range(1, mesh.numOfBoundaries + 1):       # bc_2D.jl:163-241
        if ib not in bc:
            raise KeyError(f"boundary {ib} missing from srhhydro BC")
        t = bc[ib].lower()
        if t == "inlet-q":
            idx["inletQ"].append(ib)
        elif t == "exit-h":
            idx["exitH"].append(ib)
        elif t == "wall":
            idx["wall"].append(ib)
        elif t == "symm":
            idx["symm"].append(ib)
    kinds = {}
    for kind, ids in idx.items():
        lst = []
        for ib in ids:                                   # 321-377 and twins
            faces = mesh.boundaryFaces[ib]
            dirs = mesh.boundaryFaces_direction[ib]
            n = len(faces)
            normals = np.zeros((n, 2))
            lengths = np.zeros(n)
            for i, fid in enumerate(faces):
                normals[i] = dirs[i] * mesh.face_normals[fid - 1]
                lengths[i] = mesh.face_lengths[fid - 1]
            lst.append(dict(boundary=ib, faceIDs=list(faces),
                            ghostCellIDs=[mesh.boundaryFaceID_to_ghostCellID[f] for f in faces],
                            internalCellIDs=[mesh.boundaryFaceID_to_internalCellID[f] for f in faces],
                            normals=normals, lengths=lengths))
        kinds[kind] = lst
    Q = np.array([float(hydro["IQParams"][b["boundary"]][0]) for b in kinds["inletQ"]])
    W = np.array([float(hydro["EWSParamsC"][b["boundary"]][0]) for b in kinds["exitH"]])
    ghost_ids = []
    for kind in ("inletQ", "exitH", "wall", "symm"):     # 279-295
        for b in kinds[kind]:
            ghost_ids += b["ghostCellIDs"]
    pos = {g: i + 1 for i, g in reversed(list(enumerate(ghost_ids)))}  # findfirst (298)
    indices = [pos[i] for i in range(1, len(ghost_ids) + 1)]
    return BoundaryConditions2D(kinds, Q, W, ghost_ids, indices)


# --------------------------------------------------------------------------- fields
def nodes_to_cells_scalar(mesh: Mesh2D, zn):
    """fvm_schemes_2D.jl:108-117 (mean over the cell's nodes, sequential sum)."""
    out = np.zeros(mesh.numOfCells)
    for c in range(mesh.numOfCells):
        s = 0.0
        for j in range(mesh.cellNodesCount[c]):
            s += float(zn[mesh.cellNodesList[c, j] - 1])
        out[c] = s / int(mesh.cellNodesCount[c])
    return out


def update_bed_data(mesh: Mesh2D, zb):
    """process_bed_2D.jl:46-66 -> zb_ghost[B], zb_faces[F], S0_cells[N,2] (S0_faces unused by the flux)."""
    zb = np.asarray(zb, dtype=np.float64)
    B = mesh.numOfAllBounaryFaces
    zg = np.array([zb[mesh.faceCells[mesh.allBoundaryFacesIDs_List[b]][0] - 1] for b in range(B)])
    zf = np.zeros(mesh.numOfFaces)
    for fid in range(1, mesh.numOfFaces + 1):           # fvm_schemes_2D.jl:89-105
        fc = mesh.faceCells[fid]
        zf[fid - 1] = (zb[fc[0] - 1] + zb[fc[1] - 1]) / 2.0 if len(fc) == 2 else zb[fc[0] - 1]
    S0 = np.zeros((mesh.numOfCells, 2))
    for c in range(mesh.numOfCells):                    # fvm_schemes_2D.jl:133-167
        gx = gy = 0.0
        for j in range(mesh.cellNodesCount[c]):
            fid = int(mesh.cellFacesList[c, j])
            nx, ny = mesh.cell_normals[c][j]
            gx = gx + nx * zf[fid - 1] * mesh.face_lengths[fid - 1]
            gy = gy + ny * zf[fid - 1] * mesh.face_lengths[fid - 1]
        S0[c, 0] = -1.0 * (gx / mesh.cell_areas[c])
        S0[c, 1] = -1.0 * (gy / mesh.cell_areas[c])
    return zg, zf, S0


def matID_cells(mesh: Mesh2D, zones: dict):
    """process_SRH_2D_input.jl:136-153 (first zone containing the cell; 0 = default)."""
    m = np.zeros(mesh.numOfCells, dtype=np.int64)
    owner = {}
    for k, v in zones.items():
        for c in v:
            owner.setdefault(c, k)
    for c in range(1, mesh.numOfCells + 1):
        m[c - 1] = owner.get(c, 0)
    return m


# --------------------------------------------------------------------------- a full case
@dataclass
class Case:
    """Everything `SWE2D_Extra_Parameters` (applications/application_commons.jl:7-44) holds for the RHS."""
    mesh: Mesh2D
    bc: BoundaryConditions2D
    nodeCoordinates: np.ndarray
    zb_cells: np.ndarray
    zb_ghost: np.ndarray
    zb_faces: np.ndarray
    S0_cells: np.ndarray          # [N,2]
    matID: np.ndarray
    ManningN_zone: np.ndarray     # zone values 0..nMat-1 (solve_swe_2D.jl:168)
    ManningN_cells: np.ndarray
    wstill: np.ndarray
    hstill: np.ndarray
    hstill_ghost: np.ndarray
    Q0: np.ndarray                # [3N] = vcat(xi, q_x, q_y)  (solve_swe_2D.jl:224)
    g: float = 9.81
    k_n: float = 1.0
    h_small: float = 1.0e-3


def load_case(case_dir, srhhydro_name, ic=None):
    """solve_swe_2D.jl:46-277 restricted to what the RHS reads.

    ic: ("constant", [wse, wstill, q_x, q_y]) or ("from_file", path_to_json)."""
    hydro = read_srhhydro(os.path.join(case_dir, srhhydro_name))
    grid = hydro["Grid"].strip('"')
    matf = hydro["HydroMat"].strip('"')
    bcd = dict(hydro["BC"])
    geom = read_srhgeom(os.path.join(case_dir, grid), bcd)
    mesh = initialize_mesh_2D(geom, hydro["BC"])
    _, zones = read_srhmat(os.path.join(case_dir, matf))
    mid = matID_cells(mesh, zones)
    bc = initialize_boundary_conditions_2D(mesh, hydro)
    zb = nodes_to_cells_scalar(mesh, geom.nodeCoordinates[:, 2])
    zg, zf, S0 = update_bed_data(mesh, zb)
    nz = hydro["ManningsN"]
    n_zone = np.array([nz[i] for i in range(len(nz))])
    n_cells = np.array([nz[int(m)] for m in mid])
    N = mesh.numOfCells
    if ic is None:
        ic = ("constant", [1.0, 0.5, 0.0, 0.0])
    if ic[0] == "constant":
        wse_c, wst_c, qx_c, qy_c = ic[1]
        wse = np.full(N, float(wse_c)); wst = np.full(N, float(wst_c))
        qx = np.full(N, float(qx_c)); qy = np.full(N, float(qy_c))
    else:
        d = json.load(open(ic[1]))
        wse, wst = np.array(d["wse"], float), np.array(d["wstill"], float)
        qx, qy = np.array(d["q_x"], float), np.array(d["q_y"], float)
    hstill = wst - zb
    xi = wse - wst
    hg = np.array([hstill[mesh.faceCells[f][0] - 1] for f in mesh.allBoundaryFacesIDs_List])
    return Case(mesh, bc, geom.nodeCoordinates, zb, zg, zf, S0, mid, n_zone, n_cells, wst, hstill, hg,
                np.concatenate([xi, qx, qy]))


def flatten(case: Case):
    """Flat arrays of the C-ABI descriptors (include/hydrograd_b200.h), index_base = 1.

    Layouts are the Julia ones: N x ld tables column-major, cell_normals (i,j,k) -> i + N*(j + ld*k)."""
    m, bc = case.mesh, case.bc
    N, F, B, ld = m.numOfCells, m.numOfFaces, m.numOfAllBounaryFaces, G_MAX_NODES_PER_ELEMENT
    neigh = np.zeros((N, ld), dtype=np.int64)
    normals = np.zeros((N, ld, 2))
    for c in range(N):
        k = int(m.cellNodesCount[c])
        neigh[c, :k] = m.cellNeighbors[c]
        normals[c, :k, :] = m.cell_normals[c]
    ptr = [0]
    gh, ic, nrm, ln = [], [], [], []
    counts = []
    for kind in ("inletQ", "exitH", "wall", "symm"):
        counts.append(len(bc.kinds[kind]))
        for b in bc.kinds[kind]:
            gh += b["ghostCellIDs"]; ic += b["internalCellIDs"]
            nrm.append(b["normals"]); ln.append(b["lengths"])
            ptr.append(len(gh))
    nrm = np.concatenate(nrm) if nrm else np.zeros((0, 2))
    ln = np.concatenate(ln) if ln else np.zeros(0)
    return dict(
        n_cells=N, n_faces=F, n_ghost=B, ld=ld, index_base=1,
        cell_nfaces=np.ascontiguousarray(m.cellNodesCount, dtype=np.int64),
        cell_faces=np.asfortranarray(m.cellFacesList.astype(np.int64)).ravel(order="F").copy(),
        cell_neighbors=np.asfortranarray(neigh).ravel(order="F").copy(),
        cell_normals=np.asfortranarray(normals).ravel(order="F").copy(),
        face_is_boundary=m.bFace_is_boundary.astype(np.uint8),
        face_lengths=m.face_lengths.copy(), cell_areas=m.cell_areas.copy(),
        cell_centroids=np.asfortranarray(m.cell_centroids).ravel(order="F").copy(),
        n_inletq=counts[0], n_exith=counts[1], n_wall=counts[2], n_symm=counts[3],
        bc_ptr=np.array(ptr, dtype=np.int64), bc_ghost_ids=np.array(gh, dtype=np.int64),
        bc_internal_cells=np.array(ic, dtype=np.int64),
        bc_normals=np.asfortranarray(nrm).ravel(order="F").copy(), bc_lengths=ln.copy(),
        g=case.g, k_n=case.k_n, h_small=case.h_small,
        hstill=case.hstill.copy(), hstill_ghost=case.hstill_ghost.copy(),
        zb_cells=case.zb_cells.copy(), zb_ghost=case.zb_ghost.copy(),
        S0_cells=np.asfortranarray(case.S0_cells).ravel(order="F").copy(),
        ManningN_cells=case.ManningN_cells.copy(), matID_cells=case.matID.astype(np.int64),
        n_mat=len(case.ManningN_zone), inletQ_TotalQ=case.bc.inletQ_TotalQ.copy(),
        exitH_WSE=case.bc.exitH_WSE.copy())
