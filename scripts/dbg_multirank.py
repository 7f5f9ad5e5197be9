import sys, numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import parallel as P, synthetic as S
from tests import cases
from tests.test_gpu_multirank import _route
flat, Q0 = S.dam_break(40); Pn = 4
N = flat["n_cells"]
part = P.rcb_partition(flat["cell_centroids"][:N], flat["cell_centroids"][N:], Pn)
Q = cases.random_state_flat(flat, 7, dry_frac=0.05)
single = hg.Context(flat, tile_cells=128)
ref = single.rhs(Q)
locs = [P.extract_local(flat, part, r, Q) for r in range(Pn)]
ctxs = [hg.Context(loc, tile_cells=128) for loc, _ in locs]
infos = [info for _, info in locs]
for c, info in zip(ctxs, infos):
    c.set_state(info["Q"])
_route(ctxs, infos, with_lambda=False)
got = np.zeros(3 * N)
cutcell = np.zeros(N, bool); flipcell = np.zeros(N, bool)
for c, (loc, info) in zip(ctxs, locs):
    c.rhs_resident(); d = c.get_rhs(); n = info["own"].size
    for k in range(3): got[k * N + info["own"]] = d[k * n:(k + 1) * n]
    cutcell[info["own"][info["halo_cells"]]] = True
    nph = loc["n_ghost"] - len(info["halo_cells"])
    fl = loc["halo_flip"][nph:].astype(bool)
    flipcell[info["own"][info["halo_cells"][fl]]] = True
diff = (got != ref).reshape(3, N).any(0)
print("differing cells", diff.sum(), "of which cut-adjacent", (diff & cutcell).sum(), "flip-adjacent", (diff & flipcell).sum(), "cut cells total", cutcell.sum(), "flip cells", flipcell.sum())
h = Q[:N] + flat["hstill"]
print("dry among differing", (h[diff] <= 1e-3).sum(), " comp diffs", [(got[k*N:(k+1)*N] != ref[k*N:(k+1)*N]).sum() for k in range(3)])
i = np.nonzero(diff)[0][:5]; print(i, got[i] - ref[i], got[N+i]-ref[N+i])
ld = flat["ld"]
nb = flat["cell_neighbors"].reshape(ld, N); nfc = flat["cell_nfaces"]; fc = flat["cell_faces"].reshape(ld, N)
isb = flat["face_is_boundary"]
for i in np.nonzero(diff)[0]:
    row = []
    for j in range(nfc[i]):
        if isb[fc[j, i]]:
            row.append(("B",))
        else:
            r = nb[j, i]
            row.append((int(r), int(part[r]), "flip" if r < i else "", "dry" if h[r] <= 1e-3 else "", round(float(h[r]), 5)))
    print(i, int(part[i]), round(float(h[i]), 5), row)
