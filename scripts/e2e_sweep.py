"""Host-buffer calls (hg_rhs / hg_rhs_vjp through pinned memory) vs the number of pipeline chunks: python scripts/e2e_sweep.py [million cells]"""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N = flat["n_cells"]
hQ = torch.empty(3 * N, dtype=torch.float64).pin_memory(); hD = torch.empty(3 * N, dtype=torch.float64).pin_memory()
hL = torch.empty(3 * N, dtype=torch.float64).pin_memory(); hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
hQ.numpy()[:] = Q0; hL.numpy()[:] = 1.0
p = S.RIVER_N_ZONES[:flat["n_mat"]].copy(); pb = np.zeros(p.size)
for K in (int(a) for a in (sys.argv[2:] or ["0", "8", "12", "16", "24", "48", "64"])):
    ctx = hg.Context(flat, pipeline_chunks=K)
    ctx.rhs(hQ.numpy(), p, "ManningN", out=hD.numpy()); ctx.rhs_vjp_into(hQ.numpy(), hL.numpy(), hB.numpy(), p, "ManningN", pb)
    t0 = time.perf_counter()
    for _ in range(3): ctx.rhs(hQ.numpy(), p, "ManningN", out=hD.numpy())
    t1 = time.perf_counter()
    for _ in range(3): ctx.rhs_vjp_into(hQ.numpy(), hL.numpy(), hB.numpy(), p, "ManningN", pb)
    t2 = time.perf_counter()
    r, v = (t1 - t0) / 3 * 1e3, (t2 - t1) / 3 * 1e3
    print(json.dumps(dict(chunks=K, rhs_ms=round(r, 2), vjp_ms=round(v, 2), rhs_GBs_each=round(24 * N / r / 1e6, 1), vjp_h2d_GBs=round(48 * N / v / 1e6, 1),
                          e2e_G_cells_s=round(N / (r + v) / 1e6, 3))), flush=True)
    del ctx
