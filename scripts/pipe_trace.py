"""Timeline of the host-buffer RHS pipeline (HG_DEBUG_PIPE=1): python scripts/pipe_trace.py [million cells]"""
import os, sys, time
os.environ["HG_DEBUG_PIPE"] = "1"
import numpy as np, torch
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N = flat["n_cells"]
hQ = torch.empty(3 * N, dtype=torch.float64).pin_memory(); hD = torch.empty(3 * N, dtype=torch.float64).pin_memory()
hQ.numpy()[:] = Q0
p = S.RIVER_N_ZONES[:flat["n_mat"]].copy()
ctx = hg.Context(flat)
for i in range(3):
    t0 = time.perf_counter()
    ctx.rhs(hQ.numpy(), p, "ManningN", out=hD.numpy())
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.3f} ms wall", file=sys.stderr, flush=True)
hL = torch.empty(3 * N, dtype=torch.float64).pin_memory(); hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
hL.numpy()[:] = 1.0
pb = np.zeros(p.size)
for i in range(3):
    t0 = time.perf_counter()
    ctx.rhs_vjp_into(None, hL.numpy(), hB.numpy(), p, "ManningN", pb)
    print(f"pullback {i} (resident state, not traced): {(time.perf_counter() - t0) * 1e3:.3f} ms wall", file=sys.stderr, flush=True)
