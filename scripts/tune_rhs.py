"""RHS kernel tuning sweep on one mesh: python scripts/tune_rhs.py [million cells]"""
import sys, time
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
nx = int(M * 1e6 / 1.1 / 1000); flat, Q0 = S.river(nx, 1000)
N, F = flat["n_cells"], flat["n_faces"]
B = 100 * N + 32 * F + 4 * int(flat["cell_nfaces"].sum())
print("N", N, "bytes/cell", B / N, flush=True)
import numpy as np
lam = np.random.default_rng(0).standard_normal(3 * N)
for tile, var, fb in [(256, 0, True), (256, 1, True), (192, 0, True)]:
    ctx = hg.Context(flat, tile_cells=tile, vjp_variant=var, face_blocks=fb)
    ctx.set_state(Q0); ctx.set_lambda(lam)
    ctx.time_rhs(5); ctx.time_vjp(5)
    t = min(ctx.time_rhs(20) / 20 for _ in range(3))
    tv = min(ctx.time_vjp(20) / 20 for _ in range(3))
    print(f"tile {tile} face_blocks {fb}: rhs {t:.4f} ms {B / t / 1e6 / 6448.1:.3f}   vjp {tv:.4f} ms {(B + 32 * N) / tv / 1e6 / 6448.1:.3f}", flush=True)
    del ctx
