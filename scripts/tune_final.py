"""Launch-shape sweep of both kernels on the 16M-cell river (after the face-block regrouping)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(int(16e6 / 1.1 / 1000), 1000)
N, F = flat["n_cells"], flat["n_faces"]
B = 100 * N + 32 * F + 4 * int(flat["cell_nfaces"].sum())
lam = np.random.default_rng(0).standard_normal(3 * N)
for tile, threads, var in [(256, 0, 0), (256, 192, 1), (256, 1010, 2), (192, 0, 0), (192, 128, 1)]:
    ctx = hg.Context(flat, tile_cells=tile, threads=threads, vjp_variant=var)
    ctx.set_state(Q0); ctx.set_lambda(lam)
    ctx.time_rhs(5); ctx.time_vjp(5)
    t = min(ctx.time_rhs(20) / 20 for _ in range(3))
    tv = min(ctx.time_vjp(20) / 20 for _ in range(3))
    print(f"tile {tile} rhs-threads {threads} vjp-variant {var}: rhs {t:.4f} ms {B / t / 1e6 / 6448.1:.3f}   vjp {tv:.4f} ms {(B + 32 * N) / tv / 1e6 / 6448.1:.3f}", flush=True)
    del ctx
