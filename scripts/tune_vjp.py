"""VJP kernel tuning sweep on one mesh: python scripts/tune_vjp.py [million cells]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N, F = flat["n_cells"], flat["n_faces"]
B = 132 * N + 32 * F + 4 * int(flat["cell_nfaces"].sum())
lam = np.random.default_rng(0).standard_normal(3 * N)
print("N", N, "vjp bytes/cell", B / N, flush=True)
for tile, var in [(256, 0), (256, 1), (192, 2)]:
    ctx = hg.Context(flat, tile_cells=tile, vjp_variant=var)
    ctx.set_state(Q0); ctx.set_lambda(lam)
    ctx.time_vjp(5)
    t = min(ctx.time_vjp(20) / 20 for _ in range(3))
    print(f"tile {tile} variant {var}: {t:.4f} ms  {B / t / 1e6:.0f} GB/s  {B / t / 1e6 / 6448.1:.3f}", flush=True)
    del ctx
