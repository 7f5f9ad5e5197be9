"""L2 prefetch distance sweep (RHS and VJP): python scripts/tune_prefetch.py [million cells]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N, F = flat["n_cells"], flat["n_faces"]
B = 100 * N + 32 * F + 4 * int(flat["cell_nfaces"].sum())
lam = np.random.default_rng(0).standard_normal(3 * N)
for tile, var, pf in [(256, 0, -1), (256, 0, 0), (256, 0, 370), (256, 0, 1480), (256, 0, 2960), (192, 1, -1), (192, 1, 0), (192, 1, 1184), (192, 1, 2400)]:
    ctx = hg.Context(flat, tile_cells=tile, vjp_variant=var, prefetch=pf)
    ctx.set_state(Q0); ctx.set_lambda(lam)
    ctx.time_rhs(5); ctx.time_vjp(5)
    tr = min(ctx.time_rhs(20) / 20 for _ in range(3))
    tv = min(ctx.time_vjp(20) / 20 for _ in range(3))
    print(f"tile {tile} vjpvar {var} prefetch {pf}: rhs {tr:.4f} ms ({B / tr / 1e6 / 6448.1:.3f})  vjp {tv:.4f} ms ({(B + 32 * N) / tv / 1e6 / 6448.1:.3f})", flush=True)
    del ctx
