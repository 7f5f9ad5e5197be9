import sys, numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(3636, 1000)
N = flat["n_cells"]
ctx = hg.Context(flat)
ctx.set_state(Q0)
print("ones  :", ctx.time_vjp(3) / 3, ctx.time_vjp(10) / 10)
ctx.set_lambda(np.random.default_rng(0).standard_normal(3 * N))
print("random:", ctx.time_vjp(3) / 3, ctx.time_vjp(10) / 10)
print("rhs   :", ctx.time_rhs(3) / 3, ctx.time_rhs(10) / 10)
