"""Three RHS + three VJP launches with the Manning zones active on a river mesh (for an ncu launch list): python scripts/vjp_launches.py [million cells]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
ctx = hg.Context(flat, threads=128)
ctx.set_state(Q0); ctx.set_lambda(np.random.default_rng(0).standard_normal(3 * flat["n_cells"]))
ctx.set_params(S.RIVER_N_ZONES[:flat["n_mat"]], "ManningN")
ctx.time_rhs(3); ctx.time_vjp(3)
print("rhs ms", ctx.time_rhs(10) / 10, "vjp ms", ctx.time_vjp(10) / 10)
