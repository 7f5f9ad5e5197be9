import sys, numpy as np, time
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(1000, 1000)     # ~1.1M cells with C3-style BCs (config C5)
N = flat["n_cells"]
ctx = hg.Context(flat)
ctx.set_state(Q0)
t1 = ctx.time_rhs(20, True, 1e-4) / 20
print("single member euler step ms", t1, "cells/s", N / t1 * 1e3)
for M in (8, 32, 128):
    ctx.ensemble_alloc(M, per_member_manning=True)
    rng = np.random.default_rng(1234)
    for m in range(M):
        ctx.ensemble_set_member(m, Q0, flat_n := np.array([0.02,0.04,0.05,0.03,0.045,0.05]) * (1 + 0.2 * rng.uniform(-1, 1, 6)), "ManningN")
    ctx.time_ensemble(3, 1e-4)
    t = ctx.time_ensemble(10, 1e-4) / 10
    print("M", M, "ms/step", t, "member-cell-updates/s", M * N / t * 1e3, "speedup vs M singles", M * t1 / t)
