"""Device time of the fused forward-mode kernel on the river mesh: python scripts/jvp_time.py [million cells]"""
import sys, json
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N, F, sn = flat["n_cells"], flat["n_faces"], int(flat["cell_nfaces"].sum())
ctx = hg.Context(flat)
ctx.set_state(Q0); ctx.set_lambda(np.random.default_rng(0).standard_normal(3 * N))
ctx.set_params(S.RIVER_N_ZONES[:flat["n_mat"]], "ManningN")
rhs = ctx.time_rhs(10) / 10
PEAK = 6539.9
for K in (1, 2, 6):
    ctx.time_jvp(K, 2)
    t = min(ctx.time_jvp(K, 5) / 5 for _ in range(2))
    # algorithmic bytes: state + hstill + mesh tables once per tile (shared by the K directions through L2) + per direction
    # tangent in (24) + out (24) + per-cell parameter tangent (8); values out once (24)
    mesh = 100 * N + 32 * F + 4 * sn
    b = mesh + K * 56 * N
    print(json.dumps(dict(cells=N, K=K, ms=round(t, 4), ms_per_direction=round(t / K, 4), rhs_ms=round(rhs, 4),
                          direction_cell_updates_per_s=K * N / t * 1e3, algorithmic_GBs=b / t / 1e6, roofline_frac=round(b / t / 1e6 / PEAK, 3))), flush=True)

# the reference's sensitivity driver end to end (hg_solve_tsit5_sens, six Manning zones): fused vs plain-table forward mode
import time
from oracle import srh2d_ref as R
from tests import cases
res = {}
c = cases.load("savannah")
sflat = R.flatten(c)
z = np.load("tests/golden/savannah_sens/sensitivity.npz")
for name, kw in (("fused", {}), ("strict", {"strict": True})):
    cx = hg.Context(sflat, **kw)
    cx.set_controller_pow("fastpow")
    cx.solve_tsit5_sens(c.Q0, z["params_vector"], "ManningN", 0.0, 5.0, 0.02, True, 1e-6, 1e-3)
    t0 = time.perf_counter()
    _, _, st = cx.solve_tsit5_sens(c.Q0, z["params_vector"], "ManningN", 0.0, 200.0, 0.02, True, 1e-6, 1e-3)
    res["savannah_%s_s" % name] = round(time.perf_counter() - t0, 3)
    res["savannah_rhs_evals"] = st["rhs"]
del ctx
bflat, bQ = S.river(909, 1000)
for name, kw in (("fused", {}), ("strict", {"strict": True})):
    cx = hg.Context(bflat, **kw)
    p = S.RIVER_N_ZONES[:bflat["n_mat"]]
    cx.solve_tsit5_sens(bQ, p, "ManningN", 0.0, 0.002, 0.001, False)
    t0 = time.perf_counter()
    _, _, st = cx.solve_tsit5_sens(bQ, p, "ManningN", 0.0, 0.01, 0.001, False)
    res["river1M_%s_s" % name] = round(time.perf_counter() - t0, 3)
    res["river1M_rhs_evals"] = st["rhs"]
    del cx
print(json.dumps(res))
