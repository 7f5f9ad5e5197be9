"""Round-2 launch-shape sweep of both hot kernels on one river mesh: python scripts/tune_r2.py [million cells]"""
import sys, json, time
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
N, F = flat["n_cells"], flat["n_faces"]
sn = int(flat["cell_nfaces"].sum())
Brhs = 100 * N + 32 * F + 4 * sn
Bvjp = 132 * N + 32 * F + 4 * sn
PEAK = 6539.9
lam = np.random.default_rng(0).standard_normal(3 * N)
p = np.array([0.02, 0.04, 0.05, 0.03, 0.045, 0.05])[:flat["n_mat"]]
print("N", N, "rhs B/cell", Brhs / N, "vjp B/cell", Bvjp / N, flush=True)
res = []
CFG = [(256, 0, 0), (256, 128, 0), (192, 0, 0)] if len(sys.argv) < 3 else [tuple(int(v) for v in c.split(",")) for c in sys.argv[2:]]
for tile, threads, var in CFG:
    t0 = time.time()
    ctx = hg.Context(flat, tile_cells=tile, threads=threads, vjp_variant=var)
    tc = time.time() - t0
    ctx.set_state(Q0); ctx.set_lambda(lam)
    ctx.time_rhs(5); ctx.time_vjp(5)
    tv0 = min(ctx.time_vjp(20) / 20 for _ in range(3))      # no active parameter: the tile kernel + inlet follow-ups only
    ctx.set_params(p, "ManningN")
    ctx.time_vjp(5)
    tr = min(ctx.time_rhs(20) / 20 for _ in range(3))
    tv = min(ctx.time_vjp(20) / 20 for _ in range(3))
    # sustained: ~1.5 s of back-to-back launches each
    nr = int(1500 / tr / 2); ctx.time_rhs(nr); trs = ctx.time_rhs(nr) / nr
    nv = int(1500 / tv / 2); ctx.time_vjp(nv); tvs = ctx.time_vjp(nv) / nv
    r = dict(tile=tile, threads=threads, vjp_variant=var, create_s=round(tc, 1), rhs_ms=round(tr, 4), rhs_frac=round(Brhs / tr / 1e6 / PEAK, 3),
             vjp_noparam_ms=round(tv0, 4), vjp_ms=round(tv, 4), vjp_frac=round(Bvjp / tv / 1e6 / PEAK, 3), rhs_sust_ms=round(trs, 4), rhs_sust_frac=round(Brhs / trs / 1e6 / PEAK, 3),
             vjp_sust_ms=round(tvs, 4), vjp_sust_frac=round(Bvjp / tvs / 1e6 / PEAK, 3))
    print(json.dumps(r), flush=True)
    res.append(r)
    del ctx
json.dump(res, open("gpurun_out/tune_r2.json", "w"), indent=1)
