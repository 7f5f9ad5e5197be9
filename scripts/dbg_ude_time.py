"""Cost of the (unfused) UDE closure next to the plain RHS / VJP on one B200: wall clock around back-to-back resident calls
(hg_sync on both sides), 2M-cell synthetic river.  Prints one JSON line."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S, ude as hude

t0 = time.time()
ni = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
flat, Q0 = S.river(ni, 1000)
N = int(flat["n_cells"])
ctx = hg.Context(flat)
t_setup = time.time() - t0
rng = np.random.default_rng(0)
ctx.set_state(Q0)
ctx.set_lambda(rng.standard_normal(3 * N))


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


out = dict(cells=N, setup_s=round(t_setup, 1))
out["rhs_plain_ms"] = timed(ctx.rhs_resident)
out["vjp_plain_ms"] = timed(ctx.vjp_resident)
cfg = dict(input_dim=3, output_dim=1, hidden_layers=[3, 3], activations=["tanh", "tanh"], h_bounds=[0.05, 8.0], Umag_bounds=[0.0, 3.0],
           ks_bounds=[0.02, 0.3], output_bounds=[0.02, 0.06])
ks = rng.uniform(0.02, 0.3, N)
for ln in ("whole", "cell"):
    m = hude.UDEModel("ManningN_h_Umag_ks", cfg, layernorm=ln)
    ctx.set_ude_model(m, ks)
    th = rng.uniform(-0.5, 0.5, m.n_params)
    ctx.set_params(th, "UDE")
    l0 = ctx.kernel_launches()
    ctx.rhs_resident()
    l1 = ctx.kernel_launches()
    ctx.vjp_resident()
    l2 = ctx.kernel_launches()
    out[f"rhs_ude_{ln}_ms"] = timed(ctx.rhs_resident)
    out[f"vjp_ude_{ln}_ms"] = timed(ctx.vjp_resident)
    out[f"launches_{ln}"] = [l1 - l0, l2 - l1]
    ctx.set_ude_model(None)
print(json.dumps(out))
