"""Does the VJP launch time depend on the cotangent's values?  (power / clocks under random operands)"""
import sys, subprocess
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(int(16e6 / 1.1 / 1000), 1000)
N = flat["n_cells"]
ctx = hg.Context(flat)
ctx.set_state(Q0)
rng = np.random.default_rng(0)
cx = flat["cell_centroids"][:N]
lams = {"ones": np.ones(3 * N), "random": rng.standard_normal(3 * N),
        "smooth": np.concatenate([np.sin(cx / 500.0), np.cos(cx / 300.0), np.sin(cx / 700.0 + 1.0)])}
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for rep in range(2):
    for name, lam in lams.items():
        ctx.set_lambda(lam)
        ctx.time_vjp(5)
        t = [ctx.time_vjp(40) / 40 for _ in range(3)]
        print(rep, name, ["%.4f" % x for x in t], clocks(), flush=True)
# the RHS with the smooth bench state vs a random state
for name, Q in (("bench state", Q0), ("random state", Q0 * (1 + 0.3 * rng.standard_normal(3 * N)))):
    ctx.set_state(Q)
    ctx.time_rhs(5)
    print("rhs", name, ["%.4f" % (ctx.time_rhs(40) / 40) for _ in range(3)], clocks(), flush=True)
