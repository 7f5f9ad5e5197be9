"""Experiment (needs a library built with HG_NVCC_EXTRA=-DHG_PHASE_CLOCKS): where a VJP CTA's lifetime goes.
python scripts/phase_clocks.py [million cells] [tile ...]; the library prints thread 0's clock deltas per tile to stderr."""
import sys
import numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
M = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
flat, Q0 = S.river(int(M * 1e6 / 1.1 / 1000), 1000)
lam = np.random.default_rng(0).standard_normal(3 * flat["n_cells"])
for tile in [int(v) for v in sys.argv[2:]] or [256]:
    ctx = hg.Context(flat, tile_cells=tile)
    ctx.set_state(Q0); ctx.set_lambda(lam)
    print("tile", tile, "vjp ms", ctx.time_vjp(40) / 40, ctx.time_vjp(40) / 40, flush=True)
    del ctx
