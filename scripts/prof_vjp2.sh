# ncu capture of the VJP kernel at 16M cells: bash scripts/prof_vjp2.sh <tile> <variant> <tag>
mkdir -p gpurun_out
cat > /tmp/run_vjp.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(int(16e6 / 1.1 / 1000), 1000)
N = flat["n_cells"]
ctx = hg.Context(flat, tile_cells=$1, vjp_variant=$2)
ctx.set_state(Q0); ctx.set_lambda(np.random.default_rng(0).standard_normal(3 * N))
ctx.time_vjp(3)
PY
ncu --set full --import-source on --clock-control none -k regex:k_fused_vjp -s 2 -c 1 -o gpurun_out/prof_vjp_$3 -f python /tmp/run_vjp.py > gpurun_out/ncu_vjp_$3.log 2>&1
tail -3 gpurun_out/ncu_vjp_$3.log
