# ncu capture of the RHS kernel at 16M cells: bash scripts/prof_rhs2.sh <tile> <pipeline> <tag>
mkdir -p gpurun_out
cat > /tmp/run_rhs.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import _pkg; hg = _pkg.load()
from hydrograd_jl_b200 import synthetic as S
flat, Q0 = S.river(int(16e6 / 1.1 / 1000), 1000)
ctx = hg.Context(flat, tile_cells=$1)
ctx.set_state(Q0)
ctx.time_rhs(3)
PY
ncu --set full --import-source on --clock-control none -k regex:k_fused_rhs -s 2 -c 1 -o gpurun_out/prof_rhs_$3 -f python /tmp/run_rhs.py > gpurun_out/ncu_rhs_$3.log 2>&1
tail -3 gpurun_out/ncu_rhs_$3.log
