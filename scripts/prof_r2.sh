#!/bin/bash
# Round-2 captures on the 16M-cell river: one full-set capture per hot kernel + the launch list of the bench command.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_rhs -s 6 -c 1 -f -o gpurun_out/prof_16m_rhs_r2 python scripts/vjp_launches.py 16 > gpurun_out/ncu_rhs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_vjp -s 6 -c 1 -f -o gpurun_out/prof_16m_vjp_r2 python scripts/vjp_launches.py 16 > gpurun_out/ncu_vjp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/round2_launches_bench16M.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-side --sustained-s 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/ncu_vjp.log | cut -c1-200
