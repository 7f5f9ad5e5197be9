mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_rhs -s 3 -c 1 -o gpurun_out/prof_16m_rhs_r1 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_vjp -s 3 -c 1 -o gpurun_out/prof_16m_vjp_r1 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu16v.log 2>&1
tail -1 gpurun_out/ncu16v.log | cut -c1-200
