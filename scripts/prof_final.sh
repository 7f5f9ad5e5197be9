# Round-1 final captures on the bench command itself (16M-cell river): one full-set capture per hot kernel.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_rhs -s 3 -c 1 -f -o gpurun_out/prof_16m_rhs_r1g python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-s 0 > gpurun_out/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_vjp -s 3 -c 1 -f -o gpurun_out/prof_16m_vjp_r1g python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-s 0 > gpurun_out/ncu16v.log 2>&1
tail -2 gpurun_out/ncu16v.log | cut -c1-200
