mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_vjp -s 2 -c 1 -o gpurun_out/prof_vjp_${1:-r1a} python scripts/dbg_vjp_time.py > gpurun_out/ncu_vjp.log 2>&1; tail -3 gpurun_out/ncu_vjp.log
