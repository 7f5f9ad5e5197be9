set -x
mkdir -p gpurun_out
T=${1:-256}
ncu --set full --clock-control none --import-source on -k regex:k_fused_rhs -s 3 -c 1 -o gpurun_out/prof_fused_${2:-r1b} python bench.py --steps 3 --warmup 3 --cells-m 4 --no-cpu --e2e-steps 1 --tile $T > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log
