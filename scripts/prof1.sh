set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_rhs -s 3 -c 2 -o gpurun_out/prof_fused_r1a python bench.py --steps 3 --warmup 3 --cells-m 4 --no-cpu --e2e-steps 1 > gpurun_out/ncu_run.log 2>&1; tail -3 gpurun_out/ncu_run.log
