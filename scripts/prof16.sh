mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 14 -c 2 -o gpurun_out/prof_16m_r1 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu16.log 2>&1
tail -2 gpurun_out/ncu16.log | cut -c1-300
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_vjp.py -m gpu -q -x -k "simple or ensemble or error" 2>&1 | tail -6 > gpurun_out/sanitizer_memcheck.log; cat gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_vjp.py -m gpu -q -x -k "simple" 2>&1 | tail -6 > gpurun_out/sanitizer_racecheck.log; cat gpurun_out/sanitizer_racecheck.log
