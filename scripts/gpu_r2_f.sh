#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_adjoint_time.py tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider --tb=short 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vjp.csv python scripts/vjp_launches.py 16 2>&1 | tail -2
timeout 900 python scripts/tune_r2.py 16 256,128,0 2>&1 | tee gpurun_out/tune_r2.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
