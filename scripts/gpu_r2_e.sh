#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_adjoint_time.py tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider --tb=short 2>&1 | tail -5
timeout 900 python scripts/tune_r2.py 16 256,128,0 2>&1 | tee gpurun_out/tune_r2.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/launches_vjp.csv python scripts/tune_r2.py 4 256,128,0 > /dev/null 2>&1
