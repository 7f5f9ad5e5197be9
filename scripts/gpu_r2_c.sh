#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vjp.py -m gpu -q -x -s -p no:cacheprovider --tb=short 2>&1 | tail -12 | tee gpurun_out/gpu_partial.log
timeout 900 python scripts/tune_r2.py 16 256,0,0 256,128,0 2>&1 | tee gpurun_out/tune_r2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_vjp.csv python scripts/tune_r2.py 4 256,0,0 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_vjp.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
