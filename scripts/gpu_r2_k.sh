#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vjp.py tests/test_gpu_multirank.py -m gpu -q -x -p no:cacheprovider --tb=short 2>&1 | tail -4
timeout 1200 python scripts/tune_r2.py 16 256,0,0 384,0,0 384,0,1 512,0,0 2>&1 | tee gpurun_out/tune_r2.log
