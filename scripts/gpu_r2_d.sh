#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -8 gpurun_out/gpu_full.log
HG_DEBUG_TIMING=1 timeout 900 python scripts/tune_r2.py 16 256,0,0 256,128,0 2>&1 | tee gpurun_out/tune_r2.log
