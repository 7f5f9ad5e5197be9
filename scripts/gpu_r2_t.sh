#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzy_fused_jvp.py tests/test_gpu_zzy_jvp.py -m gpu -q -x -s -p no:cacheprovider --tb=short > gpurun_out/fjvp.log 2>&1
echo "exit $?" >> gpurun_out/fjvp.log
tail -40 gpurun_out/fjvp.log | cut -c1-250
