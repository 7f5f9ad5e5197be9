#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider --tb=short > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -4 gpurun_out/gpu_full.log
HG_DEBUG_TIMING=1 timeout 600 python scripts/vjp_launches.py 16 2>&1 | tail -8
