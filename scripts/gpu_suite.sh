#!/bin/bash
# The whole device suite with its printed figures kept (profiles/roundN_gpu_tests_full.log is a copy of the log).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider --tb=short --durations=10 > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -16 gpurun_out/gpu_full.log | cut -c1-200
