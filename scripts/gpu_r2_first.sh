#!/bin/bash
# Round 2, first GPU call: the whole device suite WITHOUT -x (so one failure cannot hide the rest), then ncu captures of both hot kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 --tb=short > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -60 gpurun_out/gpu_full.log
bash scripts/prof_final.sh
