#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python -m pytest tests/test_gpu_torchrun.py tests/test_gpu_multirank.py -m gpu -q -x -s -p no:cacheprovider --tb=long > gpurun_out/torchrun_$i.log 2>&1
echo "run $i exit $?"; tail -3 gpurun_out/torchrun_$i.log | cut -c1-300
done
