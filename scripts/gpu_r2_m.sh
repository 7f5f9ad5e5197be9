#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzy_jvp.py tests/test_gpu_zzz_forward_driver.py tests/test_gpu_zzz_reference_replay.py -m gpu -q -x -p no:cacheprovider --tb=short --durations=8 2>&1 | tail -16
