#!/bin/bash
# First GPU call of the next round: the tests written after this round's GPU budget was spent, then the whole suite,
# then the cost of the UDE closure.  ~2 minutes of box time.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zzy_jvp.py tests/test_gpu_zzy_symm_boundaries.py tests/test_gpu_zzz_forward_driver.py tests/test_gpu_zzz_reference_replay.py -q --tb=short -p no:cacheprovider -s > gpurun_out/replay_tests.log 2>&1
echo "exit $?" >> gpurun_out/replay_tests.log
tail -40 gpurun_out/replay_tests.log
timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=10 > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -15 gpurun_out/gpu_full.log
timeout 60 python scripts/dbg_ude_time.py 4000 > gpurun_out/ude_time.json 2> gpurun_out/ude_time.err
cat gpurun_out/ude_time.json
