#!/bin/bash
# one short GPU call: the UDE closure tests (small fixtures first) and the Tsit5 dense-output tests
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_zz_ude.py tests/test_gpu_tsit5.py -q --tb=short -p no:cacheprovider \
  -k "(ude and not savannah) or dense" > gpurun_out/ude_tests.log 2>&1
echo "exit $?" >> gpurun_out/ude_tests.log
tail -40 gpurun_out/ude_tests.log
