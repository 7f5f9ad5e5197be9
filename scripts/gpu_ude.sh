#!/bin/bash
# one short GPU call: the UDE closure tests (small fixtures), then the cost of the closure next to the plain RHS / VJP
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_zz_ude.py -q --tb=short -p no:cacheprovider -x \
  -k "not savannah or specialised" > gpurun_out/ude_tests2.log 2>&1
echo "exit $?" >> gpurun_out/ude_tests2.log
tail -15 gpurun_out/ude_tests2.log
timeout 40 python scripts/dbg_ude_time.py 4000 > gpurun_out/ude_time2.json 2> gpurun_out/ude_time2.err
cat gpurun_out/ude_time2.json; tail -3 gpurun_out/ude_time2.err
