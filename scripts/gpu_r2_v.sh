#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_ude.py -m gpu -q -x -p no:cacheprovider --tb=short 2>&1 | tail -4
timeout 300 python scripts/dbg_ude_time.py 4000 > gpurun_out/ude_time.json 2> gpurun_out/ude_time.err; cat gpurun_out/ude_time.json
