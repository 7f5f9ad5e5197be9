#!/bin/bash
mkdir -p gpurun_out
true


timeout 600 python scripts/jvp_time.py 16 2>&1 | tee gpurun_out/jvp_time.log
