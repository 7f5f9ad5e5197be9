#!/bin/bash
# full device suite + create-time breakdown + microbenchmark
mkdir -p gpurun_out
bash scripts/gpu_suite.sh
HG_DEBUG_TIMING=1 timeout 300 python scripts/tune_r2.py 16 256,0,0 > gpurun_out/exp_create.log 2>&1
grep "hg\]" gpurun_out/exp_create.log | head -12; tail -1 gpurun_out/exp_create.log | cut -c1-400
timeout 120 ./hydrograd.jl_b200/bulk_issue_micro > gpurun_out/bulk_issue.txt 2>&1; tail -3 gpurun_out/bulk_issue.txt
