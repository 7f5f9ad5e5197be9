#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box variation is +-4 %): hydrograd.jl_b200/libhg_rhsA.so (reference build)
# against the in-tree build, alternating; then the phase clocks of libhg_clk.so (HG_NVCC_EXTRA=-DHG_PHASE_CLOCKS) if present.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vjp.py tests/test_gpu_multirank.py tests/test_gpu_ensemble.py -m gpu -x -q -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1
tail -3 gpurun_out/exp_tests.log | cut -c1-300
cp hydrograd.jl_b200/libhydrograd_b200.so /tmp/libB.so
for v in B A B A; do
  if [ $v = A ]; then cp hydrograd.jl_b200/libhg_rhsA.so hydrograd.jl_b200/libhydrograd_b200.so; else cp /tmp/libB.so hydrograd.jl_b200/libhydrograd_b200.so; fi
  echo "variant $v"; timeout 300 python scripts/tune_r2.py 16 256,0,0 2>&1 | tail -1 | cut -c1-400
done | tee gpurun_out/exp_ab.log
if [ -f hydrograd.jl_b200/libhg_clk.so ]; then
  cp hydrograd.jl_b200/libhg_clk.so hydrograd.jl_b200/libhydrograd_b200.so
  timeout 400 python scripts/phase_clocks.py 16 256 > gpurun_out/exp_clocks.log 2>&1
  tail -3 gpurun_out/exp_clocks.log | cut -c1-300
fi
