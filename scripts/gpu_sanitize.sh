#!/bin/bash
# compute-sanitizer over small-mesh device tests (memcheck, then racecheck on the tile kernels' shared memory)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py::test_rhs_fixture_meshes tests/test_gpu_vjp.py::test_vjp_matches_bruteforce_oracle tests/test_gpu_zzy_fused_jvp.py::test_fused_chunk_of_directions tests/test_gpu_multirank.py::test_library_owned_exchange_matches_single_context -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitizer_memcheck.txt 2>&1
tail -5 gpurun_out/sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest "tests/test_gpu_parity.py::test_rhs_fixture_meshes" "tests/test_gpu_zzy_fused_jvp.py::test_fused_chunk_of_directions" -m gpu -q -x -p no:cacheprovider -k "savannah or ManningN" > gpurun_out/sanitizer_racecheck.txt 2>&1
tail -5 gpurun_out/sanitizer_racecheck.txt
