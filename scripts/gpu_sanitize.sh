mkdir -p gpurun_out
SEL='test_rhs_synthetic_meshes or test_tiling_and_ordering or test_vjp_adjoint_identity_synthetic or test_vjp_launch_shapes_agree or test_l2_prefetch or test_variable_manning_on_tiled or test_partitioned_rhs_and_vjp_match_single_context_bitwise'
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "test_tiling_and_ordering or test_vjp_launch_shapes_agree" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
# forward mode + sensitivity solve (hg_jvp.cu), added after the round-1 sanitizer run
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zzy_jvp.py -q -x -k "matches_the_oracle or transpose or error_behaviour" > gpurun_out/sanitizer_memcheck_jvp.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_jvp.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_zzy_jvp.py -q -x -k "Q-oneD_bump or error_behaviour" > gpurun_out/sanitizer_racecheck_jvp.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_jvp.log
