mkdir -p gpurun_out
SEL='test_rhs_synthetic_meshes or test_tiling_and_ordering or test_vjp_adjoint_identity_synthetic or test_vjp_launch_shapes_agree or test_l2_prefetch or test_variable_manning_on_tiled or test_partitioned_rhs_and_vjp_match_single_context_bitwise'
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "test_tiling_and_ordering or test_vjp_launch_shapes_agree" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
