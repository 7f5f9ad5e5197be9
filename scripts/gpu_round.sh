mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
python bench.py 2>gpurun_out/bench_n1.err | tee gpurun_out/bench_n1.json | cut -c1-400
tail -2 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --cells-m 16 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
