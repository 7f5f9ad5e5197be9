mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_final.err > gpurun_out/bench_final.json; tail -c 600 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/bench_ref_final.json; cat gpurun_out/bench_ref_final.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --sustained-s 0 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches_r1f.csv
python -c "import __graft_entry__ as g; g.smoke()"
