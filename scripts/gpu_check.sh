set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for T in 128 256 512 1024; do
python bench.py --steps 20 --warmup 3 --cells-m 4 --no-cpu --e2e-steps 1 --tile $T 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TILE', d['config']['tile_cells'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'value', d['value'])"
done
