set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for cfg in "256 192" "256 160" "192 128" "192 160" "512 384"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --cells-m 4 --no-cpu --e2e-steps 1 --tile $1 --threads $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TILE', d['config']['tile_cells'], '$2', 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'value', d['value'])"
done
