set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for cfg in "256 128" "256 192" "256 256" "512 256" "512 384" "128 128"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --cells-m 4 --no-cpu --e2e-steps 1 --tile $1 --threads $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TILE', d['config']['tile_cells'], '$2', 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'value', d['value'])"
done
