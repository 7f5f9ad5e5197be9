mkdir -p gpurun_out
N=${1:-2}
for mode in "" "--no-overlap"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e $mode 2>gpurun_out/ov.err | tee gpurun_out/overlap_n${N}${mode}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N $mode', d['rhs'], d['vjp'], 'value', d['value'])"
grep -i "error\|Traceback" -A5 gpurun_out/ov.err | head -20
done
