mkdir -p gpurun_out
free -g | head -2; nproc
for N in ${@:-4 8}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --e2e-steps 1 2>gpurun_out/scale_n$N.err | tee gpurun_out/scale_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', d['rhs'], d['vjp'], 'value', d['value'], 'e2e', d['e2e']['value'])"
grep -v "^\[rank\|OMP\|\*\*\*\|^$" gpurun_out/scale_n$N.err | tail -5
done
