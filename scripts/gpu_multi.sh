set -x
mkdir -p gpurun_out
N=${1:-2}
CM=${2:-4}
python bench.py --steps 10 --warmup 3 --cells-m $CM --no-cpu --e2e-steps 1 2>gpurun_out/b1.err | tee gpurun_out/b1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['rhs'], d['vjp'], d['value'], d['e2e']['value'])"
tail -3 gpurun_out/b1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --cells-m $CM --no-cpu --e2e-steps 1 2>gpurun_out/bN.err | tee gpurun_out/bN.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', d['rhs'], d['vjp'], d['value'], d['e2e']['value'])"
tail -8 gpurun_out/bN.err
