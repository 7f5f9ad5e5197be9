#!/bin/bash
# bench.py on N GPUs of this box, the way the driver launches it: bash scripts/gpu_bench_n.sh N [extra bench flags]
mkdir -p gpurun_out
N=${1:-1}; shift
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
tail -3 gpurun_out/bench_n$N.err | cut -c1-200
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ("value","ms_per_step","rhs","vjp")}, d["config"]["transport"], d["roofline"]["frac"], d["roofline_vjp"]["frac"], (d.get("e2e") or {}).get("value"), d.get("strong"), d.get("sustained"))
PY
