#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ("value","ms_per_step","rhs","vjp")}, d["config"]["transport"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("strong"))
PY
