#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -6 gpurun_out/gpu_full.log
for T in ipc nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --transport $T --no-e2e > gpurun_out/bench_n2_$T.json 2> gpurun_out/bench_n2_$T.err
tail -3 gpurun_out/bench_n2_$T.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n2_$T.json'))
print("$T", {k:d[k] for k in ("value","ms_per_step","rhs","vjp")}, d["config"]["transport"], d.get("strong"))
PY
done
