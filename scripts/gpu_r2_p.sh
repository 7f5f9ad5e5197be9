#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_torchrun.py tests/test_gpu_multirank.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e > gpurun_out/bench_n2_ipc.json 2> gpurun_out/bench_n2_ipc.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n2_ipc.json'))
print({k:d[k] for k in ("value","ms_per_step","rhs","vjp")}, d["config"]["transport"], d.get("strong"))
PY
