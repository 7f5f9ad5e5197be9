#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -6 gpurun_out/gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ("value","ms_per_step","rhs","vjp")}, d["sustained"], d["roofline"]["frac"], d["roofline_vjp"]["frac"], d["e2e"]["value"], d["c2"]["value"], d["c5"]["value"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vjp.csv python scripts/vjp_launches.py 16 2>&1 | tail -1
