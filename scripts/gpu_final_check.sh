#!/bin/bash
# The whole device suite with its printed figures kept (profiles/roundN_gpu_tests_full.log is a copy of gpurun_out/gpu_full.log),
# the smoke entry, then the default bench line (N = 1): the last call of a round.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider --tb=short --durations=10 > gpurun_out/gpu_full.log 2>&1
echo "exit $?" >> gpurun_out/gpu_full.log
tail -6 gpurun_out/gpu_full.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 170 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1_final.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline_vjp"]["frac"])
    print(d["e2e"])
except Exception as e:
    print("no bench line:", e)
PY
