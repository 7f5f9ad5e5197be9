"""bench.py's reference arm (`--impl reference`): runs on the host cores without a GPU, prints ONE JSON line with the
contract's keys; under a launcher only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_sec" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["steps"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    # the launcher pins OMP_NUM_THREADS=1; the arm takes its threads from the affinity mask instead (round-1 review)
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "OMP_NUM_THREADS": "1"}) == []
    d = json.loads(_run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})[0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
