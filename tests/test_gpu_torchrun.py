"""The multi-rank path under a REAL multi-process launch (torch.distributed.run, world size 2): partitioned RHS / VJP /
Euler steps of every rank equal the single-context bits -- through the library-owned CUDA-IPC transport and, when the box
has a GPU per rank, through NCCL send/recv too.  On a one-GPU box both ranks share the device (IPC still applies; NCCL
cannot run two ranks on one device and is skipped)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_single_context(tmp_path):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_torchrun_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="4", HG_WORKER_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    res = [json.load(open(tmp_path / f)) for f in sorted(os.listdir(tmp_path)) if f.startswith("rank")]
    assert r.returncode == 0 and len(res) == 2, r.stdout[-2000:] + r.stderr[-3000:]
    print(res)
    for o in res:
        assert o["ok"], o
        assert o["ipc_rhs_bitwise"] and o["ipc_euler_bitwise"] and o["ipc_rk4_bitwise"] and o["ipc_ab3_bitwise"], o
        assert o["ipc_vjp_err"] <= 1e-13, o
        a1, r1, a2, r2 = o["ipc_tsit5_counts"]
        assert (a1, r1) == (a2, r2) and a1 > 3 and o["ipc_tsit5_err"] <= 1e-11 and o["ipc_tsit5_fixed_bitwise"], o
        for name in ("euler", "rk4", "tsit5"):
            assert o[f"adj_{name}_state_bitwise"] and o[f"adj_{name}_q0bar_err"] <= 1e-12 and o[f"adj_{name}_pbar_err"] <= 1e-12, o
        assert o["pipe_rhs_bitwise"] and o["pipe_vjp_err"] <= 1e-13, o
        assert o["pipe_vjp_resident_state_bitwise"], o
        if o["one_gpu_each"]:
            assert o["nccl_rhs_bitwise"] and o["nccl_vjp_err"] <= 1e-13, o
