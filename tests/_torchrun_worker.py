"""Worker of tests/test_gpu_torchrun.py: one process per rank under torch.distributed.run.  Every rank builds the same
small mesh, keeps its RCB / slab part, and checks its own cells of the partitioned RHS / VJP / Euler steps against a
single-context run of the whole mesh on its device -- through the library-owned transport (CUDA IPC peer stores,
hg_comm.cu) and, with one GPU per rank, through the NCCL send/recv path as well."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import _pkg
    hg = _pkg.load()
    from hydrograd_jl_b200 import parallel as P
    from hydrograd_jl_b200 import synthetic as S
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = torch.cuda.device_count()
    one_gpu_each = ndev >= world
    dev = rank % ndev
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if one_gpu_each else "gloo", rank=rank, world_size=world)
    out = {"rank": rank, "devices": ndev, "one_gpu_each": one_gpu_each}
    try:
        flat, Q0 = S.river(96, 32)
        N = flat["n_cells"]
        cx, cy = flat["cell_centroids"][:N], flat["cell_centroids"][N:]
        part = P.rcb_partition(cx, cy, world, keep_together=P.inlet_cell_groups(flat))
        lam = np.random.default_rng(4).standard_normal(3 * N)
        single = hg.Context(flat, device=dev, tile_cells=128)
        ref = single.rhs(Q0)
        ref_bar, _ = single.rhs_vjp(Q0, lam)
        single.set_state(Q0)
        single.step_euler(1e-3, 5)
        ref_Q5 = single.get_state()
        loc, info = P.extract_local(flat, part, rank, Q0)
        own, n = info["own"], info["own"].size
        pick = lambda v: np.concatenate([v[k * N + own] for k in range(3)])
        # ---- library-owned transport
        ctx = hg.Context(loc, device=dev, tile_cells=128)
        P.connect_ranks(ctx, info)
        ctx.set_state(info["Q"]); ctx.set_lambda(pick(lam))
        ctx.rhs_resident()
        out["ipc_rhs_bitwise"] = bool(np.array_equal(ctx.get_rhs(), pick(ref)))
        ctx.vjp_resident()
        bar = ctx.get_vjp()[0]
        out["ipc_vjp_err"] = float(np.abs(bar - pick(ref_bar)).max() / np.abs(ref_bar).max())
        ctx.step_euler(1e-3, 5)          # five exchanges inside one call
        out["ipc_euler_bitwise"] = bool(np.array_equal(ctx.get_state(), pick(ref_Q5)))
        dist.barrier()
        ctx.comm_disconnect()
        # ---- NCCL send/recv path (needs one GPU per rank)
        if one_gpu_each:
            ctx2 = hg.Context(loc, device=dev, tile_cells=128)
            ex = P.attach_exchanger(ctx2, info["neighbors"])
            ctx2.set_state(info["Q"]); ctx2.set_lambda(pick(lam))
            ctx2.halo_pack(True); ex.exchange(with_lambda=True)
            ctx2.rhs_resident()
            out["nccl_rhs_bitwise"] = bool(np.array_equal(ctx2.get_rhs(), pick(ref)))
            ctx2.vjp_resident()
            out["nccl_vjp_err"] = float(np.abs(ctx2.get_vjp()[0] - pick(ref_bar)).max() / np.abs(ref_bar).max())
            torch.cuda.synchronize()
        out["ok"] = True
    except Exception as e:  # noqa: BLE001
        out["ok"] = False
        out["error"] = repr(e)
    print("WORKER " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
