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
        # the other resident steppers loop the RHS too (four exchanges per RK4 step, one or two per AB3 step)
        single.set_state(Q0); single.step_rk4(1e-3, 3); ref_rk = single.get_state()
        ctx.set_state(info["Q"]); ctx.step_rk4(1e-3, 3)
        out["ipc_rk4_bitwise"] = bool(np.array_equal(ctx.get_state(), pick(ref_rk)))
        single.set_state(Q0); single.step_ab3(1e-3, 4); ref_ab = single.get_state()
        ctx.set_state(info["Q"]); ctx.step_ab3(1e-3, 4)
        out["ipc_ab3_bitwise"] = bool(np.array_equal(ctx.get_state(), pick(ref_ab)))
        # adaptive Tsit5: the stages exchange halos inside the call; the error norm is the one global scalar (host all-reduce).
        # Its sum runs in another order than on one context, so the steps agree to rounding, not bitwise
        P.attach_allreduce(ctx)
        single.set_state(Q0); _, st1 = single.solve_tsit5(0.0, 0.05, 1e-3, adaptive=True, abstol=1e-8, reltol=1e-6); ref_ts = single.get_state()
        ctx.set_state(info["Q"]); _, st2 = ctx.solve_tsit5(0.0, 0.05, 1e-3, adaptive=True, abstol=1e-8, reltol=1e-6)
        out["ipc_tsit5_counts"] = [st1["accepted"], st1["rejected"], st2["accepted"], st2["rejected"]]
        out["ipc_tsit5_err"] = float(np.abs(ctx.get_state() - pick(ref_ts)).max() / np.abs(ref_ts).max())
        single.set_state(Q0); single.solve_tsit5(0.0, 4e-3, 1e-3, adaptive=False); ref_tf = single.get_state()
        ctx.set_state(info["Q"]); ctx.solve_tsit5(0.0, 4e-3, 1e-3, adaptive=False)
        out["ipc_tsit5_fixed_bitwise"] = bool(np.array_equal(ctx.get_state(), pick(ref_tf)))
        # time adjoints: every RHS / VJP of the forward and reverse sweeps exchanges state / cotangent halos; Q0bar covers the
        # owned cells, pbar is the rank's partial sum
        par = dict(params=S.RIVER_N_ZONES.copy(), active="ManningN")
        def summed(v):
            t = torch.tensor(v, dtype=torch.float64, device="cuda" if one_gpu_each else "cpu"); dist.all_reduce(t); return t.cpu().numpy()
        for name, run in (("euler", lambda c, q, l: c.euler_adjoint(q, l, 1e-3, 6, **par)),
                          ("rk4", lambda c, q, l: c.rk_adjoint("RK4", q, l, 1e-3, 3, **par)),
                          ("tsit5", lambda c, q, l: c.rk_adjoint("Tsit5", q, l, 1e-3, 3, **par))):
            rQT, rbar, rp = run(single, Q0, lam)
            gQT, gbar, gp = run(ctx, info["Q"], pick(lam))
            out[f"adj_{name}_state_bitwise"] = bool(np.array_equal(gQT, pick(rQT)))
            out[f"adj_{name}_q0bar_err"] = float(np.abs(gbar - pick(rbar)).max() / np.abs(rbar).max())
            if rp.size:
                out[f"adj_{name}_pbar_err"] = float(np.abs(summed(gp) - rp).max() / max(np.abs(rp).max(), 1e-300))
        dist.barrier()
        ctx.comm_disconnect()
        # ---- host-buffer calls on a mesh large enough for the three-stream pipeline (>= 1M cells per rank): hg_rhs / hg_rhs_vjp
        # push the cut cells once the last chunk has landed; the band tiles run in the last stage
        bflat, bQ = S.river(2100, 1000)
        bN = bflat["n_cells"]
        bpart = (np.arange(bN) * world // bN).astype(np.int32)
        blam = np.random.default_rng(5).standard_normal(3 * bN)
        big = hg.Context(bflat, device=dev)
        bref = big.rhs(bQ)
        bbar, _ = big.rhs_vjp(bQ, blam)
        del big
        bloc, binfo = P.extract_local(bflat, bpart, rank, bQ)
        bown = binfo["own"]
        bpick = lambda v: np.concatenate([v[k * bN + bown] for k in range(3)])
        bctx = hg.Context(bloc, device=dev)
        # this context is wired by the library's own rendezvous (POSIX shared memory): what a host without a messaging layer uses
        bctx.comm_init_shm("t" + os.path.basename(os.environ["HG_WORKER_OUT"]) + os.environ.get("MASTER_PORT", "0"), rank, world, binfo["neighbors"])
        out["pipe_rhs_bitwise"] = bool(np.array_equal(bctx.rhs(binfo["Q"]), bpick(bref)))
        got_bar, _ = bctx.rhs_vjp(binfo["Q"], bpick(blam))
        out["pipe_vjp_err"] = float(np.abs(got_bar - bpick(bbar)).max() / np.abs(bbar).max())
        # the pullback of a forward call: Q = None reuses the state hg_rhs uploaded (only lambda crosses PCIe) -- same bits
        bctx.rhs(binfo["Q"])
        got_bar2, _ = bctx.rhs_vjp(None, bpick(blam))
        out["pipe_vjp_resident_state_bitwise"] = bool(np.array_equal(got_bar2, got_bar))
        dist.barrier()
        bctx.comm_disconnect()
        del bctx
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
    with open(os.path.join(os.environ["HG_WORKER_OUT"], f"rank{rank}.json"), "w") as fh:   # (stdout of the ranks interleaves)
        json.dump(out, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
