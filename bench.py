#!/usr/bin/env python
"""bench.py -- throughput of the fp64 2-D shallow-water RHS hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells-m M]

A "step" is one pass of the hot path over the whole mesh: one fused RHS evaluation (swe_2d_rhs,
semi_discretize_swe_2D.jl:18-277) of the resident state -- plus, once the adjoint kernel is enabled
(--vjp), one VJP of the same call.  Workload at every N: config C3, the synthetic ~16M-cell
meandering river per GPU (6 Manning zones, inlet-Q / exit-H / walls), i.e. WEAK scaling: rank r owns
slab r of a river N times as long.  All inputs (5.6 GB of state + mesh tables) are far larger than the
126 MB L2, so no flush is needed between iterations (config.l2 says so).

Printed JSON keys follow the driver contract; see DESIGN.md section "Measurement" for the arithmetic:
  value      cell-updates/s, inputs resident in HBM, CUDA-event time of K steps (max over ranks)
  e2e        same metric through the host-buffer C-ABI call hg_rhs (pinned host Q in, dQdt out,
             H2D + D2H inside the timed region)
  roofline   algorithmic bytes (100 N + 32 F + 4 sum_nF) / measured kernel time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++ restatement of the reference, NOT Hydrograd.jl itself -- Julia is absent)
             on a bounded sample of the same workload on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell_updates_per_sec"
UNIT = "cell-updates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~1 kHz; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                if nv is not None:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    flags = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if r & m else "Not Active" for _, m in flags])
                    self._stop_evt.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_traffic(kernel, N):
    """DRAM bytes per launch from the committed `ncu --set full` capture of the same kernel (profiles/), scaled per cell."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    return None if t is None else t["dram_bytes_per_cell"] * N


def algorithmic_bytes(N, F, sum_nf):
    """compulsory traffic with every array touched once, int32 indices, fp64 data."""
    return 100 * N + 32 * F + 4 * sum_nf


class CpuSample:
    """Oracle on a bounded sample of the C3 workload (a ~cells_m-million-cell slab of the same river), built once.
    One CPU 'step' mirrors the GPU step: one RHS + one derivative pass.  The reference differentiates its RHS with
    ForwardDiff / Zygote; the cheapest such pass is ONE forward-mode (dual-number) sweep, i.e. one direction of the
    Jacobian -- the hand-written VJP delivers all directions at once, so this is generous to the CPU side."""

    def __init__(self, cells_m=2.0, threads=0, seed=1234):
        from hydrograd_jl_b200 import synthetic as S
        from oracle.oracle import Oracle
        self.ni = max(64, int(cells_m * 1e6 / 1.1 / 1000))
        flat, self.Q0 = S.river(self.ni, 1000, seed=seed)
        self.o = Oracle(flat)
        self.n = flat["n_cells"]
        self.threads = threads or self.o.max_threads()
        self.v = np.ones_like(self.Q0)
        self.o.rhs(self.Q0, nthreads=self.threads)          # first touch (both passes: the CPU side is timed warm)
        self.o.jvp(self.Q0, self.v, nthreads=self.threads)

    def step(self, reps=1):
        """(seconds per RHS, seconds per derivative pass), mean of reps"""
        t = time.perf_counter()
        for _ in range(reps):
            self.o.rhs(self.Q0, nthreads=self.threads)
        dt_rhs = (time.perf_counter() - t) / reps
        t = time.perf_counter()
        for _ in range(reps):
            self.o.jvp(self.Q0, self.v, nthreads=self.threads)
        return dt_rhs, (time.perf_counter() - t) / reps

    def describe(self, reps):
        return (f"{reps} x (RHS + one forward-mode dual-number derivative pass) on a {self.n}-cell slab of the C3 river "
                f"({self.ni}x1000 quads), OpenMP over cells on {self.threads} threads")


def cpu_sample(cells_m=2.0, reps=3, threads=0, seed=1234):
    c = CpuSample(cells_m, threads, seed)
    dt_rhs, dt_jvp = c.step(reps)
    return {"value": c.n / (dt_rhs + dt_jvp), "rhs_only": c.n / dt_rhs, "cores": c.threads, "sample": c.describe(reps)}


def workload_text(cells_millions):
    """config.workload, shared by both arms (the reference arm times a bounded sample of this workload)"""
    return (f"C3 synthetic {cells_millions:.1f}M-cell meandering river per GPU (mixed tri/quad, 6 Manning "
            "zones, inlet-Q/exit-H/walls); step = one fused fp64 RHS + one hand-written VJP of the resident state")


def run_reference(args):
    """--impl reference: the reference's CPU path.  Hydrograd.jl is Julia-only and Julia is not in this image, so
    this times the oracle port (C++ restatement, reference evaluation order) on all host cores."""
    import _pkg
    _pkg.load()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CpuSample(cells_m=1.0)
    t_rhs, t_all = [], []
    for s in range(args.warmup + args.steps):
        dt_rhs, dt_jvp = c.step(1)
        if s >= args.warmup:
            t_rhs.append(dt_rhs)
            t_all.append(dt_rhs + dt_jvp)
    step_s = float(np.mean(t_all))
    val, val_rhs = c.n / step_s, c.n / float(np.mean(t_rhs))
    cores, sample = c.threads, c.describe(1) + f", per step; {args.steps} timed steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args.cells_m), "sample": sample,
                       "note": "CPU arm: one RHS + one forward-mode derivative pass of the C++ port of the reference algorithm "
                               "(the cheapest derivative pass the reference's AD performs), bounded sample of the same mesh family"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "rhs_only": val_rhs},
            "rhs": {"value": val_rhs, "unit": UNIT},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells-m", type=float, default=16.0, help="million cells per GPU")
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--threads", type=int, default=0, help="threads per CTA of the fused kernel (tuning)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="multi-GPU: run the tiles without halo faces while the halo is exchanged on a second stream "
                    "(hg_*_resident_phase).  Measured slower at 16M cells/GPU (exchange ~25 us; two launches + NCCL beside the kernel cost more): off by default")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` key (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs: the launch list then holds the timed region only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE line (the JSON record of rank 0): everything else any library prints while we
    # run (NCCL's version banner, warnings) is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import _pkg
    hg = _pkg.load()
    from hydrograd_jl_b200 import synthetic as S

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")

    # ---- workload: slab `rank` of a river `world` times as long (weak scaling).  The slab decomposition is what
    # recursive coordinate bisection yields for this elongated domain; each rank generates its slab plus one
    # column on each cut side and extracts its local mesh (owned cells + halo boundaries) with the general
    # partitioner code (hydrograd.jl_b200/parallel.py).
    from hydrograd_jl_b200 import parallel as PAR
    ni = int(args.cells_m * 1e6 / 1.1 / 1000)
    t0 = time.time()
    if world == 1:
        flat, Q0 = S.river(ni, 1000)
        ranks = []
    else:
        lo = rank * ni - (1 if rank > 0 else 0)
        hi = (rank + 1) * ni + (1 if rank < world - 1 else 0)
        gflat, gQ = S.river(hi - lo, 1000, i0=lo, ni_total=world * ni)
        # owner of every cell of the extended slab, from the stream-wise column of its centroid's quad
        col = S.cell_columns(gflat, hi - lo, 1000)
        part = np.full(gflat["n_cells"], rank, dtype=np.int32)
        if rank > 0:
            part[col == 0] = rank - 1
        if rank < world - 1:
            part[col == (hi - lo - 1)] = rank + 1
        flat, info = PAR.extract_local(gflat, part, rank, gQ)
        Q0, ranks = info["Q"], info["neighbors"]
        del gflat, gQ
    N, F = flat["n_cells"], flat["n_faces"]
    log(f"[rank {rank}] mesh: N={N} F={F} ({time.time() - t0:.1f}s)")
    t0 = time.time()
    ctx = hg.Context(flat, device=local, tile_cells=args.tile, threads=args.threads)
    st = ctx.mesh_stats()
    log(f"[rank {rank}] context: {st} ({time.time() - t0:.1f}s)")
    ctx.set_state(Q0)
    rng = np.random.default_rng(99 + rank)
    ctx.set_lambda(rng.standard_normal(3 * N))
    # one explicit (non-default) stream for everything: pack kernel, NCCL transfers, RHS / VJP kernels, timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ex = PAR.attach_exchanger(ctx, ranks) if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # multi-GPU step: pack the halo, exchange it on a second stream (NCCL send/recv over NVLink) WHILE the tiles without halo
    # faces run, then the band of tiles with halo faces once the received block is complete
    comm = torch.cuda.Stream() if ex is not None else None
    ev_pack, ev_recv = torch.cuda.Event(), torch.cuda.Event()

    def overlapped(pack_lambda, run):
        ctx.halo_pack(pack_lambda)
        ev_pack.record(stream)
        comm.wait_event(ev_pack)
        with torch.cuda.stream(comm):
            ex.exchange(pack_lambda)
            ev_recv.record(comm)
        run(1)
        stream.wait_event(ev_recv)
        run(2)

    def rhs_step():
        if ex is None:
            ctx.rhs_resident()
        elif not args.overlap:
            ctx.halo_pack(False)
            ex.exchange(False)
            ctx.rhs_resident()
        else:
            overlapped(False, ctx.rhs_resident)

    def vjp_step():
        if ex is None:
            ctx.vjp_resident()
        elif not args.overlap:
            ctx.halo_pack(True)
            ex.exchange(True)
            ctx.vjp_resident()
        else:
            overlapped(True, ctx.vjp_resident)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def allmax(x):
        if dist is None:
            return x
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- warm-up, then exactly K timed steps; CUDA events on the stream every kernel and transfer is ordered on
    W = max(args.warmup, 3)
    timed(rhs_step, W); timed(vjp_step, W)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = ctx.kernel_launches()
    ms = timed(rhs_step, args.steps)
    barrier()
    ms_vjp = timed(vjp_step, args.steps)
    barrier()
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop()
    ctx.sync()
    ms, ms_vjp = allmax(ms), allmax(ms_vjp)
    if dist is not None:
        tn = torch.tensor([N], device="cuda", dtype=torch.float64)
        dist.all_reduce(tn)
        N_total = int(tn.item())
    else:
        N_total = N
    ms_rhs_step = ms / args.steps
    ms_vjp_step = ms_vjp / args.steps
    ms_per_step = ms_rhs_step + ms_vjp_step
    value = N_total / (ms_per_step * 1e-3)

    # ---- sustained regime (single GPU): ~2 s of alternating steps back to back.  The K timed steps above last a few tens
    # of milliseconds -- the same "burst" regime the HBM peak in MEASURED_PEAKS.json was measured in (best of 10 copies);
    # under seconds of load this board hits its power cap (sw_power_cap) and both kernels slow down by a few per cent.
    sustained = None
    if world == 1 and args.sustained_s > 0:
        n_pairs = max(20, int(args.sustained_s * 1e3 / ms_per_step))
        s2 = ClockSampler(local)
        s2.start()
        t_r = t_v = 0.0
        for _ in range(n_pairs // 20):
            t_r += timed(rhs_step, 20)
            t_v += timed(vjp_step, 20)
        n_done = (n_pairs // 20) * 20
        c2 = s2.stop()
        sustained = {"rhs_ms": t_r / n_done, "vjp_ms": t_v / n_done, "value": N_total / ((t_r + t_v) / n_done * 1e-3), "unit": UNIT,
                     "steps": n_done, "clocks": c2}

    # ---- roofline of the dominant kernel (k_fused_rhs): algorithmic bytes / measured launch time
    peak, peak_src = measured_peak()
    abytes = algorithmic_bytes(N, F, st["sum_cell_faces"])
    achieved = abytes / (ms_rhs_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": measured_traffic("k_fused_rhs", N), "kernel": "k_fused_rhs", "algorithmic_bytes_per_launch": abytes,
                "bytes_per_cell": abytes / N, "peak_source": peak_src, "ms_per_launch": ms_rhs_step}
    vbytes = abytes + 32 * N          # RHS inputs re-read + lambda (24 B) + Qbar (24 B) + nbar (8 B) - dQ (24 B)
    vach = vbytes / (ms_vjp_step * 1e-3) / 1e9
    roofline_vjp = {"bound": "hbm", "achieved": vach, "peak": peak, "unit": "GB/s", "frac": vach / peak, "traffic": measured_traffic("k_fused_vjp", N),
                    "kernel": "k_fused_vjp", "algorithmic_bytes_per_launch": vbytes, "bytes_per_cell": vbytes / N,
                    "ms_per_launch": ms_vjp_step}

    # ---- end to end through the host-buffer ABI (pinned host memory; H2D + D2H inside the timed region): at N = 1
    # this is exactly hg_rhs; at N > 1 the same three stages with the halo exchange in between
    e2e = None
    if not args.no_e2e:
        hQ = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hD = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hL = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hQ.numpy()[:] = Q0
        hL.numpy()[:] = 1.0
        out, outb = hD.numpy(), hB.numpy()

        def e2e_rhs():
            if ex is None:
                ctx.rhs(hQ.numpy(), out=out)
            else:
                ctx.set_state(hQ.numpy())
                rhs_step()
                ctx.get_rhs(out=out)

        def e2e_vjp():
            if ex is None:
                ctx.rhs_vjp_into(hQ.numpy(), hL.numpy(), outb)
            else:
                ctx.set_state(hQ.numpy())
                ctx.set_lambda(hL.numpy())
                vjp_step()
                ctx.get_vjp_into(outb)

        e2e_rhs(); e2e_vjp()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_rhs()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_vjp()
        barrier()
        t2 = time.perf_counter()
        e2e_rhs_s = allmax((t1 - t0) / args.e2e_steps)
        e2e_vjp_s = allmax((t2 - t1) / args.e2e_steps)
        e2e_s = e2e_rhs_s + e2e_vjp_s
        e2e = {"value": N_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 72 * N, "d2h_bytes_per_step": 48 * N,
               "ms_per_step": e2e_s * 1e3, "rhs_ms": e2e_rhs_s * 1e3, "vjp_ms": e2e_vjp_s * 1e3,
               "rhs_only": N_total / e2e_rhs_s,
               "what": "one RHS + one VJP through pinned host buffers (hg_rhs: H2D state, kernel, D2H dQdt; hg_rhs_vjp: H2D state "
                       "and lambda, kernel, D2H Qbar), chunked over three streams"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_sample()
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"], "rhs_only": c["rhs_only"],
               "note": "C++ oracle port of the reference algorithm (OpenMP); Hydrograd.jl itself cannot run here (no Julia)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(N / 1e6),
                           "cells_per_gpu": N, "faces_per_gpu": F, "tile_cells": args.tile, "n_tiles": st["n_tiles"],
                           "l2": "inputs (state + mesh tables >> 126 MB L2) larger than L2, no flush needed",
                           "parallelism": (f"rcb-slab x{world}, one-layer halo, NCCL send/recv per step" + (" overlapped with the tiles without halo faces" if args.overlap else "")) if world > 1 else "single GPU"},
                "roofline": roofline, "roofline_vjp": roofline_vjp,
                "rhs": {"value": N_total / (ms_rhs_step * 1e-3), "unit": UNIT, "ms": ms_rhs_step},
                "vjp": {"value": N_total / (ms_vjp_step * 1e-3), "unit": UNIT, "ms": ms_vjp_step},
                "sustained": sustained, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
