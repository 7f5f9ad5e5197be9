#!/usr/bin/env python
"""bench.py -- throughput of the fp64 2-D shallow-water RHS hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells-m M]

A "step" is one pass of the hot path over the whole mesh: one fused RHS evaluation (swe_2d_rhs,
semi_discretize_swe_2D.jl:18-277) of the resident state plus one hand-written VJP of the same call with the Manning
zone values as the active parameter (BASELINE config C4: Qbar AND pbar are inside the step).  Workload at every N:
config C3, the synthetic ~16M-cell meandering river per GPU (6 Manning zones, inlet-Q / exit-H / walls), i.e. WEAK
scaling: rank r owns slab r of a river N times as long.  All inputs (5.6 GB of state + mesh tables) are far larger
than the 126 MB L2, so no flush is needed between iterations (config.l2 says so).  At N > 1 the halo moves through the
library's own transport (hg_comm.cu: peer stores over NVLink, consumed inside the one tile-kernel launch);
--transport nccl selects the round-1 pack -> NCCL send/recv -> kernel sequence for comparison.

Extra keys (same JSON line): `sustained` (seconds of back-to-back steps, with roofline fractions), `c2` (1M-cell dam
break, fused Euler steps) and `c5` (128-member parameter ensemble on a 1M-cell mesh) at N = 1; `strong` (the 16M-cell
river split over the N GPUs) at N > 1.

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


TRAFFIC_SOURCE = ("static: dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of this kernel on the 16M-cell "
                  "river (profiles/traffic.json), scaled per cell -- not measured in this run")


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


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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
        # torchrun exports OMP_NUM_THREADS=1 to every rank: ask the OS which cores this process may run on instead
        self.threads = threads or host_threads()
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


def side_configs(hg, S, device, peak):
    """BASELINE configs C2 and C5 on one GPU (extra keys of the N = 1 line).  Small meshes: built, timed and freed."""
    out = {}
    # ---- C2: 1M-cell dam break, forward run with the customized Euler stepper fused into the RHS kernel
    flat, Q0 = S.dam_break(953)
    N, F, sn = flat["n_cells"], flat["n_faces"], int(flat["cell_nfaces"].sum())
    ctx = hg.Context(flat, device=device)
    ctx.set_state(Q0)
    dt = 1e-4
    ctx.time_rhs(50, True, dt)
    ms = min(ctx.time_rhs(200, True, dt) / 200 for _ in range(3))
    ab = algorithmic_bytes(N, F, sn)
    out["c2"] = {"workload": f"C2 synthetic {N / 1e6:.2f}M-cell unstructured dam break (mixed tri/quad, walls), forward run: fused RHS + Euler update per step",
                 "cells": N, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "value": N / (ms * 1e-3), "unit": UNIT,
                 "roofline_frac": ab / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": ab,
                 "l2": f"working set {ab / 1e6:.0f} MB vs 126 MB L2: partly L2-resident, no flush between steps (steady state of a forward run)"}
    del ctx
    # ---- C5: parameter ensemble, 128 members per GPU (1024 over 8 GPUs: members shard with no communication)
    flat, Q0 = S.river(909, 1000)
    N = flat["n_cells"]
    M = 128
    ctx = hg.Context(flat, device=device)
    ctx.ensemble_alloc(M, per_member_manning=True)
    rng = np.random.default_rng(1234)
    for m in range(M):
        ctx.ensemble_set_member(m, Q0, S.RIVER_N_ZONES[:flat["n_mat"]] * (1 + 0.2 * rng.uniform(-1, 1, flat["n_mat"])), "ManningN")
    ctx.time_ensemble(3, dt)
    ms = min(ctx.time_ensemble(10, dt) / 10 for _ in range(2))
    bpmc = (48.0 * M + 132.0) / M          # state in + out per member, mesh / bed tables once per tile
    out["c5"] = {"workload": f"C5 ensemble of {M} Manning-n parameter sets on a {N / 1e6:.2f}M-cell river (per GPU; x8 GPUs = 1024 members), fused Euler step of all members per launch",
                 "members": M, "cells": N, "ms_per_step": ms, "value": M * N / (ms * 1e-3), "unit": "member-cell-updates/s",
                 "bytes_per_member_cell": bpmc, "roofline_frac": bpmc * M * N / (ms * 1e-3) / 1e9 / peak,
                 "note": "bound by the fp64 / shared-memory work of the tile kernel, not by HBM (the mesh blocks of a tile are shared by the members through L2)"}
    del ctx
    return out


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
    ap.add_argument("--transport", default="ipc", choices=["ipc", "nccl"],
                    help="multi-GPU halo transport: ipc = the library's own peer stores over NVLink (hg_comm.cu), nccl = pack -> NCCL send/recv -> kernel")
    ap.add_argument("--overlap", action="store_true", help="nccl transport only: run the tiles without halo faces while the halo is exchanged on a second stream "
                    "(hg_*_resident_phase); measured slower than the serial sequence in round 1")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` key (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs: the launch list then holds the timed region only)")
    ap.add_argument("--no-side", action="store_true", help="skip the c2 / c5 / strong extra keys")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE line (the JSON record of rank 0): everything else any library prints while we
    # run (NCCL's version banner, warnings) is sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; hg_create's mesh preprocessing is OpenMP (setup, not in
    # any timed region): give each rank its share of the host cores before the library's OpenMP runtime reads the variable
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))
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

    from hydrograd_jl_b200 import parallel as PAR
    peak, peak_src = measured_peak()
    p_zones = S.RIVER_N_ZONES.copy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def allsum(x):
        if dist is None:
            return x
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt)
        return float(tt.item())

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    # one explicit (non-default) stream for everything: halo push / pack kernel, NCCL transfers, RHS / VJP kernels, timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    class Slab:
        """Slab `rank` of a river of world * ni columns: mesh, context, halo transport, step functions."""

        def __init__(self, ni):
            t0 = time.time()
            if world == 1:
                flat, Q0 = S.river(ni, 1000)
                info = {"neighbors": [], "counts": []}
            else:
                # the slab decomposition is what recursive coordinate bisection yields for this elongated domain; each rank
                # generates its slab plus one column on each cut side and extracts its local mesh (owned cells + halo
                # boundaries) with the general partitioner code (hydrograd.jl_b200/parallel.py)
                lo = rank * ni - (1 if rank > 0 else 0)
                hi = (rank + 1) * ni + (1 if rank < world - 1 else 0)
                gflat, gQ = S.river(hi - lo, 1000, i0=lo, ni_total=world * ni)
                col = S.cell_columns(gflat, hi - lo, 1000)
                part = np.full(gflat["n_cells"], rank, dtype=np.int32)
                if rank > 0:
                    part[col == 0] = rank - 1
                if rank < world - 1:
                    part[col == (hi - lo - 1)] = rank + 1
                flat, info = PAR.extract_local(gflat, part, rank, gQ)
                Q0 = info["Q"]
                del gflat, gQ
            self.N, self.F, self.Q0, self.info = flat["n_cells"], flat["n_faces"], Q0, info
            log(f"[rank {rank}] mesh: N={self.N} F={self.F} ({time.time() - t0:.1f}s)")
            t0 = time.time()
            self.ctx = ctx = hg.Context(flat, device=local, tile_cells=args.tile, threads=args.threads)
            self.st = ctx.mesh_stats()
            self.create_s = time.time() - t0
            log(f"[rank {rank}] context: {self.st} ({self.create_s:.1f}s)")
            ctx.set_state(Q0)
            ctx.set_lambda(np.random.default_rng(99 + rank).standard_normal(3 * self.N))
            ctx.set_params(p_zones[:flat["n_mat"]], "ManningN")     # C4: the Manning-n gradient is part of the VJP step
            ctx.set_stream(stream.cuda_stream)
            self.transport, self.ex = "none", None
            if world > 1:
                self.transport = args.transport
                if self.transport == "ipc":
                    ok = 1.0
                    try:
                        PAR.connect_ranks(ctx, info)
                    except Exception as e:  # noqa: BLE001 -- e.g. CUDA IPC not permitted in this container
                        log(f"[rank {rank}] library-owned transport unavailable ({e}); falling back to NCCL send/recv")
                        ok = 0.0
                    if allsum(ok) < world:      # all ranks or none
                        if ok:
                            ctx.comm_disconnect()
                        self.transport = "nccl (ipc unavailable)"
                if self.transport != "ipc":
                    self.ex = PAR.attach_exchanger(ctx, info["neighbors"])
                    self.comm, self.ev_pack, self.ev_recv = torch.cuda.Stream(), torch.cuda.Event(), torch.cuda.Event()

        def _overlapped(self, pack_lambda, run):
            self.ctx.halo_pack(pack_lambda)
            self.ev_pack.record(stream)
            self.comm.wait_event(self.ev_pack)
            with torch.cuda.stream(self.comm):
                self.ex.exchange(pack_lambda)
                self.ev_recv.record(self.comm)
            run(1)
            stream.wait_event(self.ev_recv)
            run(2)

        def rhs_step(self):
            ctx = self.ctx
            if self.ex is None:
                ctx.rhs_resident()            # multi-rank: the halo push is part of the call (auto exchange)
            elif not args.overlap:
                ctx.halo_pack(False); self.ex.exchange(False); ctx.rhs_resident()
            else:
                self._overlapped(False, ctx.rhs_resident)

        def vjp_step(self):
            ctx = self.ctx
            if self.ex is None:
                ctx.vjp_resident()
            elif not args.overlap:
                ctx.halo_pack(True); self.ex.exchange(True); ctx.vjp_resident()
            else:
                self._overlapped(True, ctx.vjp_resident)

        def close(self):
            barrier()
            if self.transport == "ipc":
                self.ctx.comm_disconnect()
            self.ctx.close()

    ni = int(args.cells_m * 1e6 / 1.1 / 1000)
    sl = Slab(ni)
    ctx, N, F, st, Q0 = sl.ctx, sl.N, sl.F, sl.st, sl.Q0
    rhs_step, vjp_step = sl.rhs_step, sl.vjp_step

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
    N_total = int(allsum(float(N)))
    ms_rhs_step = ms / args.steps
    ms_vjp_step = ms_vjp / args.steps
    ms_per_step = ms_rhs_step + ms_vjp_step
    value = N_total / (ms_per_step * 1e-3)

    uncoupled = None

    abytes = algorithmic_bytes(N, F, st["sum_cell_faces"])
    vbytes = abytes + 32 * N          # RHS inputs re-read + lambda (24 B) + Qbar (24 B) + nbar (8 B) - dQ (24 B)

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
                     "rhs_roofline_frac": abytes / (t_r / n_done * 1e-3) / 1e9 / peak,
                     "vjp_roofline_frac": vbytes / (t_v / n_done * 1e-3) / 1e9 / peak,
                     "steps": n_done, "clocks": c2}

    # ---- forward mode (the reference's sensitivity / ForwardDiff paths): the fused forward-mode tile kernel, six directions
    # (one per Manning zone) per launch
    jvp = None
    if world == 1 and not args.no_side:
        Kd = 6
        ctx.time_jvp(Kd, 1)
        ms_j = min(ctx.time_jvp(Kd, 3) / 3 for _ in range(2))
        jbytes = abytes + Kd * 56 * N     # tile tables + state once per tile (the K CTAs of a tile share them through L2); per direction 24 B in, 24 B out, 8 B parameter tangent
        jvp = {"kernel": "k_fused_jvp", "directions": Kd, "ms_per_launch": ms_j, "value": Kd * N / (ms_j * 1e-3), "unit": "direction-cell-updates/s",
               "roofline_frac": jbytes / (ms_j * 1e-3) / 1e9 / peak, "note": "values + six tangents (dQ/dt and J_Q v + J_p e_k per Manning zone); bound by the fp64 pipe (dual-number Roe flux), not by HBM"}

    # ---- roofline of the dominant kernel (k_fused_rhs): algorithmic bytes / measured launch time
    achieved = abytes / (ms_rhs_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": measured_traffic("k_fused_rhs", N), "traffic_source": TRAFFIC_SOURCE, "kernel": "k_fused_rhs",
                "algorithmic_bytes_per_launch": abytes, "bytes_per_cell": abytes / N, "peak_source": peak_src, "ms_per_launch": ms_rhs_step}
    vach = vbytes / (ms_vjp_step * 1e-3) / 1e9
    roofline_vjp = {"bound": "hbm", "achieved": vach, "peak": peak, "unit": "GB/s", "frac": vach / peak, "traffic": measured_traffic("k_fused_vjp", N),
                    "traffic_source": TRAFFIC_SOURCE, "kernel": "k_fused_vjp + parameter follow-ups (active parameter ManningN: Qbar and pbar)",
                    "algorithmic_bytes_per_launch": vbytes, "bytes_per_cell": vbytes / N, "ms_per_launch": ms_vjp_step}

    # ---- end to end through the host-buffer ABI (pinned host memory; H2D + D2H inside the timed region): at N = 1
    # this is exactly hg_rhs / hg_rhs_vjp (three-stream pipeline); at N > 1 upload, halo push + kernel, download per rank
    e2e = None
    if not args.no_e2e:
        hQ = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hD = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hL = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hQ.numpy()[:] = Q0
        hL.numpy()[:] = 1.0
        out, outb = hD.numpy(), hB.numpy()
        pz = p_zones[:len(p_zones)]
        pbar_host = np.zeros(pz.size)

        host_api = world == 1 or sl.transport == "ipc"    # hg_rhs / hg_rhs_vjp with host pointers: the three-stream pipeline

        def e2e_rhs():
            if host_api:
                ctx.rhs(hQ.numpy(), pz, "ManningN", out=out)
            else:
                ctx.set_state(hQ.numpy())
                rhs_step()
                ctx.get_rhs(out=out)

        def e2e_vjp():
            # the pullback of the forward call just made (Zygote.pullback / the shim's rrule): the primal's state is still on
            # the device, so only the cotangent is uploaded (hg_rhs_vjp with Q = NULL)
            if host_api:
                ctx.rhs_vjp_into(None, hL.numpy(), outb, pz, "ManningN", pbar_host)
            else:
                ctx.set_lambda(hL.numpy())
                vjp_step()
                ctx.get_vjp_into(outb)

        e2e_rhs(); e2e_vjp()  # warm-up
        barrier()
        t_rhs = t_vjp = 0.0
        for _ in range(args.e2e_steps):          # one step = the forward call, then its pullback
            t0 = time.perf_counter()
            e2e_rhs()
            t1 = time.perf_counter()
            e2e_vjp()
            t2 = time.perf_counter()
            t_rhs += t1 - t0
            t_vjp += t2 - t1
        barrier()
        e2e_rhs_s = allmax(t_rhs / args.e2e_steps)
        e2e_vjp_s = allmax(t_vjp / args.e2e_steps)
        e2e_s = allmax((t_rhs + t_vjp) / args.e2e_steps)
        e2e = {"value": N_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 48 * N, "d2h_bytes_per_step": 48 * N,
               "ms_per_step": e2e_s * 1e3, "rhs_ms": e2e_rhs_s * 1e3, "vjp_ms": e2e_vjp_s * 1e3,
               "rhs_only": N_total / e2e_rhs_s,
               "pcie_gbs_per_gpu": {"rhs_h2d": 24 * N / e2e_rhs_s / 1e9, "rhs_d2h": 24 * N / e2e_rhs_s / 1e9,
                                    "vjp_h2d": 24 * N / e2e_vjp_s / 1e9, "vjp_d2h": 24 * N / e2e_vjp_s / 1e9,
                                    "note": "bytes of each direction over the whole call time (the directions overlap at N = 1)"},
               "what": ("one forward call + its pullback through pinned host buffers, as Zygote.pullback(swe_2d_rhs, Q, p) makes them "
                        "(hg_rhs: H2D state, kernel, D2H dQdt; hg_rhs_vjp with Q = NULL: the state of the forward call is still resident, "
                        "H2D lambda, kernel, D2H Qbar and pbar), chunked over three streams" +
                        ("; per rank, the cut cells are pushed to the neighbours once the last chunk has landed; all ranks share one host's PCIe / pinned memory" if world > 1 else ""))
                       if host_api else
                       "per rank: hg_set_state (H2D) -> pack + NCCL send/recv + tile kernel -> hg_get_rhs; hg_set_lambda (H2D) -> ... -> hg_get_vjp (D2H); all ranks share one host's PCIe / pinned memory"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_sample()
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port", "sample": c["sample"], "rhs_only": c["rhs_only"],
               "note": "C++ oracle port of the reference algorithm (OpenMP); Hydrograd.jl itself cannot run here (no Julia)"}

    # ---- extra records: C2 / C5 on one GPU; strong scaling (the 16M-cell river split over the GPUs) at N > 1
    side, strong = {}, None
    create_s = sl.create_s
    transport = sl.transport
    if not args.no_side:
        if world == 1:
            sl.close()
            del sl, ctx
            side = side_configs(hg, S, local, peak)
        else:
            # what the halo exchange costs ON THESE BOARDS: the same launches with the library's transport disconnected (no
            # push, no waiting; the halo faces read stale values -- a timing probe, not a result).  Boards differ by a few per
            # cent and a multi-GPU step is the max over ranks, so this ratio -- not the comparison with a single-GPU run on
            # another board -- isolates the cost of the exchange.
            if sl.transport == "ipc":
                barrier()
                ctx.comm_disconnect()
                timed(ctx.rhs_resident, W); timed(ctx.vjp_resident, W)
                barrier()
                u_r = timed(ctx.rhs_resident, args.steps)
                barrier()
                u_v = timed(ctx.vjp_resident, args.steps)
                u_ms = (allmax(u_r) + allmax(u_v)) / args.steps
                uncoupled = {"ms_per_step": u_ms, "same_board_efficiency": u_ms / ms_per_step,
                             "note": "identical work without the halo push / wait on the same GPUs (stale halos: timing probe only)"}
            sl.close()
            del sl, ctx
            s2 = Slab(max(8, ni // world))
            timed(s2.rhs_step, W); timed(s2.vjp_step, W)
            barrier()
            m_r = allmax(timed(s2.rhs_step, args.steps)) / args.steps
            barrier()
            m_v = allmax(timed(s2.vjp_step, args.steps)) / args.steps
            Ns_total = int(allsum(float(s2.N)))
            strong = {"workload": f"the {Ns_total / 1e6:.1f}M-cell C3 river split into {world} slabs ({s2.N / 1e6:.2f}M cells per GPU)",
                      "cells_total": Ns_total, "rhs_ms": m_r, "vjp_ms": m_v, "ms_per_step": m_r + m_v,
                      "value": Ns_total / ((m_r + m_v) * 1e-3), "unit": UNIT, "transport": s2.transport,
                      "note": "strong-scaling record: compare with the N = 1 line's ms_per_step on the same total mesh"}
            s2.close()

    if rank == 0:
        par = "single GPU"
        if world > 1:
            par = (f"rcb-slab x{world}, one-layer halo, " +
                   ("library-owned transport: peer stores over NVLink + epoch flags, consumed by the band tiles of the single tile-kernel launch"
                    if transport == "ipc" else "pack -> NCCL send/recv -> kernel per step" + (" overlapped with the tiles without halo faces" if args.overlap else "")))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(N / 1e6),
                           "cells_per_gpu": N, "faces_per_gpu": F, "tile_cells": args.tile, "n_tiles": st["n_tiles"],
                           "l2": "inputs (state + mesh tables >> 126 MB L2) larger than L2, no flush needed",
                           "parallelism": par, "transport": transport, "hg_create_s": create_s},
                "roofline": roofline, "roofline_vjp": roofline_vjp,
                "rhs": {"value": N_total / (ms_rhs_step * 1e-3), "unit": UNIT, "ms": ms_rhs_step},
                "vjp": {"value": N_total / (ms_vjp_step * 1e-3), "unit": UNIT, "ms": ms_vjp_step},
                "sustained": sustained, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if jvp is not None:
            line["jvp"] = jvp
        line.update(side)
        if strong is not None:
            line["strong"] = strong
        if uncoupled is not None:
            line["uncoupled"] = uncoupled
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
