"""Host-side mirror of the reference's interface for the RHS path, on top of the C ABI.

Julia is not available in this image, so this Python layer plays the role of the Julia shim
(hydrograd.jl_b200/julia/HydrogradB200.jl): same names, argument meaning and error behaviour as

  swe_2d_rhs(dQdt, Q, params_vector, t, p_extra)     src/fvm/discretization/semi_discretize_swe_2D.jl:18-19
  custom_ODE_solve(ode_f, Q0, params_vector, extra)  src/ode_solvers/custom_ODE_solvers.jl:36
  SWE2D_Extra_Parameters                             src/applications/application_commons.jl:7-44

All arithmetic happens in the CUDA library; numpy is only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L


class HydrogradError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{L.ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=L.c_f64p):
    return a.ctypes.data_as(t) if a is not None else None


def _descs(f: dict):
    """ctypes descriptors over (contiguous, correctly typed copies of) the flat tables."""
    keep = {}
    for k in ("cell_nfaces", "cell_faces", "cell_neighbors", "bc_ptr", "bc_ghost_ids", "bc_internal_cells"):
        keep[k] = np.ascontiguousarray(f[k], dtype=np.int64)
    keep["matID_cells"] = (np.ascontiguousarray(f["matID_cells"], dtype=np.int64)
                           if f.get("matID_cells") is not None else None)
    keep["face_is_boundary"] = np.ascontiguousarray(f["face_is_boundary"], dtype=np.uint8)
    for k in ("cell_normals", "face_lengths", "cell_areas", "bc_normals", "bc_lengths", "hstill", "hstill_ghost",
              "zb_cells", "zb_ghost", "S0_cells", "ManningN_cells", "inletQ_TotalQ", "exitH_WSE"):
        keep[k] = _f64(f[k])
    keep["cell_centroids"] = _f64(f["cell_centroids"]) if f.get("cell_centroids") is not None else None
    mesh = L.MeshDesc(int(f["n_cells"]), int(f["n_faces"]), int(f["n_ghost"]), int(f["ld"]), int(f["index_base"]),
                      _p(keep["cell_nfaces"], L.c_i64p), _p(keep["cell_faces"], L.c_i64p),
                      _p(keep["cell_neighbors"], L.c_i64p), _p(keep["cell_normals"]),
                      _p(keep["face_is_boundary"], L.c_u8p), _p(keep["face_lengths"]), _p(keep["cell_areas"]),
                      _p(keep["cell_centroids"]))
    n_halo = int(f.get("n_halo", 0))
    keep["halo_flip"] = np.ascontiguousarray(f["halo_flip"], dtype=np.uint8) if n_halo else None
    keep["halo_area"] = _f64(f["halo_area"]) if n_halo else None
    bc = L.BcDesc(int(f["n_inletq"]), int(f["n_exith"]), int(f["n_wall"]), int(f["n_symm"]),
                  _p(keep["bc_ptr"], L.c_i64p), _p(keep["bc_ghost_ids"], L.c_i64p),
                  _p(keep["bc_internal_cells"], L.c_i64p), _p(keep["bc_normals"]), _p(keep["bc_lengths"]),
                  n_halo, _p(keep["halo_flip"], L.c_u8p), _p(keep["halo_area"]))
    keep["solver"] = f.get("riemann_solver", "Roe").encode()
    fields = L.FieldsDesc(float(f["g"]), float(f["k_n"]), float(f["h_small"]), keep["solver"],
                          _p(keep["hstill"]), _p(keep["hstill_ghost"]), _p(keep["zb_cells"]), _p(keep["zb_ghost"]),
                          _p(keep["S0_cells"]), _p(keep["ManningN_cells"]),
                          _p(keep["matID_cells"], L.c_i64p), int(f.get("n_mat", 0)),
                          _p(keep["inletQ_TotalQ"]), _p(keep["exitH_WSE"]))
    return mesh, bc, fields, keep


def _options(lib, device=0, tile_cells=256, reorder=True, strict=False, path=0, threads=0, vjp_variant=0, prefetch=0, face_blocks=True,
             ude_generic=False, pipeline_chunks=0):
    opt = L.Options()
    lib.hg_default_options(C.byref(opt))
    opt.device, opt.tile_cells, opt.reorder, opt.strict, opt.path = device, tile_cells, int(reorder), int(strict), path
    opt.reserved[0] = threads
    opt.reserved[1] = int(pipeline_chunks)
    opt.reserved[2] = int(vjp_variant)
    opt.reserved[3] = int(prefetch)
    opt.reserved[4] = 0 if face_blocks else 1
    opt.reserved[5] = 1 if ude_generic else 0
    return opt


def plan_stats(flat: dict, tile_cells=256, reorder=True, want_perm=False):
    """Host-only preview of the hg_create preprocessing (no GPU needed): tiling statistics (+ permutation)."""
    lib = L.load()
    mesh, bc, fields, keep = _descs(flat)
    opt = _options(lib, 0, tile_cells, reorder)
    stats = np.zeros(8, dtype=np.int64)
    perm = np.zeros(int(flat["n_cells"]), dtype=np.int64) if want_perm else None
    rc = lib.hg_plan_stats(C.byref(mesh), C.byref(bc), C.byref(fields), C.byref(opt), _p(stats, L.c_i64p),
                           _p(perm, L.c_i64p))
    if rc:
        raise HydrogradError(rc, (lib.hg_last_error(None) or b"").decode())
    out = dict(zip(("n_tiles", "max_local", "max_faces", "smem_bytes", "halo_cells", "tile_faces", "interior_tile_faces",
                    "sum_cell_faces"), (int(x) for x in stats)))
    return (out, perm) if want_perm else out


def plan_pipeline(flat: dict, tile_cells=256, reorder=True, pipeline_chunks=0):
    """Host-only: stage tables of the host-buffer pipeline (hg_plan_pipeline): dict(n_chunks, chunk_cells, n_tiles, tile_cells,
    margin, tile_stage [n_tiles], chunk_done [n_chunks])."""
    lib = L.load()
    mesh, bc, fields, keep = _descs(flat)
    opt = _options(lib, 0, tile_cells, reorder, pipeline_chunks=pipeline_chunks)
    hdr = np.zeros(5, dtype=np.int64)
    i32p = C.POINTER(C.c_int32)
    rc = lib.hg_plan_pipeline(C.byref(mesh), C.byref(bc), C.byref(fields), C.byref(opt), _p(hdr, L.c_i64p), None, None)
    if rc:
        raise HydrogradError(rc, (lib.hg_last_error(None) or b"").decode())
    ts, cd = np.zeros(int(hdr[2]), dtype=np.int32), np.zeros(int(hdr[0]), dtype=np.int32)
    rc = lib.hg_plan_pipeline(C.byref(mesh), C.byref(bc), C.byref(fields), C.byref(opt), _p(hdr, L.c_i64p), _p(ts, i32p), _p(cd, i32p))
    if rc:
        raise HydrogradError(rc, (lib.hg_last_error(None) or b"").decode())
    return dict(n_chunks=int(hdr[0]), chunk_cells=int(hdr[1]), n_tiles=int(hdr[2]), tile_cells=int(hdr[3]), margin=int(hdr[4]),
                tile_stage=ts, chunk_done=cd)


def plan_tables(flat: dict, tile_cells=256, reorder=True):
    """Host-only: the tile tables hg_create would upload (hg_plan_open / hg_plan_array), copied into numpy arrays by name,
    plus the entries of "dims" as ints (N, B, n_tiles, T, NF, Ns, n_desc, n_chunks, n_interior_tiles, comm_band0)."""
    lib = L.load()
    mesh, bc, fields, keep = _descs(flat)
    opt = _options(lib, 0, tile_cells, reorder)
    h = C.c_void_p()
    rc = lib.hg_plan_open(C.byref(mesh), C.byref(bc), C.byref(fields), C.byref(opt), C.byref(h))
    if rc:
        raise HydrogradError(rc, (lib.hg_last_error(None) or b"").decode())
    types = {0: np.float64, 1: np.int64, 2: np.uint8, 3: np.int32, 4: np.uint32, 5: np.uint16}
    out = {}
    try:
        for name in ("dims", "perm", "iperm", "tile_desc", "halo", "bface_e", "tile_order", "band_order", "comm_order", "face_lr", "cf_idx", "face_nx", "face_ny", "face_len",
                     "bc_type", "bc_group", "bc_ghost", "bc_cell_ref", "inlet_ptr", "bcell_ref", "bcell_ptr", "bcell_ent", "cf_rev", "bc_nx", "bc_ny", "bc_l53", "bc_l23", "bc_hstill", "bc_zb"):
            ptr, cnt, dt = C.c_void_p(), C.c_int64(), C.c_int32()
            rc = lib.hg_plan_array(h, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(dt))
            if rc:
                raise HydrogradError(rc, f"hg_plan_array({name})")
            ty = types[dt.value]
            n = cnt.value
            out[name] = (np.frombuffer((C.c_char * (n * np.dtype(ty).itemsize)).from_address(ptr.value), dtype=ty).copy()
                         if n else np.zeros(0, dtype=ty))
    finally:
        lib.hg_plan_close(h)
    out.update(zip(("N", "B", "n_tiles", "T", "NF", "Ns", "n_desc", "n_chunks", "n_interior_tiles", "comm_band0"), (int(x) for x in out["dims"])))
    return out


class Context:
    """Owns one hg_ctx (one mesh on one GPU).  `flat` holds the flat tables of include/hydrograd_b200.h
    (see INTEGRATION.md for how the Julia structs map onto them)."""

    def __init__(self, flat: dict, device=0, tile_cells=256, reorder=True, strict=False, path=0, threads=0, vjp_variant=0, prefetch=0,
                 face_blocks=True, ude_generic=False, pipeline_chunks=0):
        self.lib = L.load()
        self._h = C.c_void_p()
        mesh, bc, fields, keep = _descs(flat)
        opt = _options(self.lib, device, tile_cells, reorder, strict, path, threads, vjp_variant, prefetch, face_blocks, ude_generic, pipeline_chunks)
        rc = self.lib.hg_create(C.byref(self._h), C.byref(mesh), C.byref(bc), C.byref(fields), C.byref(opt))
        del keep            # the library copied what it needs (ownership rule of the ABI)
        if rc:
            raise HydrogradError(rc, (self.lib.hg_last_error(None) or b"").decode())
        self.N = int(flat["n_cells"])
        self.n_ghost = int(flat["n_ghost"])

    # ---------------------------------------------------------------- plumbing
    def _ck(self, rc):
        if rc:
            raise HydrogradError(rc, (self.lib.hg_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.hg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _params(params, active):
        a = L.ACTIVE_PARAM[active] if not isinstance(active, int) else active
        if a == 0 or params is None:
            return None, 0, a
        p = _f64(params)
        return p, p.size, a

    # ---------------------------------------------------------------- host-buffer API (drop-in semantics)
    def rhs(self, Q, params=None, active=None, t=0.0, out=None):
        Q = _f64(Q)
        if Q.size != 3 * self.N:
            raise HydrogradError(1, f"Q has length {Q.size}, expected {3 * self.N}")
        out = np.empty(3 * self.N) if out is None else out
        p, n, a = self._params(params, active)
        self._ck(self.lib.hg_rhs(self._h, _p(Q), _p(p), n, a, float(t), _p(out)))
        return out

    def rhs_vjp(self, Q, lam, params=None, active=None, t=0.0, want_ncell_bar=False):
        """hg_rhs_vjp.  Q = None: at the resident state (the one the last rhs / set_state call uploaded; see
        state_generation) -- only the cotangent crosses PCIe."""
        Q, lam = (_f64(Q) if Q is not None else None), _f64(lam)
        if lam.size != 3 * self.N or (Q is not None and Q.size != 3 * self.N):
            raise HydrogradError(1, f"Q / lambda must have length {3 * self.N}")
        p, n, a = self._params(params, active)
        Qbar = np.empty(3 * self.N)
        pbar = np.zeros(max(n, 1))
        nbar = np.empty(self.N) if want_ncell_bar else None
        self._ck(self.lib.hg_rhs_vjp(self._h, _p(Q), _p(p), n, a, float(t), _p(lam), _p(Qbar), _p(pbar), _p(nbar)))
        return (Qbar, pbar[:n], nbar) if want_ncell_bar else (Qbar, pbar[:n])

    def rhs_jvp(self, Q, v, params=None, active=None, pdot=None, t=0.0, want_rhs=True):
        """Forward mode on a strict context: (dQdt, J_Q v + J_p pdot) -- one partial of a ForwardDiff.Dual pass (hg_rhs_jvp)."""
        p, n, a = self._params(params, active)
        Q, v = _f64(Q), _f64(v)
        if Q.size != 3 * self.N or v.size != 3 * self.N:
            raise HydrogradError(1, f"Q / v have lengths {Q.size} / {v.size}, expected {3 * self.N}")
        pd = _f64(pdot) if pdot is not None else None
        if pd is not None and pd.size != n:
            raise ValueError(f"pdot has length {pd.size}, expected {n}")
        out = np.empty(3 * self.N) if want_rhs else None
        jv = np.empty(3 * self.N)
        self._ck(self.lib.hg_rhs_jvp(self._h, _p(Q), _p(p), n, a, float(t), _p(v), _p(pd), _p(out), _p(jv)))
        return (out, jv) if want_rhs else jv

    def rhs_jvp_multi(self, Q, V, params=None, active=None, Pdot=None, t=0.0):
        """K directions in one call (a ForwardDiff chunk): V [K, 3N], Pdot [K, n_params] or None -> (dQdt, JV [K, 3N])."""
        p, n, a = self._params(params, active)
        Q, V = _f64(Q), _f64(V)
        if V.ndim != 2 or Q.size != 3 * self.N or V.shape[1] != 3 * self.N:
            raise HydrogradError(1, f"Q / V have shapes {Q.shape} / {V.shape}, expected ({3 * self.N},) / (K, {3 * self.N})")
        K = V.shape[0]
        Pd = _f64(Pdot) if Pdot is not None and n > 0 else None
        if Pd is not None and Pd.shape != (K, n):
            raise ValueError(f"Pdot has shape {Pd.shape}, expected {(K, n)}")
        out, JV = np.empty(3 * self.N), np.empty((K, 3 * self.N))
        self._ck(self.lib.hg_rhs_jvp_multi(self._h, _p(Q), _p(p), n, a, float(t), K, _p(V), _p(Pd), _p(out), _p(JV)))
        return out, JV

    def solve_tsit5_sens(self, Q0, params, active, t0, t1, dt, adaptive=True, abstol=1e-6, reltol=1e-3, t_save=()):
        """The reference's sensitivity driver (ForwardDiff.jacobian around the Tsit5 solve, swe_2D_sensitivity.jl:34-80) on a
        strict context: returns (Q(t1) [3N], S [n_params, 3N] with S[k] = dQ(t1)/dp_k, stats); with t_save also the values at
        the save times [len(t_save), 3N] (dense output), stored in stats["saves"]."""
        p, n, a = self._params(params, active)
        Q0 = _f64(Q0)
        if Q0.size != 3 * self.N:
            raise HydrogradError(1, f"Q0 has length {Q0.size}, expected {3 * self.N}")
        QT, S = np.empty(3 * self.N), np.empty((max(n, 1), 3 * self.N))
        stats = np.zeros(3, dtype=np.int64)
        ts = _f64(np.asarray(t_save, dtype=np.float64)) if len(t_save) else None
        saves = np.empty((len(t_save), 3 * self.N)) if len(t_save) else None
        self._ck(self.lib.hg_solve_tsit5_sens(self._h, _p(Q0), _p(p), n, a, float(t0), float(t1), float(dt), int(bool(adaptive)),
                                              float(abstol), float(reltol), _p(ts), len(t_save), _p(saves), _p(QT), _p(S),
                                              _p(stats, L.c_i64p)))
        return QT, S[:n], dict(accepted=int(stats[0]), rejected=int(stats[1]), rhs=int(stats[2]), saves=saves)

    def rhs_vjp_into(self, Q, lam, Qbar_out, params=None, active=None, pbar_out=None):
        """hg_rhs_vjp into caller-owned (e.g. pinned) buffers; pbar_out [n_params] when a parameter is active."""
        p, n, a = self._params(params, active)
        Q = _f64(Q) if Q is not None else None      # None: the resident state (hg_rhs_vjp with Q = NULL)
        self._ck(self.lib.hg_rhs_vjp(self._h, _p(Q), _p(p), n, a, 0.0, _p(_f64(lam)), _p(Qbar_out), _p(pbar_out) if n else None, None))

    def state_generation(self):
        """Changes whenever the resident state changes (hg_state_generation)."""
        return int(self.lib.hg_state_generation(self._h))

    def get_vjp_into(self, Qbar_out):
        self._ck(self.lib.hg_get_vjp(self._h, _p(Qbar_out), None, None))

    # ---------------------------------------------------------------- device-resident API
    def set_state(self, Q):
        self._ck(self.lib.hg_set_state(self._h, _p(_f64(Q))))

    def get_state(self):
        out = np.empty(3 * self.N)
        self._ck(self.lib.hg_get_state(self._h, _p(out)))
        return out

    def set_params(self, params, active):
        p, n, a = self._params(params, active)
        self._ck(self.lib.hg_set_params(self._h, _p(p), n, a))

    def set_fields(self, ManningN_cells=None, zb_cells=None, zb_ghost=None, S0_cells=None, inletQ_TotalQ=None,
                   exitH_WSE=None):
        arrs = [None if a is None else _f64(a)
                for a in (ManningN_cells, zb_cells, zb_ghost, S0_cells, inletQ_TotalQ, exitH_WSE)]
        self._ck(self.lib.hg_set_fields(self._h, *[_p(a) for a in arrs]))

    MANNING_TYPES = {"constant": 0, "power_law": 1, "sigmoid": 2, "inverse": 3, "h_Umag_ks": 4}

    def set_manning_function(self, kind="constant", n_lower=0.0, n_upper=0.0, k=0.0, h_mid=0.0, ks_cells=None):
        """forward_settings.ManningN_option = "variable" (semi_discretize_swe_2D.jl:140-149): n(h) / n(h, |U|, ks) evaluated
        inside every RHS; `kind` and the parameters are ManningN_function_type / ManningN_function_parameters of the control
        file (create_manning_function, process_ManningN_2D.jl:119-133); ks_cells[N] for "h_Umag_ks"."""
        if kind not in self.MANNING_TYPES:
            raise ValueError(f"Unknown Manning's n function type: {kind}. Supported types: {', '.join(self.MANNING_TYPES)}.")
        p = np.array([n_lower, n_upper, k, h_mid], dtype=np.float64)
        ks = None if ks_cells is None else _f64(ks_cells)
        self._ck(self.lib.hg_set_manning_function(self._h, self.MANNING_TYPES[kind], _p(p), _p(ks)))

    def set_ude_model(self, model, ks_cells=None):
        """settings.bPerform_UDE with UDE_choice "ManningN_h" / "ManningN_h_Umag_ks" (semi_discretize_swe_2D.jl:165-178):
        `model` is a hydrograd.jl_b200.ude.UDEModel (None clears it); afterwards params_vector with active = "UDE" is the
        network's parameter vector, every RHS evaluates n = NN(h, |U|, ks) on the device and every VJP returns d/d theta."""
        if model is None:
            self._ck(self.lib.hg_set_ude_model(self._h, None, None))
            return
        ks = None if ks_cells is None else _f64(ks_cells)
        if ks is not None and ks.size != self.N:
            raise HydrogradError(1, f"ks_cells has length {ks.size}, expected {self.N}")
        self._ck(self.lib.hg_set_ude_model(self._h, C.byref(model.desc), _p(ks)))

    # ---------------------------------------------------------------- adjoint on the resident state
    def set_lambda(self, lam):
        self._ck(self.lib.hg_set_lambda(self._h, _p(_f64(lam))))

    def vjp_resident(self, phase=0):
        self._ck(self.lib.hg_vjp_resident_phase(self._h, phase) if phase else self.lib.hg_vjp_resident(self._h))

    def get_vjp(self, n_params=0, want_ncell_bar=False):
        Qbar = np.empty(3 * self.N)
        pbar = np.zeros(max(n_params, 1))
        nbar = np.empty(self.N) if want_ncell_bar else None
        self._ck(self.lib.hg_get_vjp(self._h, _p(Qbar), _p(pbar), _p(nbar)))
        return (Qbar, pbar[:n_params], nbar) if want_ncell_bar else (Qbar, pbar[:n_params])

    # ---------------------------------------------------------------- multi-GPU halo plumbing
    def halo_info(self):
        nn, ne = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.hg_halo_info(self._h, C.byref(nn), C.byref(ne)))
        counts = np.zeros(max(nn.value, 1), dtype=np.int64)
        self._ck(self.lib.hg_halo_counts(self._h, _p(counts, L.c_i64p)))
        return counts[:nn.value]

    def halo_buffers(self):
        """(send_ptr, recv_ptr, n_doubles): device addresses of the context-owned halo buffers."""
        s, r, n = L.c_f64p(), L.c_f64p(), C.c_int64(0)
        self._ck(self.lib.hg_halo_buffers(self._h, C.byref(s), C.byref(r), C.byref(n)))
        return C.cast(s, C.c_void_p).value, C.cast(r, C.c_void_p).value, n.value

    # library-owned exchange over NVLink peer memory (hg_comm.cu); parallel.connect_contexts / connect_ranks wire it up
    COMM_HANDLE_BYTES = 128

    def comm_export(self):
        buf = C.create_string_buffer(self.COMM_HANDLE_BYTES)
        self._ck(self.lib.hg_comm_export(self._h, buf))
        return buf.raw

    def comm_connect(self, handles, entry_offsets, flag_indices):
        blob = b"".join(handles)
        off = np.ascontiguousarray(entry_offsets, dtype=np.int64)
        idx = np.ascontiguousarray(flag_indices, dtype=np.int64)
        self._ck(self.lib.hg_comm_connect(self._h, len(handles), blob, _p(off, L.c_i64p), _p(idx, L.c_i64p)))

    def comm_init_shm(self, job_name, rank, world, neighbor_ranks):
        nb = (C.c_int32 * len(neighbor_ranks))(*[int(r) for r in neighbor_ranks])
        self._ck(self.lib.hg_comm_init_shm(self._h, job_name.encode(), int(rank), int(world), nb))

    def comm_set_auto(self, on=True):
        self._ck(self.lib.hg_comm_set_auto(self._h, int(on)))

    def comm_exchange(self, with_lambda=False):
        self._ck(self.lib.hg_comm_exchange(self._h, int(with_lambda)))

    def comm_set_allreduce(self, fn):
        """fn(numpy view of the doubles to sum over ranks, in place).  Needed by adaptive solves on multi-rank contexts."""
        def cb(ptr, n, _user):
            try:
                fn(np.ctypeslib.as_array(ptr, shape=(n,)))
                return 0
            except Exception:  # noqa: BLE001 -- never raise through the C boundary
                return 1
        self._allreduce_cb = L.ALLREDUCE_FN(cb)          # keep the trampoline alive as long as the context
        self._ck(self.lib.hg_comm_set_allreduce(self._h, self._allreduce_cb, None))

    def comm_disconnect(self):
        self._ck(self.lib.hg_comm_disconnect(self._h))

    def halo_pack(self, with_lambda=False):
        self._ck(self.lib.hg_halo_pack(self._h, int(with_lambda)))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.hg_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def rhs_resident(self, phase=0):
        """phase 1 / 2: tiles without / with halo faces (multi-GPU overlap, hg_rhs_resident_phase)."""
        self._ck(self.lib.hg_rhs_resident_phase(self._h, phase) if phase else self.lib.hg_rhs_resident(self._h))

    def get_rhs(self, out=None):
        out = np.empty(3 * self.N) if out is None else out
        self._ck(self.lib.hg_get_rhs(self._h, _p(out)))
        return out

    def sync(self):
        self._ck(self.lib.hg_sync(self._h))

    def step_euler(self, dt, nsteps=1):
        self._ck(self.lib.hg_step_euler(self._h, float(dt), int(nsteps)))

    def step_rk4(self, dt, nsteps=1):
        self._ck(self.lib.hg_step_rk4(self._h, float(dt), int(nsteps)))

    def step_ode_euler(self, dt, nsteps=1):
        """solve(prob, Euler(), dt=dt) steps (swe_2D_forward_simulation.jl:41): no dry mask, unlike step_euler."""
        self._ck(self.lib.hg_step_ode_euler(self._h, float(dt), int(nsteps)))

    def step_ab3(self, dt, nsteps=1, restart=False):
        """solve(prob, AB3(), dt=dt) steps (swe_2D_forward_simulation.jl:47); the multistep history persists between calls."""
        self._ck(self.lib.hg_step_ab3(self._h, float(dt), int(nsteps), int(bool(restart))))

    def set_controller_pow(self, mode):
        """"exact" | "fastpow": how the PI controller of solve_tsit5 raises EEst and qold to their powers (hg_set_controller_pow)."""
        self._ck(self.lib.hg_set_controller_pow(self._h, {"exact": 0, "fastpow": 1}[mode]))

    def solve_tsit5(self, t0, t1, dt, adaptive=True, abstol=1e-6, reltol=1e-3, t_save=(), saveat="stop"):
        """solve(prob, Tsit5(), adaptive=adaptive, dt=dt, saveat=t_save; abstol, reltol) on the resident state
        (swe_2D_forward_simulation.jl:38-41); returns (saved states [len(t_save), 3N], stats dict).
        saveat="interp": OrdinaryDiffEq's own saveat (dense output, steps independent of t_save); "stop": save times are stops."""
        if saveat not in ("stop", "interp"):
            raise ValueError(f"saveat must be 'stop' or 'interp', got {saveat!r}")
        fn = self.lib.hg_solve_tsit5_dense if saveat == "interp" else self.lib.hg_solve_tsit5
        ts = _f64(np.asarray(t_save, dtype=np.float64)) if len(t_save) else None
        out = np.empty((len(t_save), 3 * self.N)) if len(t_save) else None
        stats = np.zeros(3, dtype=np.int64)
        self._ck(fn(self._h, float(t0), float(t1), float(dt), int(bool(adaptive)), float(abstol), float(reltol),
                    _p(ts), len(t_save), _p(out), _p(stats, L.c_i64p)))
        return out, dict(accepted=int(stats[0]), rejected=int(stats[1]), rhs=int(stats[2]))

    def euler_adjoint(self, Q0, lam_T, dt, nsteps, params=None, active=None):
        """Discrete adjoint of nsteps Euler steps: returns (Q_T, Q0bar, pbar)."""
        p, n, a = self._params(params, active)
        QT, Q0bar, pbar = np.empty(3 * self.N), np.empty(3 * self.N), np.zeros(max(n, 1))
        self._ck(self.lib.hg_euler_adjoint(self._h, _p(_f64(Q0)), _p(p), n, a, float(dt), int(nsteps), _p(_f64(lam_T)),
                                           _p(QT), _p(Q0bar), _p(pbar)))
        return QT, Q0bar, pbar[:n]

    RK_METHODS = {"RK4": 0, "Tsit5": 1}

    def rk_adjoint(self, method, Q0, lam_T, dt, nsteps, params=None, active=None):
        """Discrete adjoint of nsteps fixed steps of `method` ("RK4" | "Tsit5"): returns (Q_T, Q0bar, pbar)."""
        p, n, a = self._params(params, active)
        QT, Q0bar, pbar = np.empty(3 * self.N), np.empty(3 * self.N), np.zeros(max(n, 1))
        self._ck(self.lib.hg_rk_adjoint(self._h, self.RK_METHODS[method], _p(_f64(Q0)), _p(p), n, a, float(dt), int(nsteps),
                                        _p(_f64(lam_T)), _p(QT), _p(Q0bar), _p(pbar)))
        return QT, Q0bar, pbar[:n]

    def rk_adjoint_steps(self, method, Q0, lam_T, steps, params=None, active=None):
        """Discrete adjoint over the given step sizes (e.g. last_steps() of an adaptive solve): returns (Q_T, Q0bar, pbar)."""
        p, n, a = self._params(params, active)
        hs = _f64(steps)
        QT, Q0bar, pbar = np.empty(3 * self.N), np.empty(3 * self.N), np.zeros(max(n, 1))
        self._ck(self.lib.hg_rk_adjoint_steps(self._h, self.RK_METHODS[method], _p(_f64(Q0)), _p(p), n, a, _p(hs), hs.size,
                                              _p(_f64(lam_T)), _p(QT), _p(Q0bar), _p(pbar)))
        return QT, Q0bar, pbar[:n]

    def last_steps(self):
        """Accepted step sizes of the last solve_tsit5."""
        n = C.c_int64(0)
        self._ck(self.lib.hg_last_steps(self._h, None, 0, C.byref(n)))
        h = np.empty(n.value)
        self._ck(self.lib.hg_last_steps(self._h, _p(h), n.value, C.byref(n)))
        return h

    def custom_ode_solve(self, Q0, params, active, t_start, t_end, dt):
        nsteps = int(np.floor((t_end - t_start) / dt + 1e-9)) + 1 if t_end >= t_start else 0
        sol = np.empty((max(nsteps, 1), 3 * self.N))   # row s = column s of the reference's 3N x nSaves `sol`
        ns = C.c_int64(0)
        p, n, a = self._params(params, active)
        self._ck(self.lib.hg_custom_ode_solve(self._h, _p(_f64(Q0)), _p(p), n, a, float(t_start), float(t_end),
                                              float(dt), _p(sol), nsteps, C.byref(ns)))
        return sol[:ns.value].T

    # ---------------------------------------------------------------- parameter ensembles
    def ensemble_alloc(self, n_members, per_member_manning=False):
        self._ck(self.lib.hg_ensemble_alloc(self._h, int(n_members), int(per_member_manning)))

    def ensemble_set_member(self, m, Q, params=None, active=None):
        p, n, a = self._params(params, active)
        self._ck(self.lib.hg_ensemble_set_member(self._h, int(m), _p(_f64(Q)), _p(p), n, a))

    def ensemble_step_euler(self, dt, nsteps=1):
        self._ck(self.lib.hg_ensemble_step_euler(self._h, float(dt), int(nsteps)))

    def ensemble_rhs(self):
        self._ck(self.lib.hg_ensemble_rhs(self._h))

    def ensemble_get_member(self, m, what="state"):
        out = np.empty(3 * self.N)
        self._ck(self.lib.hg_ensemble_get_member(self._h, int(m), 0 if what == "state" else 1, _p(out)))
        return out

    def time_ensemble(self, n_steps, dt):
        ms = C.c_float(0)
        self._ck(self.lib.hg_time_ensemble(self._h, int(n_steps), float(dt), C.byref(ms)))
        return ms.value

    # ---------------------------------------------------------------- measurement hooks
    def time_rhs(self, n_launches, fused_euler=False, dt=0.0):
        ms = C.c_float(0)
        self._ck(self.lib.hg_time_rhs(self._h, int(n_launches), int(fused_euler), float(dt), C.byref(ms)))
        return ms.value

    def time_jvp(self, n_directions, n_launches):
        ms = C.c_float(0)
        self._ck(self.lib.hg_time_jvp(self._h, int(n_directions), int(n_launches), C.byref(ms)))
        return ms.value

    def time_vjp(self, n_launches):
        ms = C.c_float(0)
        self._ck(self.lib.hg_time_vjp(self._h, int(n_launches), C.byref(ms)))
        return ms.value

    def debug_math(self, kind, x):
        """The kernels' branch-free fp64 helpers evaluated on the device (accuracy probe): kind 0 = 1/x, 1 = x^-1/2, 2 = sqrt,
        3 = sqrt(x^2 + eps), 4 = x^(-7/3)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self._ck(self.lib.hg_debug_math(self._h, int(kind), x.size, _p(x), _p(out)))
        return out

    def kernel_launches(self):
        return int(self.lib.hg_kernel_launches(self._h))

    def mesh_stats(self):
        v = [C.c_int64(0) for _ in range(5)]
        self._ck(self.lib.hg_mesh_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("n_cells", "n_faces", "sum_cell_faces", "n_tiles", "device_bytes"), (x.value for x in v)))

    def flush_l2(self):
        self._ck(self.lib.hg_flush_l2(self._h))


# ------------------------------------------------------------------------------------------------
# Reference-shaped front end
@dataclass
class swe_2D_consts:
    """src/constants/swe_2D_constants.jl:4-17 (fields the RHS reads + the Euler stepper's time data)."""
    g: float = 9.81
    k_n: float = 1.0
    h_small: float = 1.0e-3
    dt: float = 0.0
    tspan: tuple = (0.0, 0.0)
    RiemannSolver: str = "Roe"


@dataclass
class SWE2D_Extra_Parameters:
    """src/applications/application_commons.jl:7-44 restricted to what swe_2d_rhs reads.  `flat` carries the
    flattened mesh_2D / BoundaryConditions2D tables; the device context is created on first use."""
    flat: dict
    active_param_name: str = ""
    bInPlaceODE: bool = False
    swe_2D_constants: swe_2D_consts = field(default_factory=swe_2D_consts)
    options: dict = field(default_factory=dict)
    # "forward_simulation_options" of the reference's run_control.json (ManningN_option / function_type / function_parameters,
    # forward_simulation settings of process_control_file): a "variable" Manning's n selects the device closure, with
    # ks_cells gathered per material zone exactly like process_SRH_2D_input.jl:159-164 / process_ManningN_2D.jl:56-60
    forward_settings: dict = field(default_factory=dict)
    _ctx: Optional[Context] = None

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            f = dict(self.flat)
            c = self.swe_2D_constants
            f.update(g=c.g, k_n=c.k_n, h_small=c.h_small, riemann_solver=c.RiemannSolver)
            self._ctx = Context(f, **self.options)
            fs = {k.replace("forward_simulation_", ""): v for k, v in self.forward_settings.items()}
            if fs.get("ManningN_option", "constant") == "variable":
                kind = fs["ManningN_function_type"]
                prm = dict(fs.get("ManningN_function_parameters", {}))
                ks_cells = None
                if "ks" in prm:
                    base = 0
                    ks_cells = np.asarray(prm.pop("ks"), dtype=np.float64)[np.asarray(f["matID_cells"], dtype=np.int64) - base]
                self._ctx.set_manning_function(kind, ks_cells=ks_cells, **{k: float(v) for k, v in prm.items()})
        return self._ctx


def swe_2d_rhs(dQdt, Q, params_vector, t, p_extra: SWE2D_Extra_Parameters):
    """Same contract as the reference: in-place when p_extra.bInPlaceODE, else returns a new vector."""
    if p_extra.bInPlaceODE:
        p_extra.ctx.rhs(Q, params_vector, p_extra.active_param_name, t, out=dQdt)
        return dQdt
    return p_extra.ctx.rhs(Q, params_vector, p_extra.active_param_name, t)


def swe_2d_rhs_pullback(Q, params_vector, t, p_extra: SWE2D_Extra_Parameters):
    """`y, back = Zygote.pullback(Q -> swe_2d_rhs(Q, p, t, extra), Q)` (debug_AD.jl:60,75; what ZygoteVJP / the shim's
    `rrule` do per adjoint stage): returns (dQdt, back) with back(lam) -> (Qbar, pbar).  The primal's state stays on the
    device, so the pullback uploads only the cotangent as long as nothing has moved the state in between."""
    ctx = p_extra.ctx
    dQdt = ctx.rhs(Q, params_vector, p_extra.active_param_name, t)
    gen = ctx.state_generation()

    def back(lam):
        same = ctx.state_generation() == gen
        return ctx.rhs_vjp(None if same else Q, lam, params_vector, p_extra.active_param_name, t)

    return dQdt, back


def custom_ODE_solve(ode_f, Q0, params_vector, swe2d_extra_params: SWE2D_Extra_Parameters):
    """custom_ODE_solvers.jl:36-95.  `ode_f` is accepted for signature parity; the stepping runs on the device."""
    c = swe2d_extra_params.swe_2D_constants
    return swe2d_extra_params.ctx.custom_ode_solve(Q0, params_vector, swe2d_extra_params.active_param_name,
                                                   c.tspan[0], c.tspan[1], c.dt)
