"""ORACLE (test infrastructure, NOT product code) -- ctypes front end of oracle/swe_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The reference (Julia) cannot run in this image; see the header of swe_oracle.cpp for what pins it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "liboracle.so")

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class MeshDesc(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_faces", C.c_int64), ("n_ghost", C.c_int64), ("ld", C.c_int64),
                ("index_base", C.c_int32), ("cell_nfaces", c_i64p), ("cell_faces", c_i64p),
                ("cell_neighbors", c_i64p), ("cell_normals", c_f64p), ("face_is_boundary", c_u8p),
                ("face_lengths", c_f64p), ("cell_areas", c_f64p), ("cell_centroids", c_f64p)]


class BcDesc(C.Structure):
    _fields_ = [("n_inletq", C.c_int64), ("n_exith", C.c_int64), ("n_wall", C.c_int64), ("n_symm", C.c_int64),
                ("bc_ptr", c_i64p), ("ghost_ids", c_i64p), ("internal_cells", c_i64p),
                ("outward_normals", c_f64p), ("face_lengths", c_f64p),
                ("n_halo", C.c_int64), ("halo_flip", c_u8p), ("halo_area", c_f64p)]


class FieldsDesc(C.Structure):
    _fields_ = [("g", C.c_double), ("k_n", C.c_double), ("h_small", C.c_double), ("riemann_solver", C.c_char_p),
                ("hstill", c_f64p), ("hstill_ghost", c_f64p), ("zb_cells", c_f64p), ("zb_ghost", c_f64p),
                ("S0_cells", c_f64p), ("ManningN_cells", c_f64p), ("matID_cells", c_i64p), ("n_mat", C.c_int64),
                ("inletQ_TotalQ", c_f64p), ("exitH_WSE", c_f64p)]


def build(force=False):
    """g++ -O2, no FMA contraction, OpenMP for the 'generous' CPU baseline."""
    src = os.path.join(HERE, "swe_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "hydrograd_b200.h")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp", "-o", LIB, src]
    subprocess.run(cmd, check=True)
    return LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Oracle:
    """Holds the flat arrays (dict from oracle.srh2d_ref.flatten or any producer of the same keys)."""

    def __init__(self, flat: dict):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        self.flat = f = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()}
        self.N, self.F, self.B = int(f["n_cells"]), int(f["n_faces"]), int(f["n_ghost"])
        cc = f.get("cell_centroids")
        self.mesh = MeshDesc(f["n_cells"], f["n_faces"], f["n_ghost"], f["ld"], f["index_base"],
                             _p(f["cell_nfaces"], c_i64p), _p(f["cell_faces"], c_i64p),
                             _p(f["cell_neighbors"], c_i64p), _p(f["cell_normals"], c_f64p),
                             _p(f["face_is_boundary"], c_u8p), _p(f["face_lengths"], c_f64p),
                             _p(f["cell_areas"], c_f64p), _p(cc, c_f64p))
        self.bc = BcDesc(f["n_inletq"], f["n_exith"], f["n_wall"], f["n_symm"], _p(f["bc_ptr"], c_i64p),
                         _p(f["bc_ghost_ids"], c_i64p), _p(f["bc_internal_cells"], c_i64p),
                         _p(f["bc_normals"], c_f64p), _p(f["bc_lengths"], c_f64p), 0, None, None)
        self.fields = FieldsDesc(f["g"], f["k_n"], f["h_small"], b"Roe", _p(f["hstill"], c_f64p),
                                 _p(f["hstill_ghost"], c_f64p), _p(f["zb_cells"], c_f64p), _p(f["zb_ghost"], c_f64p),
                                 _p(f["S0_cells"], c_f64p), _p(f["ManningN_cells"], c_f64p),
                                 _p(f.get("matID_cells"), c_i64p), f.get("n_mat", 0),
                                 _p(f["inletQ_TotalQ"], c_f64p), _p(f["exitH_WSE"], c_f64p))
        self.lib.oracle_max_threads.restype = C.c_int

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def _args(self):
        return C.byref(self.mesh), C.byref(self.bc), C.byref(self.fields)

    @staticmethod
    def _params(params):
        p = np.ascontiguousarray(params if params is not None else np.zeros(1), dtype=np.float64)
        return p, (0 if params is None else p.size)

    def rhs(self, Q, params=None, active=0, nthreads=1, ghosts=False):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        p, npar = self._params(params)
        out = np.empty(3 * self.N)
        gh = np.empty(4 * self.B) if ghosts else None
        nthreads = nthreads or self.max_threads()      # 0 = every host core ("ref-generous")
        rc = self.lib.oracle_rhs(*self._args(), _p(Q, c_f64p), _p(p, c_f64p), C.c_int64(npar), C.c_int(active),
                                 _p(out, c_f64p), C.c_int(nthreads), _p(gh, c_f64p))
        if rc:
            raise RuntimeError(f"oracle_rhs failed with code {rc}")
        return (out, gh.reshape(4, self.B)) if ghosts else out

    def jvp(self, Q, vQ, params=None, vP=None, active=0, nthreads=1):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        vQ = np.ascontiguousarray(vQ if vQ is not None else np.zeros_like(Q), dtype=np.float64)
        p, npar = self._params(params)
        vp = np.ascontiguousarray(vP if vP is not None else np.zeros_like(p), dtype=np.float64)
        out, jv = np.empty(3 * self.N), np.empty(3 * self.N)
        rc = self.lib.oracle_rhs_jvp(*self._args(), _p(Q, c_f64p), _p(vQ, c_f64p), _p(p, c_f64p), _p(vp, c_f64p),
                                     C.c_int64(npar), C.c_int(active), _p(out, c_f64p), _p(jv, c_f64p),
                                     C.c_int(nthreads or self.max_threads()))
        if rc:
            raise RuntimeError(f"oracle_rhs_jvp failed with code {rc}")
        return out, jv

    def vjp_bruteforce(self, Q, lam, params=None, active=0, nthreads=0):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        p, npar = self._params(params)
        Qbar, pbar = np.zeros(3 * self.N), np.zeros(max(npar, 1))
        nthreads = nthreads or self.max_threads()
        rc = self.lib.oracle_rhs_vjp_bruteforce(*self._args(), _p(Q, c_f64p), _p(p, c_f64p), C.c_int64(npar),
                                                C.c_int(active), _p(lam, c_f64p), _p(Qbar, c_f64p), _p(pbar, c_f64p),
                                                C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"oracle_rhs_vjp_bruteforce failed with code {rc}")
        return Qbar, pbar[:npar]

    MANNING_TYPES = {"constant": 0, "power_law": 1, "sigmoid": 2, "inverse": 3, "h_Umag_ks": 4}

    def set_manning_function(self, kind="constant", n_lower=0.0, n_upper=0.0, k=0.0, h_mid=0.0, ks_cells=None):
        """Variable Manning's n of forward simulations (semi_discretize_swe_2D.jl:140-149) for the following calls."""
        p = np.array([n_lower, n_upper, k, h_mid], dtype=np.float64)
        ks = None if ks_cells is None else np.ascontiguousarray(ks_cells, dtype=np.float64)
        rc = self.lib.oracle_set_manning_function(C.c_int(self.MANNING_TYPES[kind]), _p(p, c_f64p), _p(ks, c_f64p), C.c_int64(self.N))
        if rc:
            raise RuntimeError(f"oracle_set_manning_function failed with code {rc}")

    def manning_closure(self, kind, h, Umag=None, ks=None, n_lower=0.0, n_upper=0.0, k=0.0, h_mid=0.0):
        h = np.ascontiguousarray(h, dtype=np.float64)
        U = None if Umag is None else np.ascontiguousarray(Umag, dtype=np.float64)
        K = None if ks is None else np.ascontiguousarray(ks, dtype=np.float64)
        p = np.array([n_lower, n_upper, k, h_mid], dtype=np.float64)
        out = [np.zeros(h.size) for _ in range(4)]
        self.lib.oracle_manning_closure(C.c_int(self.MANNING_TYPES[kind]), _p(p, c_f64p), C.c_int64(h.size), _p(h, c_f64p),
                                        _p(U, c_f64p), _p(K, c_f64p), *[_p(o, c_f64p) for o in out])
        return dict(zip(("n", "h_ks", "f", "Re"), out))

    def euler(self, Q, dt, nsteps, params=None, active=0, nthreads=1):
        Q = np.array(Q, dtype=np.float64, copy=True)
        p, npar = self._params(params)
        nthreads = nthreads or self.max_threads()
        rc = self.lib.oracle_euler(*self._args(), _p(Q, c_f64p), _p(p, c_f64p), C.c_int64(npar), C.c_int(active),
                                   C.c_double(dt), C.c_int64(nsteps), C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"oracle_euler failed with code {rc}")
        return Q

    def bed(self, zb):
        zb = np.ascontiguousarray(zb, dtype=np.float64)
        zbg, S0 = np.empty(self.B), np.empty(2 * self.N)
        self.lib.oracle_bed(*self._args(), _p(zb, c_f64p), _p(zbg, c_f64p), _p(S0, c_f64p))
        return zbg, S0

    def roe(self, s12, nx, ny, g=9.81, hmin=1e-3):
        s = np.ascontiguousarray(s12, dtype=np.float64)
        out = np.empty(3)
        self.lib.oracle_roe(_p(s, c_f64p), C.c_double(g), C.c_double(nx), C.c_double(ny), C.c_double(hmin),
                            _p(out, c_f64p))
        return out

    def friction(self, h, qx, qy, mann, g=9.81, kn=1.0, hs=1e-3):
        h, qx, qy, mann = (np.ascontiguousarray(a, dtype=np.float64) for a in (h, qx, qy, mann))
        fx, fy = np.empty_like(h), np.empty_like(h)
        self.lib.oracle_friction(C.c_int64(h.size), _p(h, c_f64p), _p(qx, c_f64p), _p(qy, c_f64p), _p(mann, c_f64p),
                                 C.c_double(g), C.c_double(kn), C.c_double(hs), _p(fx, c_f64p), _p(fy, c_f64p))
        return fx, fy
