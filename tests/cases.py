"""Shared loaders for the committed reference fixtures (tests/golden) and random states for fuzzing."""
import os

import numpy as np

from oracle import srh2d_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_STEMS = {
    "savannah": "savana_SI",
    "oneD_bump": "oneD_channel_with_bump_refined",
    "oneD_uniform": "oneD_channel_uniform_flow_refined",
    "simple": "simple",
    "oneD_bump_sens": "oneD_channel_with_bump_refined",
    "oneD_uniform_sens": "oneD_channel_uniform_flow_refined",
}
_IC = {
    "oneD_bump": ("constant", [0.33, 0.2, 0.0, 0.0]),          # run_control.json of the case
    "oneD_bump_sens": ("constant", [0.33, 0.2, 0.0, 0.0]),
    "oneD_uniform": ("constant", [3.857205, 3.0, 0.0, 0.0]),
    "oneD_uniform_sens": ("constant", [3.857205, 3.0, 0.0, 0.0]),
    "simple": ("constant", [1.0, 0.5, 0.0, 0.0]),
}
_cache = {}


def load(name):
    if name not in _cache:
        d = os.path.join(GOLD, name)
        ic = _IC.get(name)
        if name == "savannah":
            z = np.load(os.path.join(d, "ic.npz"))
            import json, tempfile
            with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
                json.dump({k: z[k].tolist() for k in z.files}, f)
            ic = ("from_file", f.name)
        _cache[name] = R.load_case(d, _STEMS[name] + ".srhhydro", ic)
    return _cache[name]


def truth(name):
    return np.load(os.path.join(GOLD, name, "truth.npz"))


def random_state(case, seed, dry_frac=0.05):
    """SURVEY 8(d) fuzz: h ~ LogUniform(1e-4, 10) (about 5 % below h_small), |u| ~ U(0,3), random direction."""
    rng = np.random.default_rng(seed)
    N = case.mesh.numOfCells
    h = np.exp(rng.uniform(np.log(1e-4), np.log(10.0), N))
    # force the requested share of (nearly) dry cells, including exact ties with h_small
    k = rng.random(N) < dry_frac
    h[k] = rng.choice([5e-4, 1e-3, 9.999e-4], size=int(k.sum()))
    sp = rng.uniform(0, 3, N)
    th = rng.uniform(0, 2 * np.pi, N)
    xi = h - case.hstill
    return np.concatenate([xi, h * sp * np.cos(th), h * sp * np.sin(th)])


def flux_scale(case, Q):
    """Denominator for 'relative' RHS errors: sum |flux*L|/A + |source| magnitude proxy per cell
    (BASELINE.md parity gates: relative to the un-cancelled magnitude, not to the residual)."""
    N = case.mesh.numOfCells
    h = np.maximum(Q[:N] + case.hstill, case.h_small)
    per = np.array([sum(case.mesh.face_lengths[f - 1] for f in case.mesh.cellFacesList[c, :case.mesh.cellNodesCount[c]])
                    for c in range(N)])
    u2 = (Q[N:2 * N] ** 2 + Q[2 * N:] ** 2) / h ** 2
    c = np.sqrt(case.g * h)
    s = (0.5 * case.g * h * h + h * u2 + h * np.sqrt(u2) * c + c * h) * per / case.mesh.cell_areas
    return np.concatenate([s, s, s]) + 1e-300


def flat_scale(flat, Q):
    """Same denominator as flux_scale, from the flat ABI tables (works for synthetic meshes too)."""
    N, ld = int(flat["n_cells"]), int(flat["ld"])
    base = int(flat["index_base"])
    faces = np.abs(np.asarray(flat["cell_faces"]).reshape(ld, N)) - base
    valid = np.arange(ld)[:, None] < np.asarray(flat["cell_nfaces"])[None, :]
    per = (np.asarray(flat["face_lengths"])[np.where(valid, faces, 0)] * valid).sum(0)
    hs, g = flat["h_small"], flat["g"]
    h = np.maximum(Q[:N] + flat["hstill"], hs)
    u2 = (Q[N:2 * N] ** 2 + Q[2 * N:] ** 2) / h ** 2
    c = np.sqrt(g * h)
    xi, hst = np.abs(Q[:N]), np.abs(np.asarray(flat["hstill"]))
    # the momentum flux carries the xi-form pressure 0.5 g (xi^2 + 2 xi hstill) (swe_2D_solvers.jl:122), whose
    # magnitude -- not 0.5 g h^2 -- is what cancels around a cell
    s = (0.5 * g * (h * h + xi * xi + 2 * xi * hst) + h * u2 + h * np.sqrt(u2) * c + c * (h + xi)) * per / np.asarray(flat["cell_areas"])
    # a cell also feels its neighbours' pressure: use the max over the cell and its face neighbours
    nb = np.asarray(flat["cell_neighbors"]).reshape(ld, N) - base
    isb = np.asarray(flat["face_is_boundary"])[np.where(valid, faces, 0)].astype(bool)
    nbs = np.where(valid & ~isb, s[np.where(valid & ~isb, nb, 0)], 0.0).max(0)
    s = np.maximum(s, nbs)
    # source magnitudes: Manning friction and bed-slope terms (semi_discretize_swe_2D.jl:463-478, 544-547)
    q = np.sqrt(Q[N:2 * N] ** 2 + Q[2 * N:] ** 2)
    fr = g * np.asarray(flat["ManningN_cells"]) ** 2 / flat["k_n"] ** 2 / (h + hs) ** (7.0 / 3.0) * q * q
    S0 = np.asarray(flat["S0_cells"])
    s = s + fr + g * np.abs(Q[:N]) * np.hypot(S0[:N], S0[N:])
    return np.concatenate([s, s, s]) + 1e-300


def random_state_flat(flat, seed, dry_frac=0.05):
    rng = np.random.default_rng(seed)
    N = int(flat["n_cells"])
    h = np.exp(rng.uniform(np.log(1e-4), np.log(10.0), N))
    k = rng.random(N) < dry_frac
    h[k] = rng.choice([5e-4, 1e-3, 9.999e-4], size=int(k.sum()))
    # the reference asserts a positive inlet conveyance (bc_2D.jl:678-680): keep inlet-adjacent cells wet
    ni = int(np.asarray(flat["bc_ptr"])[int(flat["n_inletq"])])
    ic = np.asarray(flat["bc_internal_cells"])[:ni] - int(flat["index_base"])
    h[ic] = np.maximum(h[ic], 0.05)
    sp = rng.uniform(0, 3, N)
    th = rng.uniform(0, 2 * np.pi, N)
    return np.concatenate([h - flat["hstill"], h * sp * np.cos(th), h * sp * np.sin(th)])
