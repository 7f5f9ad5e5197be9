"""TEST INFRASTRUCTURE -- a SECOND, independently written restatement of the reference RHS, used only to cross-check the
primary oracle (oracle/swe_oracle.cpp).  Never imported by the product.

Where the C++ oracle expands the algebra into scalars for speed, this file follows the Julia source statement by statement in
its own shapes: the Roe dissipation as `R_mat * (absLamda * (L_mat * dQ))` with 3x3 arrays, the boundary updates as
per-boundary vectors concatenated and permuted by `update_1d_array`, the cell loop as the comprehension of
`compute_inviscid_fluxes`.  Plain Python / numpy loops: fixture-sized meshes only (a Savannah RHS takes ~0.3 s).
Two transcriptions by different routes that agree to rounding on fuzzed states (tests/test_oracle_literal_cpu.py: all five
wet/dry branches, wall / symmetry / inlet-q / exit-h boundaries, zb / ManningN / Q parameter binding) are the pin for the
branches the reference's committed trajectories do not reach.

Reference (under /root/reference/src):
  fvm/discretization/semi_discretize_swe_2D.jl:18-277   swe_2d_rhs
  fvm/discretization/semi_discretize_swe_2D.jl:281-446  compute_inviscid_fluxes, rearrange_vector_of_vectors
  fvm/discretization/semi_discretize_swe_2D.jl:449-559  compute_source_terms, compute_friction_terms
  fvm/discretization/Riemman_solvers/swe_2D_solvers.jl:4-164  Riemann_2D_Roe
  fvm/boundary_conditions/bc_2D.jl:575-875               process_all_boundaries_2d
  utilities/smooth_functions.jl:10-52, utilities/misc_tools.jl:4-18
  parameters/process_ManningN_2D.jl:71-98, parameters/process_bed_2D.jl:46-66 (through oracle.srh2d_ref.update_bed_data,
  which is pinned bit-exact by the reference's S0_cells_truth)
"""
import numpy as np

from . import srh2d_ref as R

EPS = np.finfo(np.float64).eps          # eps(Float64)

# Every function below also accepts COMPLEX arguments: predicates, clamps and `max` look at the real part only (they are
# piecewise-constant selectors for ForwardDiff / Zygote too) and everything else is complex-analytic, so
# imag(f(x + i e v)) / e with e = 1e-30 is the exact directional derivative -- an AD-free check of the oracle's dual numbers.
re = np.real


def smooth_abs(x):                      # smooth_functions.jl:10-12
    return np.sqrt(x ** 2 + EPS)


def smooth_sqrt(x):                     # smooth_functions.jl:42-44
    return np.sqrt(x + EPS)


def smooth_pow2(x):                     # smooth_functions.jl:50-52 with y = 2
    return (x + EPS) ** 2


def riemann_2d_roe(xiL, hstillL, hL, huL, hvL, zb_L, xiR, hstillR, hR, huR, hvR, zb_R, g, normal, hmin):
    """swe_2D_solvers.jl:4-164, one if / elseif chain; the two 'virtual wall' branches fall through to the main part."""
    nx, ny = normal
    if re(hL) <= hmin and re(hR) <= hmin:                                    # :16-22
        return np.zeros(3)
    elif (re(hL + zb_L) < re(zb_R + hmin)) and (re(hR) <= hmin):             # :23-37
        hR = hL
        huR = -huL
        hvR = -hvL
    elif (re(hR + zb_R) < re(zb_L + hmin)) and (re(hL) <= hmin):             # :39-52
        hL = hR
        huL = -huR
        hvL = -hvR
    elif re(hL) <= hmin:                                                     # :54-64
        h_flux = huR * nx + hvR * ny
        hu_flux = (huR * (huR / hR) + 0.5 * g * smooth_pow2(hR)) * nx + huR * (hvR / hR) * ny
        hv_flux = (hvR * (huR / hR)) * nx + (hvR * (hvR / hR) + 0.5 * g * smooth_pow2(hR)) * ny
        return np.array([h_flux, hu_flux, hv_flux])
    elif re(hR) <= hmin:                                                     # :65-76
        h_flux = huL * nx + hvL * ny
        hu_flux = (huL * (huL / hL) + 0.5 * g * smooth_pow2(hL)) * nx + huL * (hvL / hL) * ny
        hv_flux = (hvL * (huL / hL)) * nx + (hvL * (hvL / hL) + 0.5 * g * smooth_pow2(hL)) * ny
        return np.array([h_flux, hu_flux, hv_flux])
    uL = huL / hL; vL = hvL / hL; uR = huR / hR; vR = hvR / hR               # :79-82
    sqrt_hL = smooth_sqrt(hL); sqrt_hR = smooth_sqrt(hR)                     # :89-90
    hRoe = (hL + hR) / 2.0                                                   # :91 arithmetic average
    uRoe = (sqrt_hL * uL + sqrt_hR * uR) / (sqrt_hL + sqrt_hR)
    vRoe = (sqrt_hL * vL + sqrt_hR * vR) / (sqrt_hL + sqrt_hR)
    unRoe = uRoe * nx + vRoe * ny
    cRoe = smooth_sqrt(g * hRoe)                                             # :95
    over_two_cRoe = 1.0 / 2.0 / cRoe
    R_mat = np.array([[0.0, 1.0, 1.0],                                       # :103-105
                      [ny, uRoe - cRoe * nx, uRoe + cRoe * nx],
                      [-nx, vRoe - cRoe * ny, vRoe + cRoe * ny]])
    L_mat = np.array([[-(uRoe * ny - vRoe * nx), ny, -nx],                   # :107-109
                      [unRoe * over_two_cRoe + 0.5, -nx * over_two_cRoe, -ny * over_two_cRoe],
                      [-unRoe * over_two_cRoe + 0.5, nx * over_two_cRoe, ny * over_two_cRoe]])
    absLamda = np.diag([smooth_abs(unRoe), smooth_abs(unRoe - cRoe), smooth_abs(unRoe + cRoe)])   # :111-113
    dQ = np.array([xiR - xiL, huR - huL, hvR - hvL])                         # :115
    absA_dQ = R_mat @ (absLamda @ (L_mat @ dQ))                              # :118
    pL = 0.5 * g * (smooth_pow2(xiL) + 2.0 * xiL * hstillL)                  # :121-127, xi-form pressure
    pR = 0.5 * g * (smooth_pow2(xiR) + 2.0 * xiR * hstillR)
    xi_flux_L = huL * nx + hvL * ny
    hu_flux_L = (huL * uL + pL) * nx + huL * vL * ny
    hv_flux_L = (hvL * uL) * nx + (hvL * vL + pL) * ny
    xi_flux_R = huR * nx + hvR * ny
    hu_flux_R = (huR * uR + pR) * nx + huR * vR * ny
    hv_flux_R = (hvR * uR) * nx + (hvR * vR + pR) * ny
    return np.array([(xi_flux_L + xi_flux_R - absA_dQ[0]) / 2.0,             # :131-133
                     (hu_flux_L + hu_flux_R - absA_dQ[1]) / 2.0,
                     (hv_flux_L + hv_flux_R - absA_dQ[2]) / 2.0])


def process_all_boundaries_2d(case, h, q_x, q_y, ManningN_cells, zb_cells, inletQ_TotalQ, exitH_WSE):
    """bc_2D.jl:575-875.  Cell / ghost ids in the tables are 1-based like the reference's."""
    kinds = case.bc.kinds
    hs = case.h_small
    updates = []
    for k, b in enumerate(kinds["inletQ"]):                                  # :640-730
        ic = np.asarray(b["internalCellIDs"]) - 1
        L = np.asarray(b["lengths"], dtype=np.float64)
        drywet = np.array([1.0 if re(h[c]) > hs else 0.0 for c in ic])       # :665
        total_A = 0.0
        for i in range(len(ic)):                                             # :674-676, sequential generator sum
            total_A = total_A + L[i] ** (5.0 / 3.0) * h[ic[i]] / ManningN_cells[ic[i]] * drywet[i]
        assert total_A > 1e-10                                               # :678-680
        velocity_normals = inletQ_TotalQ[k] / total_A * L ** (2.0 / 3.0) / ManningN_cells[ic]   # :690-691
        fn = np.asarray(b["normals"])
        updates.append((h[ic], -h[ic] * velocity_normals * fn[:, 0] * drywet, -h[ic] * velocity_normals * fn[:, 1] * drywet))
    for k, b in enumerate(kinds["exitH"]):                                   # :748-773
        ic = np.asarray(b["internalCellIDs"]) - 1
        d = exitH_WSE[k] - zb_cells[ic]
        updates.append((np.where(re(d) > hs, d, hs), q_x[ic], q_y[ic]))      # max.(h_small, WSE - zb), :763-764
    for b in kinds["wall"]:                                                  # :777-799
        ic = np.asarray(b["internalCellIDs"]) - 1
        updates.append((h[ic], -q_x[ic], -q_y[ic]))
    for b in kinds["symm"]:                                                  # :803-834
        ic = np.asarray(b["internalCellIDs"]) - 1
        fn = np.asarray(b["normals"])
        v_dot_n = q_x[ic] * fn[:, 0] + q_y[ic] * fn[:, 1]
        updates.append((h[ic], q_x[ic] - 2.0 * v_dot_n * fn[:, 0], q_y[ic] - 2.0 * v_dot_n * fn[:, 1]))
    all_h = np.concatenate([u[0] for u in updates])                          # :837-858
    all_qx = np.concatenate([u[1] for u in updates])
    all_qy = np.concatenate([u[2] for u in updates])
    idx = np.asarray(case.bc.all_boundary_ghost_indices) - 1                 # update_1d_array, misc_tools.jl:4-18
    return all_h[idx], all_qx[idx], all_qy[idx]


def compute_friction_terms(h, q_x, q_y, ManningN_cells, g, k_n, h_small):
    """semi_discretize_swe_2D.jl:544-547, Manning branch, evaluated left to right like the Julia expression."""
    mag = smooth_sqrt(q_x ** 2 + q_y ** 2)
    friction_x = g * ManningN_cells ** 2 / k_n ** 2 / (h + h_small) ** (7.0 / 3.0) * mag * q_x
    friction_y = g * ManningN_cells ** 2 / k_n ** 2 / (h + h_small) ** (7.0 / 3.0) * mag * q_y
    return friction_x, friction_y


def swe_2d_rhs(case, Q, params_vector=None, active_param_name=""):
    """semi_discretize_swe_2D.jl:18-277 for the constant-Manning / inversion / sensitivity configurations."""
    m = case.mesh
    N = m.numOfCells
    g, k_n, h_small = case.g, case.k_n, case.h_small
    xi, q_x, q_y = Q[:N], Q[N:2 * N], Q[2 * N:3 * N]                          # :93-95
    h = xi + case.hstill                                                     # :101
    h = np.where(re(h) <= h_small, h_small, h)                               # :104-106 (the test of q uses the clamped h)
    q_x = np.where(re(h) <= h_small, 0.0, q_x)
    q_y = np.where(re(h) <= h_small, 0.0, q_y)
    ManningN_cells, zb_cells, zb_ghost, S0_cells = case.ManningN_cells, case.zb_cells, case.zb_ghost, case.S0_cells
    inletQ_TotalQ, exitH_WSE = case.bc.inletQ_TotalQ, case.bc.exitH_WSE
    if active_param_name == "zb":                                            # :114-126
        zb_cells = np.asarray(params_vector)
        zb_ghost, _zb_faces, S0_cells = R.update_bed_data(m, re(zb_cells))
        if np.iscomplexobj(zb_cells):                                        # update_bed_data is linear in zb
            zg_i, _zf_i, S0_i = R.update_bed_data(m, np.imag(zb_cells))
            zb_ghost, S0_cells = zb_ghost + 1j * zg_i, S0_cells + 1j * S0_i
    elif active_param_name == "ManningN":                                    # :153-161, process_ManningN_2D.jl:88
        ManningN_cells = np.array([params_vector[int(mid)] for mid in case.matID])
    elif active_param_name == "Q":                                           # :190-199
        inletQ_TotalQ = np.asarray(params_vector)
    h_ghost, q_x_ghost, q_y_ghost = process_all_boundaries_2d(case, h, q_x, q_y, ManningN_cells, zb_cells, inletQ_TotalQ, exitH_WSE)
    xi_ghost = h_ghost - case.hstill_ghost                                   # :220
    # compute_inviscid_fluxes, :281-434
    updates_inviscid_cells = []
    for iCell in range(N):
        flux_sum = np.zeros(3)
        for iFace in range(int(m.cellNodesCount[iCell])):
            faceID = int(m.cellFacesList[iCell, iFace])
            right = int(m.cellNeighbors[iCell][iFace]) - 1
            face_normal = m.cell_normals[iCell][iFace]
            if not m.bFace_is_boundary[faceID - 1]:
                R_state = (xi[right], case.hstill[right], h[right], q_x[right], q_y[right], zb_cells[right])
            else:
                R_state = (xi_ghost[right], case.hstill_ghost[right], h_ghost[right], q_x_ghost[right], q_y_ghost[right], zb_ghost[right])
            flux = riemann_2d_roe(xi[iCell], case.hstill[iCell], h[iCell], q_x[iCell], q_y[iCell], zb_cells[iCell], *R_state,
                                  g, face_normal, h_small)
            flux_sum = flux_sum + flux * m.face_lengths[faceID - 1]          # :387
        updates_inviscid_cells.append(-flux_sum / m.cell_areas[iCell])      # :415
    upd = np.array(updates_inviscid_cells)
    updates_inviscid = np.concatenate([upd[:, 0], upd[:, 1], upd[:, 2]])     # rearrange_vector_of_vectors, :437-446
    # compute_friction_terms :544-547 (evaluated left to right), compute_source_terms :463-478
    friction_x, friction_y = compute_friction_terms(h, q_x, q_y, ManningN_cells, g, k_n, h_small)
    above_small_h = (re(h) > h_small).astype(np.float64)
    source_x = above_small_h * (g * xi * S0_cells[:, 0] - friction_x)
    source_y = above_small_h * (g * xi * S0_cells[:, 1] - friction_y)
    return updates_inviscid + np.concatenate([np.zeros(N), source_x, source_y])


def custom_ODE_update_cells(case, Q, params_vector, dt, active_param_name=""):
    """ode_solvers/custom_ODE_solvers.jl:5-33 -- explicit Euler with the reference's dry mask, which tests the NEW xi (the first
    N entries of Q are xi, not h) against h_small and then sets xi := h_small, q := 0."""
    N = case.mesh.numOfCells
    Q_new = Q + dt * swe_2d_rhs(case, Q, params_vector, active_param_name)   # :11-16
    dry_mask = re(Q_new[:N]) < case.h_small                                  # :19
    return np.concatenate([np.where(dry_mask, case.h_small, Q_new[:N]),      # :22-26
                           np.where(dry_mask, 0.0, Q_new[N:2 * N]),
                           np.where(dry_mask, 0.0, Q_new[2 * N:3 * N])])
