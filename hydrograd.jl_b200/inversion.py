"""Inversion loss and its gradient through the device-resident forward + adjoint sweep (BASELINE config C4).

Mirrors `compute_loss_inversion` (src/applications/inversion/swe_2D_inversion.jl:247-483; loss terms 388-467, bound
loss 738-748).  The reference differentiates this with Zygote/ForwardDiff through the ODE solve; here the ODE part is
the explicit Euler stepper of custom_ODE_solvers.jl run on the GPU and its hand-written discrete adjoint
(hg_euler_adjoint).  Only O(N) host arithmetic happens here: the loss itself and its cotangent d loss / d Q(T).
Not covered: the slope-regularisation term `calc_slope_loss` (750-772), which only acts when zb is inverted.
"""
from __future__ import annotations

import numpy as np


def compute_bound_loss(params, lower, upper):
    """swe_2D_inversion.jl:738-748 -> (loss, d loss / d params)."""
    rng = upper - lower
    lo = np.maximum(0.0, lower - params) / rng
    up = np.maximum(0.0, params - upper) / rng
    return float((lo ** 2 + up ** 2).sum()), (-2.0 * lo + 2.0 * up) / rng


def loss_terms(Q_T, params, observed, flat, active_param_name, bWSE=True, buv=True, bound=None):
    """Loss at the final state and its cotangents.  observed: dict(WSE_truth, u_truth, v_truth, zb_cell_truth).

    Returns (loss_total, parts, lambda_T[3N], dloss_dp_direct[len(params)])."""
    N = int(flat["n_cells"])
    eps = np.sqrt(np.finfo(np.float64).eps)
    xi, qx, qy = Q_T[:N], Q_T[N:2 * N], Q_T[2 * N:]
    hstill = np.asarray(flat["hstill"])
    zb = np.asarray(params) if active_param_name == "zb" else np.asarray(observed["zb_cell_truth"])
    h = xi + hstill
    u, v = qx / (h + eps), qy / (h + eps)
    wse = h + zb
    lam = np.zeros(3 * N)
    dp = np.zeros(0 if params is None else len(params))
    parts = dict(WSE=0.0, uv=0.0, bound=0.0)
    if bWSE:
        rngw = observed["WSE_truth"].max() - observed["WSE_truth"].min()
        r = (wse - observed["WSE_truth"]) / (rngw + eps)
        parts["WSE"] = float((r ** 2).sum() / N)
        g = 2.0 * r / (rngw + eps) / N
        lam[:N] += g
        if active_param_name == "zb":
            dp += g
    if buv:
        scale = max(np.sqrt(observed["u_truth"] ** 2 + observed["v_truth"] ** 2).max(), np.finfo(np.float64).eps)
        ru, rv = (u - observed["u_truth"]) / scale, (v - observed["v_truth"]) / scale
        parts["uv"] = float(((ru ** 2).sum() + (rv ** 2).sum()) / N)
        gu, gv = 2.0 * ru / scale / N, 2.0 * rv / scale / N
        lam[N:2 * N] += gu / (h + eps)
        lam[2 * N:] += gv / (h + eps)
        lam[:N] += -(gu * u + gv * v) / (h + eps)
    if bound is not None and params is not None:
        lb, dlb = compute_bound_loss(np.asarray(params), bound[0], bound[1])
        parts["bound"] = lb / N
        dp += dlb / N
    return parts["WSE"] + parts["uv"] + parts["bound"], parts, lam, dp


def loss_and_gradient(ctx, flat, Q0, params, active_param_name, observed, dt, nsteps, method="Euler", **kw):
    """One optimiser iteration's work: forward sweep, loss, discrete adjoint sweep -> (loss, parts, d loss/d params).
    method: "Euler" (the reference's customized solver, custom_ODE_solvers.jl), "RK4" or "Tsit5" (fixed-step SciML solvers), or
    "Tsit5_adaptive" -- the configuration of the reference's inversion cases (inversion_ode_solver_options: Tsit5(), adaptive):
    adaptive solve over [0, dt * nsteps] with initial step dt, then the discrete adjoint over the accepted steps, whose sizes are
    constants of the differentiation exactly as for ForwardDiff through the reference's solve (hg_last_steps + hg_rk_adjoint_steps)."""
    # the terminal cotangent needs Q(T) first, so run forward once; the adjoint call repeats the forward sweep with checkpoints
    ctx.set_params(params, active_param_name)
    ctx.set_state(Q0)
    if method == "Euler":
        ctx.step_euler(dt, nsteps)
    elif method == "RK4":
        ctx.step_rk4(dt, nsteps)
    elif method == "Tsit5":
        ctx.solve_tsit5(0.0, dt * nsteps, dt, adaptive=False)
    elif method == "Tsit5_adaptive":
        ctx.solve_tsit5(0.0, dt * nsteps, dt, adaptive=True, abstol=1e-6, reltol=1e-3)
        steps = ctx.last_steps()
    else:
        raise ValueError(f"unknown method {method}")
    Q_T = ctx.get_state()
    loss, parts, lam, dp = loss_terms(Q_T, params, observed, flat, active_param_name, **kw)
    if method == "Euler":
        _, _, pbar = ctx.euler_adjoint(Q0, lam, dt, nsteps, params, active_param_name)
    elif method == "Tsit5_adaptive":
        _, _, pbar = ctx.rk_adjoint_steps("Tsit5", Q0, lam, steps, params, active_param_name)
    else:
        _, _, pbar = ctx.rk_adjoint(method, Q0, lam, dt, nsteps, params, active_param_name)
    return loss, parts, pbar + dp


def compute_loss_UDE(ctx, flat, Q0, theta, observed, dt, nsteps, method="Tsit5", bWSE=True, buv=True):
    """compute_loss_UDE (src/applications/UDE/swe_2D_UDE.jl:522-599) and its gradient with respect to the network parameters:
    the same WSE / velocity mismatch terms as the inversion loss (UDE_bWSE_loss, UDE_b_uv_loss), no bound term.  The model must
    have been set with ctx.set_ude_model; every RHS of the forward and adjoint sweeps re-evaluates n = NN_theta(state) on the
    device.  Returns (loss_total, parts, d loss / d theta)."""
    return loss_and_gradient(ctx, flat, Q0, theta, "UDE", observed, dt, nsteps, method=method, bWSE=bWSE, buv=buv)
