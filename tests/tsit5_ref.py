"""Host restatement of the time integrator the reference's forward / sensitivity drivers use by default:
`solve(prob, Tsit5(), adaptive=..., dt=dt, saveat=t_save; abstol=1e-6, reltol=1e-3)` (swe_2D_forward_simulation.jl:38-41,
swe_2D_sensitivity.jl:38-43).  Test infrastructure: drives ANY rhs callable (the oracle in the CPU tests).

Third-party algorithm, absent from /root/reference: OrdinaryDiffEq.jl (DifferentialEquations = "7.15.0", Project.toml:74;
OrdinaryDiffEq itself unpinned).  Restated from its published form:
  * tableau: Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the first column simplifying assumption";
    the coefficients below satisfy the row-sum, order-5 and embedded order-4 conditions to 1e-16 (test_tsit5_tableau);
  * error estimate: utilde = dt * sum(btilde_i k_i), EEst = sqrt(mean((utilde / (abstol + max(|u_prev|, |u|) reltol))^2));
  * step-size control: OrdinaryDiffEq's PIController with the explicit-RK defaults beta2 = 2/(5 p), beta1 = 7/(10 p) (p = 5),
    gamma = 9/10, qmin = 1/5, qmax = 10, qoldinit = 1e-4; accept when EEst <= 1;
    The two powers are evaluated by DiffEqBase's `fastpow` in the OrdinaryDiffEq generation the reference ran
    (DifferentialEquations 7.15): Float32 arithmetic, log2 by a rational approximation on the significand (Goldberg's "fast
    approximate logarithms", (x-1)(a(x-1)+b)/((x-1)+c) after reducing the significand to [0.75, 1.5)), then exp2 -- about
    1e-5 relative error, i.e. step sizes that differ from the exact-power controller in the fifth digit.  `pow="fastpow"`
    restates it; it is what makes the reference's saved trajectories reproducible to 1e-11 ... 1e-9 over the first saves
    instead of 1e-9 ... 1e-7 (tests/test_oracle_golden.py::test_reference_trajectory_hard_pin).  `pow="exact"` (default, and what
    hg_solve_tsit5 does) uses the correctly rounded power;
  * saveat: OrdinaryDiffEq does NOT stop at the save times; after every accepted step it evaluates Tsit5's fourth-order
    dense output u(t + theta h) = u + h sum_i b_i(theta) k_i at the save times the step has passed (savevalues!), and
    copies u when a save time coincides with the step end.  `saveat="interp"` (hg_solve_tsit5_dense) restates that; the
    polynomials b_i(theta) below satisfy the continuous order-4 conditions identically in theta and b_i(1) = the weights
    of the fifth-order solution (test_tsit5_dense_output_conditions).  `saveat="stop"` (hg_solve_tsit5) treats the save
    times as tstops instead -- the step is clipped to land on them and the controller's proposal is restored afterwards.
    Both are within the integration tolerance of each other; only "interp" reproduces OrdinaryDiffEq's step sequence.
"""
import numpy as np

C = (0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0)
A = ((),
     (0.161,),
     (-0.008480655492356989, 0.335480655492357),
     (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
     (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
     (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
     (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774))
BTILDE = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552,
          -0.45808210592918697, 0.015151515151515152)
# dense output: b_i(theta) = sum_p INTERP[i][p] theta^(p+1)   (OrdinaryDiffEq's Tsit5Interp r_ip)
INTERP = ((1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216),
          (0.0, 0.13169999999999998, -0.2234, 0.1017),
          (0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253),
          (0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902),
          (0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928),
          (0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661),
          (0.0, 1.5, -4.0, 2.5))
BETA2, BETA1, GAMMA, QMIN, QMAX, QOLDINIT = 2.0 / 25.0, 7.0 / 50.0, 0.9, 0.2, 10.0, 1e-4


def interp_weights(theta):
    """b_i(theta), i = 1..7, evaluated by Horner like OrdinaryDiffEq's @evalpoly."""
    return [theta * (r[0] + theta * (r[1] + theta * (r[2] + theta * r[3]))) for r in INTERP]


def fastpow(x, y):
    """DiffEqBase.fastpow(x::Float64, y::Float64): Float32(x), Float32(y), fastlog2, exp2 (restated; see the header)."""
    if x == 0.0:
        return 0.0
    f32 = np.float32
    a, b, c = f32(0.338953), f32(2.198599), f32(1.523692)
    ux = int(np.array([x], dtype=np.float32).view(np.uint32)[0])
    ex = (ux & 0x7F800000) >> 23
    if ux & 0x00400000:                      # significand >= 1.5: halve it (exponent field 126), compensate in the exponent
        sig = np.array([(ux & 0x007FFFFF) | 0x3F000000], dtype=np.uint32).view(np.float32)[0]
        fexp = f32(ex - 126)
    else:
        sig = np.array([(ux & 0x007FFFFF) | 0x3F800000], dtype=np.uint32).view(np.float32)[0]
        fexp = f32(ex - 127)
    sg = f32(sig - f32(1.0))
    lg2 = f32(fexp + f32(f32(sg * f32(f32(a * sg) + b)) / f32(sg + c)))
    return float(np.exp2(f32(f32(y) * lg2)))


def default_norm(ut, u, unew, abstol, reltol):
    """EEst of a real state: ODE_DEFAULT_NORM of calculate_residuals(utilde, uprev, u, abstol, reltol)."""
    return float(np.sqrt(np.mean((ut / (abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol)) ** 2)))


def dual_norm(ut, u, unew, abstol, reltol):
    """EEst when the state is a vector of ForwardDiff.Dual numbers (what ForwardDiff.jacobian over `solve` produces in
    swe_2D_sensitivity.jl:34-80), stored here as an array of shape (1 + K, n): row 0 the values, rows 1..K the partials.
    DiffEqBase's ForwardDiff rules: the internal norm of one Dual is sqrt(value^2 + sum(partials^2)), so the residual of
    entry i is utilde_i / (abstol + max(|uprev_i|, |u_i|) reltol) with those norms in the (real) denominator, and the norm of
    the residual vector is sqrt(sum over values AND partials of the squares / totallength) with totallength = n (1 + K)."""
    den = abstol + np.maximum(np.sqrt((u * u).sum(0)), np.sqrt((unew * unew).sum(0))) * reltol
    r = ut / den[None, :]
    return float(np.sqrt((r * r).sum() / r.size))


def solve(rhs, u0, t0, t1, dt, adaptive=True, abstol=1e-6, reltol=1e-3, t_save=(), saveat="stop", norm=default_norm, pow="exact",
          record=None):
    """Returns (u(t1), [u(t) for t in t_save], stats).  rhs(u) -> du/dt (autonomous: the reference RHS ignores t).
    record: a list that receives (t, h) of every accepted step."""
    power = {"exact": lambda x, y: x ** y, "fastpow": fastpow}[pow]
    assert saveat in ("stop", "interp")
    u = np.array(u0, dtype=np.float64)
    t = float(t0)
    stops = sorted(set([float(x) for x in t_save if t0 < x <= t1 and saveat == "stop"] + [float(t1)]))
    pending = sorted(float(x) for x in t_save if t0 < x <= t1) if saveat == "interp" else []
    saves = {}
    if any(float(x) == t0 for x in t_save):
        saves[float(t0)] = u.copy()
    dt_ctrl, qold = float(dt), QOLDINIT
    n_acc = n_rej = n_rhs = 0
    k = [None] * 7
    k[0] = rhs(u); n_rhs += 1
    for ts in stops:
        while t < ts:
            h = min(dt_ctrl, ts - t)
            clipped = h < dt_ctrl
            if ts - (t + h) < 1e-12 * max(1.0, abs(ts)):       # do not leave a sliver before the stop
                h = ts - t
            for i in range(1, 7):
                y = u.copy()
                for j, a in enumerate(A[i]):
                    y += (h * a) * k[j]
                if i < 6:
                    k[i] = rhs(y); n_rhs += 1
            unew = y
            k[6] = rhs(unew); n_rhs += 1
            tnew = ts if h == ts - t else t + h

            def dense_saves():
                while pending and pending[0] <= tnew:
                    s_ = pending.pop(0)
                    if s_ == tnew:
                        saves[s_] = unew.copy()
                    else:
                        v = u.copy()
                        for b, kk in zip(interp_weights((s_ - t) / h), k):
                            v += (h * b) * kk
                        saves[s_] = v
            if not adaptive:
                if record is not None:
                    record.append((t, h))
                dense_saves()
                u, t, k[0] = unew, tnew, k[6]
                n_acc += 1
                continue
            ut = np.zeros_like(u)
            for i in range(7):
                ut += (h * BTILDE[i]) * k[i]
            eest = norm(ut, u, unew, abstol, reltol)
            if eest == 0.0:
                q11, q = 0.0, 1.0 / QMAX
            else:
                q11 = power(eest, BETA1)
                q = q11 / power(qold, BETA2)
                q = max(1.0 / QMAX, min(1.0 / QMIN, q / GAMMA))
            if eest <= 1.0:
                if record is not None:
                    record.append((t, h))
                dense_saves()
                u, t, k[0] = unew, tnew, k[6]
                n_acc += 1
                qold = max(eest, QOLDINIT)
                prop = h / q
                # a step clipped by a stop does not shrink the controller's own proposal
                dt_ctrl = max(prop, dt_ctrl) if clipped else prop
            else:
                n_rej += 1
                dt_ctrl = h / min(1.0 / QMIN, q11 / GAMMA)
        t = ts
        if saveat == "stop" or ts in [float(x) for x in t_save]:
            saves.setdefault(ts, u.copy())
    return u, [saves[float(x)] for x in t_save if float(x) in saves], dict(accepted=n_acc, rejected=n_rej, rhs=n_rhs)


def replay(step, steps, t_save):
    """Re-runs a recorded step sequence [(t, h)] with `step(t0, t1, h, saves_inside) -> [state at each save]` (one fixed
    Tsit5 step from the current state, dense output at the save times inside it) and returns the saved states in order."""
    pending = sorted(float(x) for x in t_save)
    out = []
    for t, h in steps:
        t1 = t + h
        inside = []
        while pending and pending[0] <= t1:
            inside.append(pending.pop(0))
        out.extend(step(t, t1, h, inside))
        if not pending:
            break
    return out
