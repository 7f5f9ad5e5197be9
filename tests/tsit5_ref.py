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
  * saveat: OrdinaryDiffEq interpolates with Tsit5's dense output.  Here (and in hg_solve_tsit5) the save times are tstops --
    the step is clipped to land on them and the controller's proposal is restored afterwards -- because the dense-output
    polynomials are not restated.  Both are within the integration tolerance of each other.
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
BETA2, BETA1, GAMMA, QMIN, QMAX, QOLDINIT = 2.0 / 25.0, 7.0 / 50.0, 0.9, 0.2, 10.0, 1e-4


def solve(rhs, u0, t0, t1, dt, adaptive=True, abstol=1e-6, reltol=1e-3, t_save=()):
    """Returns (u(t1), [u(t) for t in t_save], stats).  rhs(u) -> du/dt (autonomous: the reference RHS ignores t)."""
    u = np.array(u0, dtype=np.float64)
    t = float(t0)
    stops = sorted(set([float(x) for x in t_save if t0 < x <= t1] + [float(t1)]))
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
            if not adaptive:
                u, t, k[0] = unew, t + h, k[6]
                n_acc += 1
                continue
            ut = np.zeros_like(u)
            for i in range(7):
                ut += (h * BTILDE[i]) * k[i]
            eest = float(np.sqrt(np.mean((ut / (abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol)) ** 2)))
            if eest == 0.0:
                q11, q = 0.0, 1.0 / QMAX
            else:
                q11 = eest ** BETA1
                q = q11 / qold ** BETA2
                q = max(1.0 / QMAX, min(1.0 / QMIN, q / GAMMA))
            if eest <= 1.0:
                u, t, k[0] = unew, t + h, k[6]
                n_acc += 1
                qold = max(eest, QOLDINIT)
                prop = h / q
                # a step clipped by a stop does not shrink the controller's own proposal
                dt_ctrl = max(prop, dt_ctrl) if clipped else prop
            else:
                n_rej += 1
                dt_ctrl = h / min(1.0 / QMIN, q11 / GAMMA)
        t = ts
        saves[ts] = u.copy()
    return u, [saves[float(x)] for x in t_save if float(x) in saves], dict(accepted=n_acc, rejected=n_rej, rhs=n_rhs)
