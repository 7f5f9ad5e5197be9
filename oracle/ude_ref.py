"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of the reference's UDE closure for Manning's n.

Follows
  * update_ManningN_UDE, src/parameters/process_ManningN_2D.jl:216-272 (input normalisation to [-1, 1] :233-240, the batch
    call :257-263, `vec(outputs)` :271);
  * the call site src/fvm/discretization/semi_discretize_swe_2D.jl:165-178 (u = q_x ./ h, v = q_y ./ h, Umag = sqrt.(u.^2 .+ v.^2)
    of the CLAMPED state, :101-106);
  * create_NN_model, src/UDE/process_UDE.jl:29-39: for every hidden layer `Dense(in, width, act)` then `LayerNorm(width)`, a final
    `Dense(in, 1)` and the wrapper `lo .+ (hi - lo) .* sigmoid(x)`; activations of get_activation (:91-105).

Third-party arithmetic absent from /root/reference: Lux.jl (Lux = "1.2.3", Project.toml:88).  Restated from the published
layer definitions: Dense y = act.(W x .+ b) with W of size (out, in); LayerNorm y = (x .- mean) ./ sqrt.(var .+ epsilon) .* scale
.+ bias with the uncorrected variance and epsilon = 1f-5.  `ln_mode` selects what the statistics run over: "whole" = every
entry of the (width x N) array (Lux's documented default `dims = Colon()`), "cell" = the hidden units of each cell, "none".

PARITY UNPINNED for this closure: the reference commits no network parameters (they live in .jld2 files that are not part of
the repository), so no fixture can pin the restatement; the CUDA path is compared with this file only.

Everything works on complex arrays too, which gives derivatives by complex-step differentiation (exact to rounding for these
analytic functions) without a second, hand-written derivative that could share a mistake with the product.
"""
import numpy as np

ACT = {"identity": 0, "relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4, "softplus": 5}
LN = {"none": 0, "cell": 1, "whole": 2}
LUX_EPS = float(np.float32(1e-5))


def _act(name, z):
    if name == "relu":
        return np.where(z.real > 0, z, 0 * z)
    if name == "leakyrelu":
        return np.where(z.real > 0, z, 0.01 * z)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if name == "tanh":
        return np.tanh(z)
    if name == "softplus":
        return np.log1p(np.exp(-np.where(z.real > 0, z, -z))) + np.where(z.real > 0, z, 0 * z)
    return z


class Model:
    """hidden: list of widths; acts: list of activation names; theta layout = the ComponentArray of Lux.setup(Chain(...)):
    layer_1 = Dense (weight[out x in] column-major, bias[out]), layer_2 = LayerNorm (bias[width], scale[width]), ...,
    last Dense (weight[1 x in], bias[1])."""

    def __init__(self, choice, hidden, acts, ln_mode, h_bounds, output_bounds, umag_bounds=(0.0, 1.0), ks_bounds=(0.0, 1.0),
                 eps=LUX_EPS):
        assert choice in ("ManningN_h", "ManningN_h_Umag_ks") and len(hidden) == len(acts)
        self.choice, self.hidden, self.acts, self.ln_mode, self.eps = choice, list(hidden), list(acts), ln_mode, eps
        self.n_in = 1 if choice == "ManningN_h" else 3
        self.bounds = [tuple(h_bounds), tuple(umag_bounds), tuple(ks_bounds)]
        self.out = tuple(output_bounds)
        off, n_prev = 0, self.n_in
        self.off_w, self.off_b, self.off_g, self.off_be = [], [], [], []
        for w in self.hidden:
            self.off_w.append(off); off += w * n_prev
            self.off_b.append(off); off += w
            if ln_mode != "none":
                self.off_be.append(off); off += w
                self.off_g.append(off); off += w
            else:
                self.off_be.append(0); self.off_g.append(0)
            n_prev = w
        self.off_w.append(off); off += n_prev
        self.off_b.append(off); off += 1
        self.n_params = off

    def init_theta(self, rng):
        """Glorot-uniform weights, zero biases, unit scales perturbed a little so that every parameter matters."""
        th = np.zeros(self.n_params)
        n_prev = self.n_in
        for l, w in enumerate(self.hidden + [1]):
            lim = np.sqrt(6.0 / (w + n_prev))
            th[self.off_w[l]:self.off_w[l] + w * n_prev] = rng.uniform(-lim, lim, w * n_prev)
            th[self.off_b[l]:self.off_b[l] + w] = rng.uniform(-0.3, 0.3, w)
            if l < len(self.hidden) and self.ln_mode != "none":
                th[self.off_g[l]:self.off_g[l] + w] = 1.0 + rng.uniform(-0.3, 0.3, w)
                th[self.off_be[l]:self.off_be[l] + w] = rng.uniform(-0.3, 0.3, w)
            n_prev = w
        return th

    # ---- update_ManningN_UDE on the raw state --------------------------------------------------------------------------
    def inputs(self, Q, hstill, ks, h_small):
        N = hstill.size
        xi, qx, qy = Q[:N], Q[N:2 * N], Q[2 * N:]
        h0 = xi + hstill
        dry = h0.real <= h_small
        h = np.where(dry, h_small + 0 * h0, h0)
        u = np.where(dry, 0 * qx, qx / h)
        v = np.where(dry, 0 * qy, qy / h)
        umag = np.sqrt(u * u + v * v)
        rows = [2.0 * (h - self.bounds[0][0]) / (self.bounds[0][1] - self.bounds[0][0]) - 1.0]
        if self.n_in == 3:
            rows.append(2.0 * (umag - self.bounds[1][0]) / (self.bounds[1][1] - self.bounds[1][0]) - 1.0)
            rows.append(2.0 * (ks - self.bounds[2][0]) / (self.bounds[2][1] - self.bounds[2][0]) - 1.0 + 0 * h)
        return np.stack(rows)                                  # (n_in, N), like vcat(h', Umag', ks')

    def network(self, X, th):
        a, n_prev = X, self.n_in
        for l, w in enumerate(self.hidden):
            W = th[self.off_w[l]:self.off_w[l] + w * n_prev].reshape(n_prev, w).T      # column-major (out, in)
            b = th[self.off_b[l]:self.off_b[l] + w]
            y = _act(self.acts[l], W @ a + b[:, None])
            if self.ln_mode != "none":
                ax = None if self.ln_mode == "whole" else 0
                mu = y.mean(axis=ax, keepdims=True)
                var = ((y - mu) ** 2).mean(axis=ax, keepdims=True)
                g = th[self.off_g[l]:self.off_g[l] + w]
                be = th[self.off_be[l]:self.off_be[l] + w]
                y = (y - mu) / np.sqrt(var + self.eps) * g[:, None] + be[:, None]
            a, n_prev = y, w
        L = len(self.hidden)
        z = th[self.off_w[L]:self.off_w[L] + n_prev] @ a + th[self.off_b[L]]
        return self.out[0] + (self.out[1] - self.out[0]) * (1.0 / (1.0 + np.exp(-z)))

    def manning(self, Q, th, hstill, ks, h_small):
        """ManningN_cells[N] of update_ManningN_UDE for the state Q and the parameters th."""
        return self.network(self.inputs(Q, hstill, ks, h_small), th)

    # ---- derivatives by complex step ---------------------------------------------------------------------------------------
    def pullback(self, Q, th, hstill, ks, h_small, nbar, step=1e-30):
        """(d n / d Q)^T nbar [3N] and (d n / d theta)^T nbar [n_params], one complex-step evaluation per direction."""
        Q = np.asarray(Q, dtype=np.float64)
        N = hstill.size
        Qbar, thbar = np.zeros(3 * N), np.zeros(self.n_params)
        for k in range(3 * N):
            Qc = Q.astype(np.complex128)
            Qc[k] += 1j * step
            Qbar[k] = (self.manning(Qc, th, hstill, ks, h_small).imag / step) @ nbar
        for k in range(self.n_params):
            tc = th.astype(np.complex128)
            tc[k] += 1j * step
            thbar[k] = (self.manning(Q, tc, hstill, ks, h_small).imag / step) @ nbar
        return Qbar, thbar


class UdeRhs:
    """swe_2d_rhs with settings.bPerform_UDE (semi_discretize_swe_2D.jl:165-178): ManningN_cells = update_ManningN_UDE(state,
    theta), then the ordinary RHS.  Built from the C++ oracle with one Manning zone per cell (so that its dual-number
    derivative with respect to the zone values is d/d ManningN_cells) and the network above."""

    def __init__(self, flat, model, ks=None):
        from .oracle import Oracle
        f = dict(flat)
        self.N = int(f["n_cells"])
        f["matID_cells"] = np.arange(self.N, dtype=np.int64)
        f["n_mat"] = self.N
        self.o = Oracle(f)
        self.m = model
        self.hstill = np.asarray(f["hstill"], dtype=np.float64)
        self.hs = float(f["h_small"])
        self.ks = np.ones(self.N) if ks is None else np.asarray(ks, dtype=np.float64)

    def manning(self, Q, th):
        return self.m.manning(np.asarray(Q, dtype=np.float64), th, self.hstill, self.ks, self.hs)

    def rhs(self, Q, th):
        return self.o.rhs(Q, self.manning(Q, th), 2)

    def vjp(self, Q, th, lam):
        """(Qbar, thetabar) of lam . rhs(Q, theta): dual-number Jacobian of the RHS, complex-step Jacobian of the network."""
        Qbar, nbar = self.o.vjp_bruteforce(Q, lam, self.manning(Q, th), 2)
        Qb, tb = self.m.pullback(Q, th, self.hstill, self.ks, self.hs, nbar)
        return Qbar + Qb, tb, nbar

    def jvp(self, Q, th, v, w, step=1e-30):
        """d/d eps rhs(Q + eps v, theta + eps w): one complex-step pass through the network, one dual-number pass through the RHS."""
        Q = np.asarray(Q, dtype=np.float64)
        dn = self.m.manning(Q + 1j * step * np.asarray(v), th + 1j * step * np.asarray(w), self.hstill, self.ks, self.hs).imag / step
        return self.o.jvp(Q, v, self.manning(Q, th), dn, 2)[1]
