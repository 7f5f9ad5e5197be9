// Per-cell arithmetic of the UDE network (hg_set_ude_model): Manning's n = NN_theta(h, |U|, ks) and its reverse sweep.
// Restates update_ManningN_UDE (parameters/process_ManningN_2D.jl:216-272) and the Lux chain of create_NN_model
// (UDE/process_UDE.jl:29-39): [Dense(act) -> LayerNorm] x n_hidden -> Dense(1) -> lo + (hi - lo) sigmoid.
// Plain host/device functions without any CUDA dependency, so that the very same code is compiled by nvcc into the
// kernels of hg_ude.cu and by g++ into the CPU check of tests/test_ude_cpu.py (tests/ude_host.cpp).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/hydrograd_b200.h"

#if defined(__CUDACC__)
#define HG_HD __host__ __device__ __forceinline__
#else
#define HG_HD inline
#endif

namespace hg {
namespace ude {

constexpr int MAXH = HG_UDE_MAX_HIDDEN, MAXW = HG_UDE_MAX_WIDTH;
constexpr int MAXP = 3 * MAXW + 3 * MAXW + (MAXH - 1) * (MAXW * MAXW + 3 * MAXW) + MAXW + 1;   // 233 with 3 x 8

// the descriptor in kernel-argument form (by value, ~250 bytes)
struct Model {
  int32_t n_in = 0, n_hidden = 0, ln_mode = 0, n_params = 0;
  int32_t width[MAXH] = {}, act[MAXH] = {};
  int32_t off_w[MAXH + 1] = {}, off_b[MAXH + 1] = {}, off_g[MAXH] = {}, off_be[MAXH] = {};
  double eps = 0.0, in_lo[3] = {}, in_den[3] = {}, out_lo = 0.0, out_span = 0.0;
};

// network inputs of one cell from its raw state: clamp of semi_discretize_swe_2D.jl:101-106, u = q/h, |U| = sqrt(u^2 + v^2)
// (:166-169), normalisation to [-1, 1] (process_ManningN_2D.jl:233-240)
struct Inputs {
  double x[3], h, u, v, umag;
  bool dry;
};
HG_HD void inputs(const Model& m, double xi, double qx, double qy, double hst, double ks, double hs, Inputs& in) {
  const double h0 = xi + hst;
  in.dry = h0 <= hs;
  in.h = in.dry ? hs : h0;
  in.u = in.dry ? 0.0 : qx / in.h;
  in.v = in.dry ? 0.0 : qy / in.h;
  in.umag = sqrt(in.u * in.u + in.v * in.v);
  in.x[0] = 2.0 * (in.h - m.in_lo[0]) / m.in_den[0] - 1.0;
  in.x[1] = in.x[2] = 0.0;
  if (m.n_in == 3) {
    in.x[1] = 2.0 * (in.umag - m.in_lo[1]) / m.in_den[1] - 1.0;
    in.x[2] = 2.0 * (ks - m.in_lo[2]) / m.in_den[2] - 1.0;
  }
}
// transpose of the above: the clamp is a constant selector (a clamped cell passes nothing back, like every other clamp of
// the path); d|U|/d(u, v) at |U| = 0 is taken as 0 (the reference's sqrt would give NaN there)
HG_HD void inputs_adj(const Model& m, const Inputs& in, const double* xbar, double& xib, double& qxb, double& qyb) {
  xib = qxb = qyb = 0.0;
  if (in.dry) return;
  double hb = xbar[0] * (2.0 / m.in_den[0]);
  if (m.n_in == 3 && in.umag > 0.0) {
    const double Ub = xbar[1] * (2.0 / m.in_den[1]);
    const double ub = Ub * in.u / in.umag, vb = Ub * in.v / in.umag;
    qxb = ub / in.h;
    qyb = vb / in.h;
    hb -= (ub * in.u + vb * in.v) / in.h;
  }
  xib = hb;
}

HG_HD double act_fwd(int a, double z) {
  switch (a) {
    case HG_ACT_RELU: return z > 0.0 ? z : 0.0;
    case HG_ACT_LEAKYRELU: return z > 0.0 ? z : 0.01 * z;
    case HG_ACT_SIGMOID: return 1.0 / (1.0 + exp(-z));
    case HG_ACT_TANH: return tanh(z);
    case HG_ACT_SOFTPLUS: return log1p(exp(-fabs(z))) + (z > 0.0 ? z : 0.0);
    default: return z;
  }
}
// derivative from the pre-activation z and the value y = act(z)
HG_HD double act_der(int a, double z, double y) {
  switch (a) {
    case HG_ACT_RELU: return y > 0.0 ? 1.0 : 0.0;
    case HG_ACT_LEAKYRELU: return z > 0.0 ? 1.0 : 0.01;
    case HG_ACT_SIGMOID: return y * (1.0 - y);
    case HG_ACT_TANH: return 1.0 - y * y;
    case HG_ACT_SOFTPLUS: return 1.0 / (1.0 + exp(-z));
    default: return 1.0;
  }
}

// what the reverse sweep needs from the forward pass of one cell
struct Tape {
  double z[MAXH][MAXW];      // pre-activations
  double y[MAXH][MAXW];      // activations (LayerNorm input)
  double xh[MAXH][MAXW];     // normalised values (y - mean) * rstd (= y without LayerNorm)
  double a[MAXH][MAXW];      // layer outputs xh * scale + bias
  double rstd[MAXH];
  double s;                  // sigmoid of the output unit
};

// Forward pass of one cell.  stats[2 l], stats[2 l + 1] = (mean, 1 / sqrt(var + eps)) of hidden layer l over the whole
// array (HG_LN_WHOLE_ARRAY only).  upto >= 0: stop after the activations of hidden layer `upto` (t.y[upto] is valid;
// used to accumulate that layer's statistics) and return 0.  Otherwise returns n.
HG_HD double forward(const Model& m, const double* th, const double* x, const double* stats, int upto, Tape& t) {
  const double* prev = x;
  int n_prev = m.n_in;
  for (int l = 0; l < m.n_hidden; ++l) {
    const int W = m.width[l];
    const double* w = th + m.off_w[l];
    const double* b = th + m.off_b[l];
    for (int j = 0; j < W; ++j) {
      double z = b[j];
      for (int i = 0; i < n_prev; ++i) z += w[j + i * W] * prev[i];
      t.z[l][j] = z;
      t.y[l][j] = act_fwd(m.act[l], z);
    }
    if (l == upto) return 0.0;
    if (m.ln_mode == HG_LN_NONE) {
      t.rstd[l] = 1.0;
      for (int j = 0; j < W; ++j) t.xh[l][j] = t.a[l][j] = t.y[l][j];
    } else {
      double mu, rstd;
      if (m.ln_mode == HG_LN_PER_CELL) {
        mu = 0.0;
        for (int j = 0; j < W; ++j) mu += t.y[l][j];
        mu /= W;
        double var = 0.0;
        for (int j = 0; j < W; ++j) var += (t.y[l][j] - mu) * (t.y[l][j] - mu);
        rstd = 1.0 / sqrt(var / W + m.eps);
      } else {
        mu = stats[2 * l];
        rstd = stats[2 * l + 1];
      }
      t.rstd[l] = rstd;
      const double* g = th + m.off_g[l];
      const double* be = th + m.off_be[l];
      for (int j = 0; j < W; ++j) {
        t.xh[l][j] = (t.y[l][j] - mu) * rstd;
        t.a[l][j] = t.xh[l][j] * g[j] + be[j];
      }
    }
    prev = t.a[l];
    n_prev = W;
  }
  const double* w = th + m.off_w[m.n_hidden];
  double zo = th[m.off_b[m.n_hidden]];
  for (int i = 0; i < n_prev; ++i) zo += w[i] * prev[i];
  t.s = 1.0 / (1.0 + exp(-zo));
  return m.out_lo + m.out_span * t.s;
}

// Reverse sweep of one cell from nbar = d(.)/dn.  bstats[2 l], bstats[2 l + 1] = whole-array means of g and g * xh of hidden
// layer l (g = abar * scale; HG_LN_WHOLE_ARRAY only).  stop_at >= 0: stop once g of hidden layer `stop_at` is known and
// return it in g_out[width] (used to accumulate that layer's bstats); nothing else is written.  Otherwise the input
// adjoints go to xbar[n_in] and, when acc != NULL, the cell's contribution is ADDED to acc[n_params] (= thetabar).
HG_HD void backward(const Model& m, const double* th, const double* x, const Tape& t, const double* bstats, double nbar,
                    int stop_at, double* g_out, double* acc, double* xbar) {
  double abar[MAXW], zbar[MAXW];
  const int L = m.n_hidden;
  {
    const double zob = nbar * m.out_span * t.s * (1.0 - t.s);
    const double* w = th + m.off_w[L];
    const int W = m.width[L - 1];
    for (int i = 0; i < W; ++i) abar[i] = zob * w[i];
    if (acc) {
      for (int i = 0; i < W; ++i) acc[m.off_w[L] + i] += zob * t.a[L - 1][i];
      acc[m.off_b[L]] += zob;
    }
  }
  for (int l = L - 1; l >= 0; --l) {
    const int W = m.width[l];
    // LayerNorm
    if (m.ln_mode == HG_LN_NONE) {
      for (int j = 0; j < W; ++j) zbar[j] = abar[j];
    } else {
      const double* g = th + m.off_g[l];
      double gj[MAXW];
      for (int j = 0; j < W; ++j) gj[j] = abar[j] * g[j];
      if (l == stop_at) {
        for (int j = 0; j < W; ++j) g_out[j] = gj[j];
        return;
      }
      if (acc) {
        for (int j = 0; j < W; ++j) {
          acc[m.off_g[l] + j] += abar[j] * t.xh[l][j];
          acc[m.off_be[l] + j] += abar[j];
        }
      }
      double m1, m2;
      if (m.ln_mode == HG_LN_PER_CELL) {
        m1 = m2 = 0.0;
        for (int j = 0; j < W; ++j) { m1 += gj[j]; m2 += gj[j] * t.xh[l][j]; }
        m1 /= W;
        m2 /= W;
      } else {
        m1 = bstats[2 * l];
        m2 = bstats[2 * l + 1];
      }
      for (int j = 0; j < W; ++j) zbar[j] = (gj[j] - m1 - t.xh[l][j] * m2) * t.rstd[l];
    }
    // activation
    for (int j = 0; j < W; ++j) zbar[j] *= act_der(m.act[l], t.z[l][j], t.y[l][j]);
    // Dense
    const int n_prev = l == 0 ? m.n_in : m.width[l - 1];
    const double* prev = l == 0 ? x : t.a[l - 1];
    const double* w = th + m.off_w[l];
    if (acc) {
      for (int j = 0; j < W; ++j) {
        for (int i = 0; i < n_prev; ++i) acc[m.off_w[l] + j + i * W] += zbar[j] * prev[i];
        acc[m.off_b[l] + j] += zbar[j];
      }
    }
    double pb[MAXW];
    for (int i = 0; i < n_prev; ++i) {
      double sacc = 0.0;
      for (int j = 0; j < W; ++j) sacc += w[j + i * W] * zbar[j];
      pb[i] = sacc;
    }
    if (l == 0) {
      for (int i = 0; i < n_prev; ++i) xbar[i] = pb[i];
    } else {
      for (int i = 0; i < n_prev; ++i) abar[i] = pb[i];
    }
  }
}

// running (count, mean, M2) and the pairwise combination of two of them (Chan et al.): accurate one-pass variance with a
// fixed combination order
struct Moments {
  double n, mean, m2;
};
HG_HD void moments_push(Moments& a, double x) {
  a.n += 1.0;
  const double d = x - a.mean;
  a.mean += d / a.n;
  a.m2 += d * (x - a.mean);
}
HG_HD Moments moments_merge(const Moments& a, const Moments& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  const double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

// hg_ude_desc -> Model with every range checked; returns NULL or the reason the descriptor is rejected
inline const char* make_model(const hg_ude_desc* d, Model& m) {
  if (d->choice != HG_UDE_MANNING_H && d->choice != HG_UDE_MANNING_H_UMAG_KS) return "unknown UDE choice (FlowResistance is not built)";
  if (d->n_hidden < 1 || d->n_hidden > MAXH) return "n_hidden out of range";
  if (d->layernorm < HG_LN_NONE || d->layernorm > HG_LN_WHOLE_ARRAY) return "unknown layernorm mode";
  if (d->n_params < 1 || d->n_params > MAXP) return "n_params out of range";
  if (d->layernorm != HG_LN_NONE && !(d->ln_epsilon >= 0.0)) return "ln_epsilon must be non-negative";
  m = Model();
  m.n_in = d->choice == HG_UDE_MANNING_H ? 1 : 3;
  m.n_hidden = d->n_hidden;
  m.ln_mode = d->layernorm;
  m.n_params = (int32_t)d->n_params;
  m.eps = d->ln_epsilon;
  const double* lo_hi[3] = {d->h_bounds, d->umag_bounds, d->ks_bounds};
  for (int i = 0; i < m.n_in; ++i) {
    m.in_lo[i] = lo_hi[i][0];
    m.in_den[i] = lo_hi[i][1] - lo_hi[i][0];
    if (!(m.in_den[i] != 0.0) || m.in_den[i] != m.in_den[i]) return "input bounds must differ";
  }
  m.out_lo = d->output_bounds[0];
  m.out_span = d->output_bounds[1] - d->output_bounds[0];
  int n_prev = m.n_in;
  auto fits = [&](int64_t off, int64_t len) { return off >= 0 && off + len <= d->n_params; };
  for (int l = 0; l <= m.n_hidden; ++l) {
    const int W = l < m.n_hidden ? d->width[l] : 1;
    if (l < m.n_hidden) {
      if (W < 1 || W > MAXW) return "hidden width out of range";
      if (d->activation[l] < HG_ACT_IDENTITY || d->activation[l] > HG_ACT_SOFTPLUS) return "unknown activation";
      m.width[l] = W;
      m.act[l] = d->activation[l];
      if (m.ln_mode != HG_LN_NONE) {
        if (!fits(d->off_ln_scale[l], W) || !fits(d->off_ln_bias[l], W)) return "LayerNorm offsets outside theta";
        m.off_g[l] = (int32_t)d->off_ln_scale[l];
        m.off_be[l] = (int32_t)d->off_ln_bias[l];
      }
    }
    if (!fits(d->off_weight[l], (int64_t)W * n_prev) || !fits(d->off_bias[l], W)) return "Dense offsets outside theta";
    m.off_w[l] = (int32_t)d->off_weight[l];
    m.off_b[l] = (int32_t)d->off_bias[l];
    n_prev = W;
  }
  return nullptr;
}

}  // namespace ude
}  // namespace hg
