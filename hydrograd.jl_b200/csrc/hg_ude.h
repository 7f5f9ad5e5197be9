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
#define HG_HDC __host__ __device__ constexpr
#define HG_UNROLL _Pragma("unroll")
#else
#define HG_HD inline
#define HG_HDC constexpr
#define HG_UNROLL
#endif

namespace hg {
namespace ude {

// Reciprocal, reciprocal square root and tanh of the per-cell arithmetic.  On the device: the branch-free MUFU-seed + one
// third-order step of the RHS kernels (hg_device.cuh; arguments here are positive normals: depths, 1 + exp(.), variances +
// eps) and tanh from ONE expm1 and one reciprocal -- the closure is bound by the fp64 pipe, and the library's `/`, `sqrt` and
// `tanh` each cost a branchy slow path.  On the host (the g++ check of tests/ude_host.cpp): the plain expressions.
#if defined(__CUDA_ARCH__)
HG_HD double u_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
HG_HD double u_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-(x * r), r, 1.0);
  return fma(r * e, fma(0.375, e, 0.5), r);
}
HG_HD double u_tanh(double z) {
  const double e = expm1(-2.0 * fabs(z));          // in (-1, 0]: tanh|z| = -e / (2 + e), no cancellation near 0
  const double t = -e * u_rcp(2.0 + e);
  return z < 0.0 ? -t : t;
}
#else
HG_HD double u_rcp(double x) { return 1.0 / x; }
HG_HD double u_rsqrt(double x) { return 1.0 / sqrt(x); }
HG_HD double u_tanh(double z) { return tanh(z); }
#endif

constexpr int MAXH = HG_UDE_MAX_HIDDEN, MAXW = HG_UDE_MAX_WIDTH;
constexpr int MAXP = 3 * MAXW + 3 * MAXW + (MAXH - 1) * (MAXW * MAXW + 3 * MAXW) + MAXW + 1;   // 233 with 3 x 8

// The descriptor in kernel-argument form (by value, ~250 bytes).  The off_* fields index the CANONICAL parameter vector --
// per hidden layer: weight, bias, [LayerNorm scale, LayerNorm bias]; then the output weight and bias; n_params entries --
// which depends on the shape only (canonical_offsets), so that a kernel specialised on the shape addresses it with
// compile-time constants.  ThetaMap translates to the caller's theta (hg_ude_desc.off_*).
struct Model {
  int32_t n_in = 0, n_hidden = 0, ln_mode = 0, n_params = 0;
  int32_t width[MAXH] = {}, act[MAXH] = {};
  int32_t off_w[MAXH + 1] = {}, off_b[MAXH + 1] = {}, off_g[MAXH] = {}, off_be[MAXH] = {};
  double eps = 0.0, in_lo[3] = {}, in_den[3] = {}, out_lo = 0.0, out_span = 0.0;
};

struct ThetaMap {
  int16_t to_user[MAXP];   // canonical index -> index in the caller's theta
};
HG_HD void canonical_offsets(Model& m) {
  int off = 0, n_prev = m.n_in;
  HG_UNROLL
  for (int l = 0; l < MAXH; ++l) {
    if (l < m.n_hidden) {
      const int W = m.width[l];
      m.off_w[l] = off; off += W * n_prev;
      m.off_b[l] = off; off += W;
      if (m.ln_mode != HG_LN_NONE) {
        m.off_g[l] = off; off += W;
        m.off_be[l] = off; off += W;
      }
      n_prev = W;
    }
  }
  m.off_w[m.n_hidden] = off; off += n_prev;
  m.off_b[m.n_hidden] = off; off += 1;
  m.n_params = off;
}

// ---- network shape: either read from the Model at run time (Generic) or fixed at compile time (Spec).  Every loop of the
// per-cell functions below runs to the compile-time maxima (MAXH, MAXW) under `if (index < bound)` guards, so that after
// unrolling every array index is a constant and the tape lives in registers; with a Spec the guards fold away as well.
struct Generic {
  static constexpr bool kSpec = false;
  static constexpr int P = MAXP;
};
template <int NIN_, int NH_, int W_, int ACT_, int LN_>
struct Spec {
  static constexpr bool kSpec = true;
  static constexpr int NIN = NIN_, NH = NH_, W = W_, ACT = ACT_, LN = LN_;
  static constexpr int kLnP = LN_ == HG_LN_NONE ? 0 : 2 * W_;
  static constexpr int kLayer0 = NIN_ * W_ + W_ + kLnP, kLayer = W_ * W_ + W_ + kLnP;
  static constexpr int P = kLayer0 + (NH_ - 1) * kLayer + W_ + 1;
  // canonical offsets (canonical_offsets() with these widths)
  static HG_HDC int base(int l) { return l == 0 ? 0 : kLayer0 + (l - 1) * kLayer; }
  static HG_HDC int n_prev(int l) { return l == 0 ? NIN_ : W_; }
  static HG_HDC int off_w(int l) { return base(l); }
  static HG_HDC int off_b(int l) { return l < NH_ ? base(l) + n_prev(l) * W_ : base(l) + W_; }
  static HG_HDC int off_g(int l) { return off_b(l) + W_; }
  static HG_HDC int off_be(int l) { return off_b(l) + 2 * W_; }
};
template <class S> HG_HD int s_nin(const Model& m) { if constexpr (S::kSpec) return S::NIN; else return m.n_in; }
template <class S> HG_HD int s_nh(const Model& m) { if constexpr (S::kSpec) return S::NH; else return m.n_hidden; }
template <class S> HG_HD int s_ln(const Model& m) { if constexpr (S::kSpec) return S::LN; else return m.ln_mode; }
template <class S> HG_HD int s_np(const Model& m) { if constexpr (S::kSpec) return S::P; else return m.n_params; }
template <class S> HG_HD int s_width(const Model& m, int l) { if constexpr (S::kSpec) return S::W; else return m.width[l]; }
template <class S> HG_HD int s_act(const Model& m, int l) { if constexpr (S::kSpec) return S::ACT; else return m.act[l]; }
template <class S> HG_HD int s_off_w(const Model& m, int l) { if constexpr (S::kSpec) return S::off_w(l); else return m.off_w[l]; }
template <class S> HG_HD int s_off_b(const Model& m, int l) { if constexpr (S::kSpec) return S::off_b(l); else return m.off_b[l]; }
template <class S> HG_HD int s_off_g(const Model& m, int l) { if constexpr (S::kSpec) return S::off_g(l); else return m.off_g[l]; }
template <class S> HG_HD int s_off_be(const Model& m, int l) { if constexpr (S::kSpec) return S::off_be(l); else return m.off_be[l]; }

// the shapes with a compile-time instantiation: what the reference ships (examples/SWE_2D/UDE/*/run_control.json: 1 or 3
// inputs, hidden_layers [3, 3], tanh) with either LayerNorm reading; everything else runs Generic
#define HG_UDE_UNPAREN(...) __VA_ARGS__
#define HG_UDE_SPECS(X)                                                   \
  X(0, (::hg::ude::Generic))                                              \
  X(1, (::hg::ude::Spec<1, 2, 3, HG_ACT_TANH, HG_LN_WHOLE_ARRAY>))        \
  X(2, (::hg::ude::Spec<3, 2, 3, HG_ACT_TANH, HG_LN_WHOLE_ARRAY>))        \
  X(3, (::hg::ude::Spec<1, 2, 3, HG_ACT_TANH, HG_LN_PER_CELL>))           \
  X(4, (::hg::ude::Spec<3, 2, 3, HG_ACT_TANH, HG_LN_PER_CELL>))
inline int spec_of(const Model& m) {
  if (m.n_hidden == 2 && m.width[0] == 3 && m.width[1] == 3 && m.act[0] == HG_ACT_TANH && m.act[1] == HG_ACT_TANH) {
    if (m.ln_mode == HG_LN_WHOLE_ARRAY) return m.n_in == 1 ? 1 : 2;
    if (m.ln_mode == HG_LN_PER_CELL) return m.n_in == 1 ? 3 : 4;
  }
  return 0;
}

// network inputs of one cell from its raw state: clamp of semi_discretize_swe_2D.jl:101-106, u = q/h, |U| = sqrt(u^2 + v^2)
// (:166-169), normalisation to [-1, 1] (process_ManningN_2D.jl:233-240)
struct Inputs {
  double x[3], h, u, v, umag;
  bool dry;
};
template <class S>
HG_HD void inputs(const Model& m, double xi, double qx, double qy, double hst, double ks, double hs, Inputs& in) {
  const double h0 = xi + hst;
  in.dry = h0 <= hs;
  in.h = in.dry ? hs : h0;
  const double rh = u_rcp(in.h);
  in.u = in.dry ? 0.0 : qx * rh;
  in.v = in.dry ? 0.0 : qy * rh;
  in.umag = sqrt(in.u * in.u + in.v * in.v);
  in.x[0] = 2.0 * (in.h - m.in_lo[0]) * u_rcp(m.in_den[0]) - 1.0;
  in.x[1] = in.x[2] = 0.0;
  if (s_nin<S>(m) == 3) {
    in.x[1] = 2.0 * (in.umag - m.in_lo[1]) * u_rcp(m.in_den[1]) - 1.0;
    in.x[2] = 2.0 * (ks - m.in_lo[2]) * u_rcp(m.in_den[2]) - 1.0;
  }
}
// transpose of the above: the clamp is a constant selector (a clamped cell passes nothing back, like every other clamp of
// the path); d|U|/d(u, v) at |U| = 0 is taken as 0 (the reference's sqrt would give NaN there)
template <class S>
HG_HD void inputs_adj(const Model& m, const Inputs& in, const double* xbar, double& xib, double& qxb, double& qyb) {
  xib = qxb = qyb = 0.0;
  if (in.dry) return;
  double hb = xbar[0] * (2.0 * u_rcp(m.in_den[0]));
  if (s_nin<S>(m) == 3 && in.umag > 0.0) {
    const double Ub = xbar[1] * (2.0 * u_rcp(m.in_den[1]));
    const double rU = u_rcp(in.umag), rh = u_rcp(in.h);
    const double ub = Ub * in.u * rU, vb = Ub * in.v * rU;
    qxb = ub * rh;
    qyb = vb * rh;
    hb -= (ub * in.u + vb * in.v) * rh;
  }
  xib = hb;
}

HG_HD double act_fwd(int a, double z) {
  switch (a) {
    case HG_ACT_RELU: return z > 0.0 ? z : 0.0;
    case HG_ACT_LEAKYRELU: return z > 0.0 ? z : 0.01 * z;
    case HG_ACT_SIGMOID: return u_rcp(1.0 + exp(-z));
    case HG_ACT_TANH: return u_tanh(z);
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
    case HG_ACT_SOFTPLUS: return u_rcp(1.0 + exp(-z));
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

// output of layer l - 1 (the network input for l = 0), entry i
HG_HD double layer_in(const double* x, const Tape& t, int l, int i) { return l == 0 ? x[i < 3 ? i : 0] : t.a[l > 0 ? l - 1 : 0][i]; }

// Forward pass of one cell.  stats[2 l], stats[2 l + 1] = (mean, 1 / sqrt(var + eps)) of hidden layer l over the whole
// array (HG_LN_WHOLE_ARRAY only).  upto >= 0: stop after the activations of hidden layer `upto` (t.y[upto] is valid;
// used to accumulate that layer's statistics) and return 0.  Otherwise returns n.
template <class S>
HG_HD double forward(const Model& m, const double* th, const double* x, const double* stats, int upto, Tape& t) {
  const int nh = s_nh<S>(m), ln = s_ln<S>(m);
  HG_UNROLL
  for (int l = 0; l < MAXH; ++l) {
    if (l < nh) {
      const int W = s_width<S>(m, l), n_prev = l == 0 ? s_nin<S>(m) : s_width<S>(m, l > 0 ? l - 1 : 0);
      const double* w = th + s_off_w<S>(m, l);
      const double* b = th + s_off_b<S>(m, l);
      HG_UNROLL
      for (int j = 0; j < MAXW; ++j) {
        if (j < W) {
          double z = b[j];
          HG_UNROLL
          for (int i = 0; i < MAXW; ++i)
            if (i < n_prev) z += w[j + i * W] * layer_in(x, t, l, i);
          t.z[l][j] = z;
          t.y[l][j] = act_fwd(s_act<S>(m, l), z);
        }
      }
      if (l == upto) return 0.0;
      if (ln == HG_LN_NONE) {
        t.rstd[l] = 1.0;
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j)
          if (j < W) t.xh[l][j] = t.a[l][j] = t.y[l][j];
      } else {
        double mu, rstd;
        if (ln == HG_LN_PER_CELL) {
          mu = 0.0;
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j)
            if (j < W) mu += t.y[l][j];
          const double rW = 1.0 / W;      // (compile-time for the specialised shapes)
          mu *= rW;
          double var = 0.0;
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j)
            if (j < W) var += (t.y[l][j] - mu) * (t.y[l][j] - mu);
          rstd = u_rsqrt(var * rW + m.eps);
        } else {
          mu = stats[2 * l];
          rstd = stats[2 * l + 1];
        }
        t.rstd[l] = rstd;
        const double* g = th + s_off_g<S>(m, l);
        const double* be = th + s_off_be<S>(m, l);
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j) {
          if (j < W) {
            t.xh[l][j] = (t.y[l][j] - mu) * rstd;
            t.a[l][j] = t.xh[l][j] * g[j] + be[j];
          }
        }
      }
    }
  }
  double zo = 0.0;
  HG_UNROLL
  for (int l = 0; l < MAXH; ++l) {          // the output unit reads the last hidden layer (l = nh - 1)
    if (l == nh - 1) {
      const int W = s_width<S>(m, l);
      const double* w = th + s_off_w<S>(m, l + 1);
      zo = th[s_off_b<S>(m, l + 1)];
      HG_UNROLL
      for (int i = 0; i < MAXW; ++i)
        if (i < W) zo += w[i] * t.a[l][i];
    }
  }
  t.s = u_rcp(1.0 + exp(-zo));
  return m.out_lo + m.out_span * t.s;
}

// Reverse sweep of one cell from nbar = d(.)/dn.  bstats[2 l], bstats[2 l + 1] = whole-array means of g and g * xh of hidden
// layer l (g = abar * scale; HG_LN_WHOLE_ARRAY only).  stop_at >= 0: stop once g of hidden layer `stop_at` is known and
// return it in g_out[MAXW] (used to accumulate that layer's bstats); nothing else is written.  Otherwise the input
// adjoints go to xbar[3] and, when acc != NULL, the cell's contribution is ADDED to acc[n_params] (= thetabar, canonical order).
template <class S>
HG_HD void backward(const Model& m, const double* th, const double* x, const Tape& t, const double* bstats, double nbar,
                    int stop_at, double* g_out, double* acc, double* xbar) {
  double abar[MAXW], zbar[MAXW];
  const int nh = s_nh<S>(m), ln = s_ln<S>(m);
  HG_UNROLL
  for (int j = 0; j < MAXW; ++j) abar[j] = zbar[j] = 0.0;
  const double zob = nbar * m.out_span * t.s * (1.0 - t.s);
  HG_UNROLL
  for (int l = 0; l < MAXH; ++l) {
    if (l == nh - 1) {
      const int W = s_width<S>(m, l);
      const int ow = s_off_w<S>(m, l + 1), ob = s_off_b<S>(m, l + 1);
      HG_UNROLL
      for (int i = 0; i < MAXW; ++i) {
        if (i < W) {
          abar[i] = zob * th[ow + i];
          if (acc) acc[ow + i] += zob * t.a[l][i];
        }
      }
      if (acc) acc[ob] += zob;
    }
  }
  HG_UNROLL
  for (int l = MAXH - 1; l >= 0; --l) {
    if (l < nh) {
      const int W = s_width<S>(m, l);
      // LayerNorm
      if (ln == HG_LN_NONE) {
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j)
          if (j < W) zbar[j] = abar[j];
      } else {
        const int og = s_off_g<S>(m, l), obe = s_off_be<S>(m, l);
        double gj[MAXW];
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j) gj[j] = j < W ? abar[j] * th[og + j] : 0.0;
        if (l == stop_at) {
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j) g_out[j] = gj[j];
          return;
        }
        if (acc) {
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j) {
            if (j < W) {
              acc[og + j] += abar[j] * t.xh[l][j];
              acc[obe + j] += abar[j];
            }
          }
        }
        double m1, m2;
        if (ln == HG_LN_PER_CELL) {
          m1 = m2 = 0.0;
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j)
            if (j < W) { m1 += gj[j]; m2 += gj[j] * t.xh[l][j]; }
          m1 /= W;
          m2 /= W;
        } else {
          m1 = bstats[2 * l];
          m2 = bstats[2 * l + 1];
        }
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j)
          if (j < W) zbar[j] = (gj[j] - m1 - t.xh[l][j] * m2) * t.rstd[l];
      }
      // activation
      HG_UNROLL
      for (int j = 0; j < MAXW; ++j)
        if (j < W) zbar[j] *= act_der(s_act<S>(m, l), t.z[l][j], t.y[l][j]);
      // Dense
      const int n_prev = l == 0 ? s_nin<S>(m) : s_width<S>(m, l > 0 ? l - 1 : 0);
      const int ow = s_off_w<S>(m, l), ob = s_off_b<S>(m, l);
      if (acc) {
        HG_UNROLL
        for (int j = 0; j < MAXW; ++j) {
          if (j < W) {
            HG_UNROLL
            for (int i = 0; i < MAXW; ++i)
              if (i < n_prev) acc[ow + j + i * W] += zbar[j] * layer_in(x, t, l, i);
            acc[ob + j] += zbar[j];
          }
        }
      }
      double pb[MAXW];
      HG_UNROLL
      for (int i = 0; i < MAXW; ++i) {
        double sacc = 0.0;
        if (i < n_prev) {
          HG_UNROLL
          for (int j = 0; j < MAXW; ++j)
            if (j < W) sacc += th[ow + j + i * W] * zbar[j];
        }
        pb[i] = sacc;
      }
      if (l == 0) {
        xbar[0] = pb[0];
        xbar[1] = pb[1];
        xbar[2] = pb[2];
      } else {
        HG_UNROLL
        for (int i = 0; i < MAXW; ++i) abar[i] = pb[i];
      }
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

// hg_ude_desc -> Model (canonical offsets) + ThetaMap, every range checked; returns NULL or the reason the descriptor is rejected
inline const char* make_model(const hg_ude_desc* d, Model& m, ThetaMap& map) {
  if (d->choice != HG_UDE_MANNING_H && d->choice != HG_UDE_MANNING_H_UMAG_KS) return "unknown UDE choice (FlowResistance is not built)";
  if (d->n_hidden < 1 || d->n_hidden > MAXH) return "n_hidden out of range";
  if (d->layernorm < HG_LN_NONE || d->layernorm > HG_LN_WHOLE_ARRAY) return "unknown layernorm mode";
  if (d->n_params < 1 || d->n_params > 32767) return "n_params out of range";
  if (d->layernorm != HG_LN_NONE && !(d->ln_epsilon >= 0.0)) return "ln_epsilon must be non-negative";
  m = Model();
  m.n_in = d->choice == HG_UDE_MANNING_H ? 1 : 3;
  m.n_hidden = d->n_hidden;
  m.ln_mode = d->layernorm;
  m.eps = d->ln_epsilon;
  const double* lo_hi[3] = {d->h_bounds, d->umag_bounds, d->ks_bounds};
  for (int i = 0; i < m.n_in; ++i) {
    m.in_lo[i] = lo_hi[i][0];
    m.in_den[i] = lo_hi[i][1] - lo_hi[i][0];
    if (!(m.in_den[i] != 0.0) || m.in_den[i] != m.in_den[i]) return "input bounds must differ";
  }
  m.out_lo = d->output_bounds[0];
  m.out_span = d->output_bounds[1] - d->output_bounds[0];
  for (int l = 0; l < m.n_hidden; ++l) {
    if (d->width[l] < 1 || d->width[l] > MAXW) return "hidden width out of range";
    if (d->activation[l] < HG_ACT_IDENTITY || d->activation[l] > HG_ACT_SOFTPLUS) return "unknown activation";
    m.width[l] = d->width[l];
    m.act[l] = d->activation[l];
  }
  canonical_offsets(m);
  if (m.n_params > d->n_params) return "n_params is smaller than the network's parameter count";
  for (int k = 0; k < MAXP; ++k) map.to_user[k] = 0;
  bool ok = true;
  auto place = [&](int canon, int64_t user, int len) {
    if (user < 0 || user + len > d->n_params) { ok = false; return; }
    for (int k = 0; k < len; ++k) map.to_user[canon + k] = (int16_t)(user + k);
  };
  int n_prev = m.n_in;
  for (int l = 0; l <= m.n_hidden; ++l) {
    const int W = l < m.n_hidden ? m.width[l] : 1;
    place(m.off_w[l], d->off_weight[l], W * n_prev);
    place(m.off_b[l], d->off_bias[l], W);
    if (l < m.n_hidden && m.ln_mode != HG_LN_NONE) {
      place(m.off_g[l], d->off_ln_scale[l], W);
      place(m.off_be[l], d->off_ln_bias[l], W);
    }
    n_prev = W;
  }
  if (!ok) return "parameter offsets outside theta";
  // two network arrays must not share an entry of theta (its adjoint would be ambiguous)
  for (int a = 0; a < m.n_params; ++a)
    for (int b = a + 1; b < m.n_params; ++b)
      if (map.to_user[a] == map.to_user[b]) return "parameter offsets overlap";
  return nullptr;
}

}  // namespace ude
}  // namespace hg
