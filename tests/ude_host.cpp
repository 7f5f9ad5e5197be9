// CPU check of the UDE per-cell arithmetic (hydrograd.jl_b200/csrc/hg_ude.h): the very functions the kernels of hg_ude.cu call,
// compiled by g++ and driven by plain loops in the kernels' pass structure (statistics passes layer by layer, then n; reverse:
// bstats passes from the last hidden layer down, then the full sweep).  TEST INFRASTRUCTURE: built and loaded only by
// tests/test_ude_cpu.py; the product has no CPU path.
#include <vector>

#include "../hydrograd.jl_b200/csrc/hg_ude.h"

using namespace hg::ude;

namespace {
struct Pass {
  Model m;
  int64_t N;
  const double *Q, *hstill, *ks, *th;
  double hs;
  double stats[2 * MAXH], bstats[2 * MAXH];
  void cell_inputs(int64_t i, Inputs& in) const { inputs(m, Q[i], Q[N + i], Q[2 * N + i], hstill[i], ks ? ks[i] : 1.0, hs, in); }
  void forward_stats() {
    if (m.ln_mode != HG_LN_WHOLE_ARRAY) return;
    for (int l = 0; l < m.n_hidden; ++l) {
      Moments acc{0, 0, 0};
      for (int64_t i = 0; i < N; ++i) {
        Inputs in; Tape t;
        cell_inputs(i, in);
        forward(m, th, in.x, stats, l, t);
        for (int j = 0; j < m.width[l]; ++j) moments_push(acc, t.y[l][j]);
      }
      stats[2 * l] = acc.mean;
      stats[2 * l + 1] = 1.0 / sqrt(acc.m2 / acc.n + m.eps);
    }
  }
};
}  // namespace

extern "C" {
int ude_host_max_params() { return MAXP; }

int ude_host_manning(const hg_ude_desc* d, int64_t N, const double* Q, const double* hstill, const double* ks, double hs,
                     const double* th, double* n_out) {
  Pass p{};
  if (make_model(d, p.m)) return 1;
  p.N = N; p.Q = Q; p.hstill = hstill; p.ks = ks; p.th = th; p.hs = hs;
  p.forward_stats();
  for (int64_t i = 0; i < N; ++i) {
    Inputs in; Tape t;
    p.cell_inputs(i, in);
    n_out[i] = forward(p.m, th, in.x, p.stats, -1, t);
  }
  return 0;
}

// Qbar_add[3N] = (dn/dQ)^T nbar, thbar[n_params] = (dn/dtheta)^T nbar
int ude_host_pullback(const hg_ude_desc* d, int64_t N, const double* Q, const double* hstill, const double* ks, double hs,
                      const double* th, const double* nbar, double* Qbar_add, double* thbar) {
  Pass p{};
  if (make_model(d, p.m)) return 1;
  const Model& m = p.m;
  p.N = N; p.Q = Q; p.hstill = hstill; p.ks = ks; p.th = th; p.hs = hs;
  p.forward_stats();
  if (m.ln_mode == HG_LN_WHOLE_ARRAY) {
    for (int l = m.n_hidden - 1; l >= 0; --l) {
      double s1 = 0.0, s2 = 0.0;
      for (int64_t i = 0; i < N; ++i) {
        Inputs in; Tape t; double g[MAXW];
        p.cell_inputs(i, in);
        forward(m, th, in.x, p.stats, -1, t);
        backward(m, th, in.x, t, p.bstats, nbar[i], l, g, nullptr, nullptr);
        for (int j = 0; j < m.width[l]; ++j) { s1 += g[j]; s2 += g[j] * t.xh[l][j]; }
      }
      const double M = (double)N * m.width[l];
      p.bstats[2 * l] = s1 / M;
      p.bstats[2 * l + 1] = s2 / M;
    }
  }
  std::vector<double> acc(m.n_params, 0.0);
  for (int64_t i = 0; i < N; ++i) {
    Inputs in; Tape t; double xbar[3] = {0, 0, 0};
    p.cell_inputs(i, in);
    forward(m, th, in.x, p.stats, -1, t);
    backward(m, th, in.x, t, p.bstats, nbar[i], -1, nullptr, acc.data(), xbar);
    inputs_adj(m, in, xbar, Qbar_add[i], Qbar_add[N + i], Qbar_add[2 * N + i]);
  }
  for (int k = 0; k < m.n_params; ++k) thbar[k] = acc[k];
  return 0;
}
}
