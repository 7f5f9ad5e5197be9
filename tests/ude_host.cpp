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
  ThetaMap map;
  std::vector<double> thc;   // theta in the canonical order the kernels stage it in
  int64_t N;
  const double *Q, *hstill, *ks, *th;
  double hs;
  bool setup(const hg_ude_desc* d, const double* theta_user) {
    if (make_model(d, m, map)) return false;
    thc.resize(m.n_params);
    for (int k = 0; k < m.n_params; ++k) thc[k] = theta_user[map.to_user[k]];
    th = thc.data();
    return true;
  }
  double stats[2 * MAXH], bstats[2 * MAXH];
  template <class S>
  void cell_inputs(int64_t i, Inputs& in) const { inputs<S>(m, Q[i], Q[N + i], Q[2 * N + i], hstill[i], ks ? ks[i] : 1.0, hs, in); }
  template <class S>
  void forward_stats() {
    if (m.ln_mode != HG_LN_WHOLE_ARRAY) return;
    for (int l = 0; l < m.n_hidden; ++l) {
      Moments acc{0, 0, 0};
      for (int64_t i = 0; i < N; ++i) {
        Inputs in; Tape t;
        cell_inputs<S>(i, in);
        forward<S>(m, th, in.x, stats, l, t);
        for (int j = 0; j < m.width[l]; ++j) moments_push(acc, t.y[l][j]);
      }
      stats[2 * l] = acc.mean;
      stats[2 * l + 1] = 1.0 / sqrt(acc.m2 / acc.n + m.eps);
    }
  }
};

template <class S>
void manning_impl(Pass& p, double* n_out) {
  p.template forward_stats<S>();
  for (int64_t i = 0; i < p.N; ++i) {
    Inputs in; Tape t;
    p.template cell_inputs<S>(i, in);
    n_out[i] = forward<S>(p.m, p.th, in.x, p.stats, -1, t);
  }
}

template <class S>
void pullback_impl(Pass& p, const double* nbar, double* Qbar_add, std::vector<double>& acc) {
  const Model& m = p.m;
  const double* th = p.th;
  const int64_t N = p.N;
  p.template forward_stats<S>();
  if (m.ln_mode == HG_LN_WHOLE_ARRAY) {
    for (int l = m.n_hidden - 1; l >= 0; --l) {
      double s1 = 0.0, s2 = 0.0;
      for (int64_t i = 0; i < N; ++i) {
        Inputs in; Tape t; double g[MAXW];
        p.template cell_inputs<S>(i, in);
        forward<S>(m, th, in.x, p.stats, -1, t);
        backward<S>(m, th, in.x, t, p.bstats, nbar[i], l, g, nullptr, nullptr);
        for (int j = 0; j < m.width[l]; ++j) { s1 += g[j]; s2 += g[j] * t.xh[l][j]; }
      }
      const double M = (double)N * m.width[l];
      p.bstats[2 * l] = s1 / M;
      p.bstats[2 * l + 1] = s2 / M;
    }
  }
  for (int64_t i = 0; i < N; ++i) {
    Inputs in; Tape t; double xbar[3] = {0, 0, 0};
    p.template cell_inputs<S>(i, in);
    forward<S>(m, th, in.x, p.stats, -1, t);
    backward<S>(m, th, in.x, t, p.bstats, nbar[i], -1, nullptr, acc.data(), xbar);
    inputs_adj<S>(m, in, xbar, Qbar_add[i], Qbar_add[N + i], Qbar_add[2 * N + i]);
  }
}
template <class SS>
int check_spec(int id) {
  int bad = 0;
  if constexpr (SS::kSpec) {
    Model m;
    m.n_in = SS::NIN; m.n_hidden = SS::NH; m.ln_mode = SS::LN;
    for (int l = 0; l < SS::NH; ++l) { m.width[l] = SS::W; m.act[l] = SS::ACT; }
    canonical_offsets(m);
    bad += m.n_params != SS::P || spec_of(m) != id;
    for (int l = 0; l <= SS::NH; ++l) bad += m.off_w[l] != SS::off_w(l) || m.off_b[l] != SS::off_b(l);
    for (int l = 0; l < SS::NH; ++l) bad += m.off_g[l] != SS::off_g(l) || m.off_be[l] != SS::off_be(l);
  }
  return bad;
}
}  // namespace

extern "C" {
int ude_host_max_params() { return MAXP; }

// the compile-time offsets of every Spec against canonical_offsets() of the same shape; returns the number of mismatches
int ude_host_spec_check() {
  int bad = 0;
#define X(id, S) bad += check_spec<HG_UDE_UNPAREN S>(id);
  HG_UDE_SPECS(X)
#undef X
  return bad;
}

// which instantiation ran is returned through *spec (0 = generic); force_generic = 1 always runs the generic one
int ude_host_manning(const hg_ude_desc* d, int64_t N, const double* Q, const double* hstill, const double* ks, double hs,
                     const double* th, double* n_out, int force_generic, int* spec) {
  Pass p{};
  if (!p.setup(d, th)) return 1;
  p.N = N; p.Q = Q; p.hstill = hstill; p.ks = ks; p.hs = hs;
  const int id = force_generic ? 0 : spec_of(p.m);
  if (spec) *spec = id;
  switch (id) {
#define X(id_, S) case id_: manning_impl<HG_UDE_UNPAREN S>(p, n_out); break;
    HG_UDE_SPECS(X)
#undef X
  }
  return 0;
}

// Qbar_add[3N] = (dn/dQ)^T nbar, thbar[n_params] = (dn/dtheta)^T nbar
int ude_host_pullback(const hg_ude_desc* d, int64_t N, const double* Q, const double* hstill, const double* ks, double hs,
                      const double* th, const double* nbar, double* Qbar_add, double* thbar, int force_generic) {
  Pass p{};
  if (!p.setup(d, th)) return 1;
  p.N = N; p.Q = Q; p.hstill = hstill; p.ks = ks; p.hs = hs;
  std::vector<double> acc(p.m.n_params, 0.0);
  switch (force_generic ? 0 : spec_of(p.m)) {
#define X(id_, S) case id_: pullback_impl<HG_UDE_UNPAREN S>(p, nbar, Qbar_add, acc); break;
    HG_UDE_SPECS(X)
#undef X
  }
  for (int64_t k = 0; k < d->n_params; ++k) thbar[k] = 0.0;
  for (int k = 0; k < p.m.n_params; ++k) thbar[p.map.to_user[k]] = acc[k];
  return 0;
}
}
