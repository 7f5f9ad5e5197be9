// Forward mode of the RHS: dQ/dt and its directional derivative J_Q v + J_p pdot in one sweep -- what the reference gets
// when ForwardDiff pushes Dual numbers through swe_2d_rhs (swe_2D_sensitivity.jl:34-80 wraps the whole solve in
// ForwardDiff.jacobian; the ForwardDiffSensitivity / ForwardSensitivity inversion options do the same, solve_swe_2D.jl:230-235).
//
// Per-cell / per-boundary-entry arithmetic, templated on the scalar (double or Dual) and written in the reference's own
// operation order like the strict path (hg_plain.cu), which it restates:
//   swe_2d_rhs                semi_discretize_swe_2D.jl:18-277       process_all_boundaries_2d  bc_2D.jl:575-875
//   Riemann_2D_Roe            swe_2D_solvers.jl:4-164                 update_bed_data            process_bed_2D.jl:46-66
// Branches are taken on the values; clamps, wet flags and max() select constants (zero tangent), as ForwardDiff does.
// Plain host/device functions without any CUDA dependency: nvcc compiles them into the kernels of hg_jvp.cu, g++ into the
// CPU check of tests/test_jvp_cpu.py (tests/jvp_host.cpp), which compares them with the dual-number pass of the CPU restatement the tests hold.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HGJ_HD __host__ __device__ __forceinline__
#else
#define HGJ_HD inline
#endif

namespace hg {
namespace jvp {

constexpr double kEps = 2.220446049250313e-16;   // eps(Float64), utilities/smooth_functions.jl
enum : int32_t { kInletQ = 0, kExitH = 1, kWall = 2, kSymm = 3 };                    // BcType of hg_ctx.h
enum : int32_t { kParamNone = 0, kParamZb = 1, kParamManning = 2, kParamQ = 3 };     // HG_PARAM_*

struct Dual {
  double v, d;
};
HGJ_HD Dual mk(double v, double d) { Dual r; r.v = v; r.d = d; return r; }
HGJ_HD Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
HGJ_HD Dual operator+(Dual a, double b) { return mk(a.v + b, a.d); }
HGJ_HD Dual operator+(double a, Dual b) { return mk(a + b.v, b.d); }
HGJ_HD Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
HGJ_HD Dual operator-(Dual a, double b) { return mk(a.v - b, a.d); }
HGJ_HD Dual operator-(double a, Dual b) { return mk(a - b.v, -b.d); }
HGJ_HD Dual operator-(Dual a) { return mk(-a.v, -a.d); }
HGJ_HD Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
HGJ_HD Dual operator*(Dual a, double b) { return mk(a.v * b, a.d * b); }
HGJ_HD Dual operator*(double a, Dual b) { return mk(a * b.v, a * b.d); }
HGJ_HD Dual operator/(Dual a, Dual b) { const double q = a.v / b.v; return mk(q, (a.d - q * b.d) / b.v); }
HGJ_HD Dual operator/(Dual a, double b) { return mk(a.v / b, a.d / b); }
HGJ_HD Dual operator/(double a, Dual b) { const double q = a / b.v; return mk(q, -q * b.d / b.v); }

HGJ_HD double val(double x) { return x; }
HGJ_HD double val(Dual x) { return x.v; }
HGJ_HD double tangent(double) { return 0.0; }
HGJ_HD double tangent(Dual x) { return x.d; }
HGJ_HD double sqrt_(double x) { return sqrt(x); }
HGJ_HD Dual sqrt_(Dual x) { const double s = sqrt(x.v); return mk(s, x.d / (2.0 * s)); }
HGJ_HD double pow_(double x, double e) { return pow(x, e); }
HGJ_HD Dual pow_(Dual x, double e) { return mk(pow(x.v, e), e * pow(x.v, e - 1.0) * x.d); }
template <class T> HGJ_HD T lift(double v, double d);
template <> HGJ_HD double lift<double>(double v, double) { return v; }
template <> HGJ_HD Dual lift<Dual>(double v, double d) { return mk(v, d); }

template <class T> HGJ_HD T smooth_abs(T x) { return sqrt_(x * x + kEps); }
template <class T> HGJ_HD T smooth_sqrt(T x) { return sqrt_(x + kEps); }
template <class T> HGJ_HD T smooth_pow2(T x) { T y = x + kEps; return y * y; }

struct Args {
  int32_t N, B, n_inlet, active;
  double g, k_n, h_small;
  const int32_t *cf_ptr, *cf_nb;                    // CSR of cell faces; nb >= N is a ghost (N + ghost id)
  const double *cf_nx, *cf_ny, *cf_len;             // per cell-face
  const double *area, *hstill, *zb, *S0x, *S0y, *mann;
  const int32_t* matid;
  const int32_t *bc_type, *bc_group, *bc_ghost, *bc_cell, *inlet_ptr;   // boundary entries in processing order
  const double *bc_nx, *bc_ny, *bc_l53, *bc_l23;
  const double *hstill_g, *zb_g;                    // ghost order
  double *gh, *gqx, *gqy, *gxi;                     // ghost states [B], ghost order: values ...
  double *gh_d, *gqx_d, *gqy_d, *gxi_d;             // ... and tangents
  const double *Qin, *wse;
  const double *Q, *V;                              // state and its tangent [3N] (V may be NULL: zero)
  const double *params, *pdot;                      // active parameter and its tangent (pdot may be NULL: zero)
  double *dQ, *dQ_d;                                // outputs [3N] (dQ may be NULL)
  int32_t* err;                                     // set to 3 (HG_ERR_CONVEYANCE) by inlet_coef
};

template <class T>
HGJ_HD T param(const Args& a, int64_t k) { return lift<T>(a.params[k], a.pdot ? a.pdot[k] : 0.0); }

// clamp of semi_discretize_swe_2D.jl:101-106: h <= h_small -> h = h_small, q = 0 (constants); xi is not clamped
template <class T>
HGJ_HD void load_cell(const Args& a, int32_t i, T& xi, T& h, T& qx, T& qy) {
  const int64_t N = a.N;
  xi = lift<T>(a.Q[i], a.V ? a.V[i] : 0.0);
  h = xi + a.hstill[i];
  const bool dry = val(h) <= a.h_small;
  h = dry ? lift<T>(a.h_small, 0.0) : h;
  qx = dry ? lift<T>(0.0, 0.0) : lift<T>(a.Q[N + i], a.V ? a.V[N + i] : 0.0);
  qy = dry ? lift<T>(0.0, 0.0) : lift<T>(a.Q[2 * N + i], a.V ? a.V[2 * N + i] : 0.0);
}
template <class T>
HGJ_HD T mann_of(const Args& a, int32_t i) {
  return a.active == kParamManning ? param<T>(a, a.matid[i]) : lift<T>(a.mann[i], 0.0);
}
template <class T>
HGJ_HD T zb_of(const Args& a, int32_t i) {
  return a.active == kParamZb ? param<T>(a, i) : lift<T>(a.zb[i], 0.0);
}

// inlet-q boundary k: Q_k / sum_f L_f^(5/3) h_c / n_c wet_f (bc_2D.jl:665-691), sequential left fold like the reference
template <class T>
HGJ_HD T inlet_coef(const Args& a, int32_t k) {
  T tot = lift<T>(0.0, 0.0);
  bool first = true;
  for (int32_t e = a.inlet_ptr[k]; e < a.inlet_ptr[k + 1]; ++e) {
    T xi, h, qx, qy;
    const int32_t c = a.bc_cell[e];
    load_cell<T>(a, c, xi, h, qx, qy);
    const double wet = val(h) > a.h_small ? 1.0 : 0.0;
    const T term = a.bc_l53[e] * h / mann_of<T>(a, c) * wet;
    tot = first ? term : tot + term;
    first = false;
  }
  if (!(val(tot) > 1e-10)) *a.err = 3;                       // bc_2D.jl:678-680 (every thread would write the same value)
  const T Q = a.active == kParamQ ? param<T>(a, k) : lift<T>(a.Qin[k], 0.0);
  return Q / tot;
}

// ghost state of boundary entry e (bc_2D.jl:640-834), stored in ghost order
template <class T>
HGJ_HD void ghost_entry(const Args& a, int32_t e, T coef) {
  const double hs = a.h_small;
  const int32_t c = a.bc_cell[e], gi = a.bc_ghost[e], t = a.bc_type[e], k = a.bc_group[e];
  T xi, h, qx, qy, hg, gx, gy;
  load_cell<T>(a, c, xi, h, qx, qy);
  const double nx = a.bc_nx[e], ny = a.bc_ny[e];
  if (t == kInletQ) {
    const double wet = val(h) > hs ? 1.0 : 0.0;
    const T vn = coef * a.bc_l23[e] / mann_of<T>(a, c);
    hg = h;
    gx = -h * vn * nx * wet;
    gy = -h * vn * ny * wet;
  } else if (t == kExitH) {
    const T w = a.wse[k] - zb_of<T>(a, c);
    hg = val(w) > hs ? w : lift<T>(hs, 0.0);                 // max(h_small, WSE - zb)
    gx = qx; gy = qy;
  } else if (t == kWall) {
    hg = h; gx = -qx; gy = -qy;
  } else {
    const T vdn = qx * nx + qy * ny;
    hg = h;
    gx = qx - 2.0 * vdn * nx;
    gy = qy - 2.0 * vdn * ny;
  }
  const T gxi = hg - a.hstill_g[gi];
  a.gh[gi] = val(hg); a.gqx[gi] = val(gx); a.gqy[gi] = val(gy); a.gxi[gi] = val(gxi);
  a.gh_d[gi] = tangent(hg); a.gqx_d[gi] = tangent(gx); a.gqy_d[gi] = tangent(gy); a.gxi_d[gi] = tangent(gxi);
}

// swe_2D_solvers.jl:4-164 in the reference's operation order
template <class T>
HGJ_HD void roe(T xiL, double hstL, T hL, T huL, T hvL, T zbL, T xiR, double hstR, T hR, T huR, T hvR, T zbR, double g, double nx,
                double ny, double hmin, T& o0, T& o1, T& o2) {
  if (val(hL) <= hmin && val(hR) <= hmin) { o0 = o1 = o2 = lift<T>(0.0, 0.0); return; }
  else if (val(hL + zbL) < val(zbR + hmin) && val(hR) <= hmin) { hR = hL; huR = -huL; hvR = -hvL; }
  else if (val(hR + zbR) < val(zbL + hmin) && val(hL) <= hmin) { hL = hR; huL = -huR; hvL = -hvR; }
  else if (val(hL) <= hmin) {
    const T p = (0.5 * g) * smooth_pow2(hR);
    o0 = huR * nx + hvR * ny;
    o1 = (huR * (huR / hR) + p) * nx + huR * (hvR / hR) * ny;
    o2 = (hvR * (huR / hR)) * nx + (hvR * (hvR / hR) + p) * ny;
    return;
  } else if (val(hR) <= hmin) {
    const T p = (0.5 * g) * smooth_pow2(hL);
    o0 = huL * nx + hvL * ny;
    o1 = (huL * (huL / hL) + p) * nx + huL * (hvL / hL) * ny;
    o2 = (hvL * (huL / hL)) * nx + (hvL * (hvL / hL) + p) * ny;
    return;
  }
  const T uL = huL / hL, vL = hvL / hL, uR = huR / hR, vR = hvR / hR;
  const T sL = smooth_sqrt(hL), sR = smooth_sqrt(hR);
  const T hRoe = (hL + hR) / 2.0;
  const T uRoe = (sL * uL + sR * uR) / (sL + sR);
  const T vRoe = (sL * vL + sR * vR) / (sL + sR);
  const T un = uRoe * nx + vRoe * ny;
  const T c = smooth_sqrt(g * hRoe);
  const T o2c = 1.0 / 2.0 / c;
  const T R22 = uRoe - c * nx, R23 = uRoe + c * nx, R32 = vRoe - c * ny, R33 = vRoe + c * ny;
  const T L11 = -(uRoe * ny - vRoe * nx);
  const double L12 = ny, L13 = -nx;
  const T L21 = un * o2c + 0.5, L22 = -nx * o2c, L23 = -ny * o2c;
  const T L31 = -un * o2c + 0.5, L32 = nx * o2c, L33 = ny * o2c;
  const T a1 = smooth_abs(un), a2 = smooth_abs(un - c), a3 = smooth_abs(un + c);
  const T d1 = xiR - xiL, d2 = huR - huL, d3 = hvR - hvL;
  const T w1 = (L11 * d1 + L12 * d2) + L13 * d3;
  const T w2 = (L21 * d1 + L22 * d2) + L23 * d3;
  const T w3 = (L31 * d1 + L32 * d2) + L33 * d3;
  const T z1 = a1 * w1, z2 = a2 * w2, z3 = a3 * w3;
  const T y1 = z2 + z3;
  const T y2 = (ny * z1 + R22 * z2) + R23 * z3;
  const T y3 = (-nx * z1 + R32 * z2) + R33 * z3;
  const T pL = (0.5 * g) * (smooth_pow2(xiL) + 2.0 * xiL * hstL);
  const T pR = (0.5 * g) * (smooth_pow2(xiR) + 2.0 * xiR * hstR);
  const T f1L = huL * nx + hvL * ny;
  const T f2L = (huL * uL + pL) * nx + huL * vL * ny;
  const T f3L = (hvL * uL) * nx + (hvL * vL + pL) * ny;
  const T f1R = huR * nx + hvR * ny;
  const T f2R = (huR * uR + pR) * nx + huR * vR * ny;
  const T f3R = (hvR * uR) * nx + (hvR * vR + pR) * ny;
  o0 = (f1L + f1R - y1) / 2.0;
  o1 = (f2L + f2R - y2) / 2.0;
  o2 = (f3L + f3R - y3) / 2.0;
}

// compute_inviscid_fluxes + compute_source_terms of cell i (every interior face from both of its cells, left-to-right sums)
template <class T>
HGJ_HD void cell(const Args& a, int32_t i) {
  const int64_t N = a.N;
  const double g = a.g, hs = a.h_small;
  T xi, h, qx, qy;
  load_cell<T>(a, i, xi, h, qx, qy);
  const T zbi = zb_of<T>(a, i);
  const double hsti = a.hstill[i];
  T s0 = lift<T>(0.0, 0.0), s1 = s0, s2 = s0, gx = s0, gy = s0;
  for (int32_t k = a.cf_ptr[i]; k < a.cf_ptr[i + 1]; ++k) {
    const int32_t r = a.cf_nb[k];
    const double nx = a.cf_nx[k], ny = a.cf_ny[k], L = a.cf_len[k];
    T xr, hr, qxr, qyr, zbr;
    double hstr;
    if (r < a.N) {
      load_cell<T>(a, r, xr, hr, qxr, qyr);
      zbr = zb_of<T>(a, r);
      hstr = a.hstill[r];
    } else {
      const int32_t gi = r - a.N;
      xr = lift<T>(a.gxi[gi], a.gxi_d[gi]); hr = lift<T>(a.gh[gi], a.gh_d[gi]);
      qxr = lift<T>(a.gqx[gi], a.gqx_d[gi]); qyr = lift<T>(a.gqy[gi], a.gqy_d[gi]);
      hstr = a.hstill_g[gi];
      zbr = a.active == kParamZb ? zbi : lift<T>(a.zb_g[gi], 0.0);        // zb_ghost = zb of the internal cell (fvm_schemes_2D.jl:3-30)
    }
    T f0, f1, f2;
    roe<T>(xi, hsti, h, qx, qy, zbi, xr, hstr, hr, qxr, qyr, zbr, g, nx, ny, hs, f0, f1, f2);
    s0 = s0 + f0 * L; s1 = s1 + f1 * L; s2 = s2 + f2 * L;
    if (a.active == kParamZb) {
      const T zf = (r < a.N) ? (zbi + zbr) / 2.0 : zbi;                    // cells_to_faces_scalar, fvm_schemes_2D.jl:89-105
      gx = gx + nx * zf * L; gy = gy + ny * zf * L;
    }
  }
  const double A = a.area[i];
  T S0x, S0y;
  if (a.active == kParamZb) { S0x = -1.0 * (gx / A); S0y = -1.0 * (gy / A); }
  else { S0x = lift<T>(a.S0x[i], 0.0); S0y = lift<T>(a.S0y[i], 0.0); }
  const T n = mann_of<T>(a, i);
  const T mag = smooth_sqrt(qx * qx + qy * qy);
  const T coef = g * (n * n) / (a.k_n * a.k_n) / pow_(h + hs, 7.0 / 3.0);
  const T frx = coef * mag * qx, fry = coef * mag * qy;
  const double wet = val(h) > hs ? 1.0 : 0.0;
  const T r0 = -s0 / A + 0.0;
  const T r1 = -s1 / A + wet * (g * xi * S0x - frx);
  const T r2 = -s2 / A + wet * (g * xi * S0y - fry);
  if (a.dQ) { a.dQ[i] = val(r0); a.dQ[N + i] = val(r1); a.dQ[2 * N + i] = val(r2); }
  a.dQ_d[i] = tangent(r0); a.dQ_d[N + i] = tangent(r1); a.dQ_d[2 * N + i] = tangent(r2);
}

}  // namespace jvp
}  // namespace hg
