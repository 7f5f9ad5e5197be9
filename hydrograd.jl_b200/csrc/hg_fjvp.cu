// Fused forward-mode (JVP) tile kernel: dQ/dt and J_Q v + J_p pdot of the 2-D shallow-water RHS in ONE pass over the staged
// tile -- the device counterpart of ForwardDiff.Dual flowing through swe_2d_rhs (ForwardDiff.jacobian around the solve,
// applications/sensitivity/swe_2D_sensitivity.jl:34-80; the ForwardDiffSensitivity / ForwardSensitivity options,
// applications/solve_swe_2D.jl:230-235) on the PERFORMANCE path: same tile tables, TMA staging and face-once evaluation as
// k_fused_rhs (hg_fused.cu), every quantity carried as a dual number (value, tangent).  Branches are taken on the values;
// clamps, wet flags and max() select constants -- exactly what ForwardDiff does to the reference.  K directions are batched in
// one launch (CTA b: tile b / K, direction b % K): the mesh tables and the state of a tile are shared by its K CTAs through L2.
//   phase 1  raw (xi, q_x, q_y) and their tangents -> clamped, derived duals (u, v, sqrt(h+eps), xi-form pressure) per cell
//   phase 2  every face once: dual Roe flux (Riemann_2D_Roe, swe_2D_solvers.jl:4-164); boundary faces build the dual ghost
//            state from the owned cell (bc_2D.jl:640-834) -- inlet-q through the dual conveyance coefficient, exit-h through
//            the bed tangent, wall, symm
//   phase 3  per-cell gather in the reference's face order + dual bed-slope / Manning-friction sources
// Parameter tangents: ManningN zones -> a per-cell n-dot (zone gather of pdot), zb -> bed and slope tangents (update_bed_data
// is linear: the same kernel applied to pdot), Q -> inlet discharge tangents.
#include "hg_device.cuh"

namespace hg {
namespace {
using namespace dev;

struct D1 {
  double v, d;
};
__device__ __forceinline__ D1 mk(double v, double d = 0.0) { return D1{v, d}; }
__device__ __forceinline__ D1 operator+(D1 a, D1 b) { return D1{a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ D1 operator-(D1 a, D1 b) { return D1{a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ D1 operator-(D1 a) { return D1{-a.v, -a.d}; }
__device__ __forceinline__ D1 operator*(D1 a, D1 b) { return D1{a.v * b.v, fma(a.v, b.d, a.d * b.v)}; }
__device__ __forceinline__ D1 operator*(double a, D1 b) { return D1{a * b.v, a * b.d}; }
__device__ __forceinline__ D1 operator*(D1 b, double a) { return D1{a * b.v, a * b.d}; }
__device__ __forceinline__ D1 operator+(D1 a, double b) { return D1{a.v + b, a.d}; }
__device__ __forceinline__ D1 operator-(D1 a, double b) { return D1{a.v - b, a.d}; }
__device__ __forceinline__ D1 rcp(D1 x) {                    // 1/x
  const double r = fast_rcp(x.v);
  return D1{r, -r * r * x.d};
}
__device__ __forceinline__ D1 rsq(D1 x) {                    // x^(-1/2)
  const double r = fast_rsqrt(x.v);
  return D1{r, -0.5 * r * r * r * x.d};
}
__device__ __forceinline__ D1 dsqrt(D1 x) {                  // sqrt(x), x > 0
  const double r = fast_rsqrt(x.v);
  const double s = x.v * r;
  return D1{fma(fma(-s, s, x.v), 0.5 * r, s), 0.5 * r * x.d};
}
__device__ __forceinline__ D1 sabs(D1 x) {                   // sqrt(x^2 + eps), utilities/smooth_functions.jl
  const double y = fma(x.v, x.v, EPS);
  const double r = fast_rsqrt(y);
  return D1{y * r, x.v * r * x.d};
}
__device__ __forceinline__ D1 pow_m73_d(D1 x) {              // x^(-7/3)
  const double w = rcbrt_pos(x.v), w2 = w * w, w3 = w2 * w;
  const double p = w3 * w3 * w;
  return D1{p, -(7.0 / 3.0) * p * w3 * x.d};
}

struct SideD {
  D1 xi, h, hu, hv, u, v, s, P;
};
__device__ __forceinline__ void derive_d(SideD& s, double hst, double g) {
  const D1 rh = rcp(s.h);
  s.u = s.hu * rh;
  s.v = s.hv * rh;
  s.s = dsqrt(s.h + EPS);
  const D1 xe = s.xi + EPS;
  s.P = (0.5 * g) * (xe * xe + (2.0 * hst) * s.xi);   // xi-form pressure, swe_2D_solvers.jl:122
}

// Riemann_2D_Roe on duals, face-once form, same statement order as dev::roe_flux; returns flux * len
template <class ZL, class ZR>
__device__ __forceinline__ void roe_flux_d(SideD L, SideD R, ZL zbL, ZR zbR, double nx, double ny, double len, double g,
                                           double hmin, D1& o0, D1& o1, D1& o2) {
  const bool dryL = L.h.v <= hmin, dryR = R.h.v <= hmin;
  if (dryL || dryR) {
    if (dryL && dryR) { o0 = o1 = o2 = mk(0.0); return; }
    const double zl = zbL(), zr = zbR();
    if ((L.h.v + zl) < (zr + hmin) && dryR) {
      R.h = L.h; R.hu = -L.hu; R.hv = -L.hv; R.u = -L.u; R.v = -L.v; R.s = L.s;
    } else if ((R.h.v + zr) < (zl + hmin) && dryL) {
      L.h = R.h; L.hu = -R.hu; L.hv = -R.hv; L.u = -R.u; L.v = -R.v; L.s = R.s;
    } else {
      const SideD& W = dryL ? R : L;
      const D1 hp = W.h + EPS;
      const D1 p = (0.5 * g) * (hp * hp);
      const D1 un = nx * W.u + ny * W.v;
      o0 = len * (nx * W.hu + ny * W.hv);
      o1 = len * (W.hu * un + nx * p);
      o2 = len * (W.hv * un + ny * p);
      return;
    }
  }
  const D1 rs = rcp(L.s + R.s);
  const D1 uRoe = (L.s * L.u + R.s * R.u) * rs;
  const D1 vRoe = (L.s * L.v + R.s * R.v) * rs;
  const D1 un = nx * uRoe + ny * vRoe;
  const D1 c2 = (0.5 * g) * (L.h + R.h) + EPS;
  const D1 rc = rsq(c2);
  const D1 c = c2 * rc;
  const D1 k = 0.5 * rc;
  const D1 d1 = R.xi - L.xi, d2 = R.hu - L.hu, d3 = R.hv - L.hv;
  const D1 w1 = -((ny * uRoe - nx * vRoe) * d1) + ny * d2 - nx * d3;
  const D1 m = k * (un * d1 - (nx * d2 + ny * d3));
  const D1 w2 = 0.5 * d1 + m, w3 = 0.5 * d1 - m;
  const D1 z1 = sabs(un) * w1, z2 = sabs(un - c) * w2, z3 = sabs(un + c) * w3;
  const D1 zs = z2 + z3, zd = c * (z3 - z2);
  const D1 y2 = ny * z1 + uRoe * zs + nx * zd;
  const D1 y3 = -(nx * z1) + vRoe * zs + ny * zd;
  const D1 unL = nx * L.u + ny * L.v, unR = nx * R.u + ny * R.v;
  const D1 ps = L.P + R.P;
  const double hl = 0.5 * len;
  o0 = hl * ((nx * L.hu + ny * L.hv) + (nx * R.hu + ny * R.hv) - zs);
  o1 = hl * (L.hu * unL + R.hu * unR + nx * ps - y2);
  o2 = hl * (L.hv * unL + R.hv * unR + ny * ps - y3);
}

struct FjvpArgs {
  int32_t N, n_tiles, active;
  int32_t K;                             // directions per launch: CTA b works on tile b / K, direction b % K (direction fastest:
                                         // the K CTAs of a tile run back to back and share its mesh tables and state through L2)
  int64_t Ns;
  Consts c;
  const int32_t *tile_desc, *halo, *bface_e;
  const uint32_t* face_lr;
  const uint16_t* cf_idx;
  const double *face_nx, *face_ny, *face_len;
  const double *area, *hstill, *zb, *S0x, *S0y, *mann;
  const int32_t *bc_type, *bc_group;
  const double *bc_nx, *bc_ny, *bc_l23, *bc_hstill, *bc_zb, *wse;
  const double *coef_v, *coef_d;         // [K][n_inlet] dual conveyance coefficients (k_fjvp_inlet)
  const double *Q, *V;                   // state [3Ns]; tangents [K][3Ns] (row stride sV)
  const double *mann_d, *zb_d, *S0x_d, *S0y_d;   // per-cell parameter tangents [K][Ns] or NULL
  double *out, *out_d;                   // dQ/dt [3Ns] or NULL; its tangents [K][3Ns]
  int64_t sV, sP, sI;                    // row strides: state-like (3 Ns), per-cell parameter tangents (Ns), per-inlet
};

template <int T, int ML, int MF, int NF>
struct __align__(16) FjvpSmem {
  uint64_t bar[2];
  double xi[ML], h[ML], u[ML], v[ML], s[ML], P[ML];         // values (raw xi, q_x, q_y, hstill on arrival: in place)
  double dxi[ML], dh[ML], du[ML], dv[ML], ds[ML], dP[ML];   // tangents (raw tangents of xi, q_x, q_y on arrival)
  double f[3][MF];                                          // nx, ny, len on arrival; value fluxes * len after phase 2
  double gt[3][MF];                                         // tangent fluxes * len
  uint32_t lr[MF];
  uint16_t cf[T * NF];
};

// dual conveyance-weighted inlet split (bc_2D.jl:665-691): coef_k = Q_k / sum_f L_f^(5/3) h_c / n_c wet_f; one CTA per
// (inlet, direction), fixed-shape tree reductions of value and tangent
__global__ void __launch_bounds__(256) k_fjvp_inlet(Consts c, int32_t n_inlet, int64_t Ns, const int32_t* __restrict__ inlet_ptr,
                                                    const int32_t* __restrict__ bc_cell, const double* __restrict__ bc_l53,
                                                    const double* __restrict__ Q, const double* __restrict__ V, int64_t sV,
                                                    const double* __restrict__ hstill, const double* __restrict__ mann,
                                                    const double* __restrict__ mann_d, int64_t sP, const double* __restrict__ Qin,
                                                    const double* __restrict__ Qin_d, double* __restrict__ coef_v,
                                                    double* __restrict__ coef_d, int32_t* err) {
  __shared__ double rv[256], rd[256];
  const int k = blockIdx.x, y = blockIdx.y;
  if (V) V += (int64_t)y * sV;
  if (mann_d) mann_d += (int64_t)y * sP;
  double av = 0.0, ad = 0.0;
  for (int32_t e = inlet_ptr[k] + threadIdx.x; e < inlet_ptr[k + 1]; e += 256) {
    const int32_t ci = bc_cell[e];
    const double h0 = Q[ci] + hstill[ci];
    if (h0 > c.h_small) {                           // clamped (dry) cells contribute nothing, value or tangent
      const D1 h = mk(h0, V ? V[ci] : 0.0);
      const D1 n = mk(mann[ci], mann_d ? mann_d[ci] : 0.0);
      const D1 t = bc_l53[e] * (h * rcp(n));
      av += t.v; ad += t.d;
    }
  }
  rv[threadIdx.x] = av; rd[threadIdx.x] = ad;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { rv[threadIdx.x] += rv[threadIdx.x + s]; rd[threadIdx.x] += rd[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (!(rv[0] > 1e-10)) atomicExch(err, HG_ERR_CONVEYANCE);
    const D1 A = mk(rv[0], rd[0]);
    const D1 q = mk(Qin[k], Qin_d ? Qin_d[(int64_t)y * n_inlet + k] : 0.0);
    const D1 cf = q * rcp(A);
    coef_v[(int64_t)y * n_inlet + k] = cf.v; coef_d[(int64_t)y * n_inlet + k] = cf.d;
  }
}

template <int T, int ML, int MF, int NF, int TH, int MB>
__global__ void __launch_bounds__(TH, MB) k_fused_jvp(const __grid_constant__ FjvpArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using Smem = FjvpSmem<T, ML, MF, NF>;
  Smem& sm = *reinterpret_cast<Smem*>(smraw);
  const int tid = threadIdx.x;
  const int t = (int)(blockIdx.x / (unsigned)a.K), y = (int)(blockIdx.x % (unsigned)a.K);
  const int4 d0 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc));
  const int4 d1 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc) + 1);
  const int4 d2 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc) + 2);
  const int32_t c0 = d0.x, nc = d0.y, hp = d0.z, nh = d0.w;
  const int32_t fp = d1.x, nf = d1.y, nfp = d1.z;
  const int32_t nint = d2.y, bfp = d2.z;
  const int32_t ncp = (nc + 1) & ~1;
  const double g = a.c.g, hs = a.c.h_small;
  const int64_t Ns = a.Ns;
  const double* __restrict__ V = a.V + (int64_t)y * a.sV;
  double* __restrict__ out_d = a.out_d + (int64_t)y * a.sV;
  double* __restrict__ out = (a.out && y == 0) ? a.out : nullptr;     // the values are delivered once
  const double* mann_d = a.mann_d ? a.mann_d + (int64_t)y * a.sP : nullptr;
  const double* zb_d = a.zb_d ? a.zb_d + (int64_t)y * a.sP : nullptr;
  const double* S0x_d = a.S0x_d ? a.S0x_d + (int64_t)y * a.sP : nullptr;
  const double* S0y_d = a.S0y_d ? a.S0y_d + (int64_t)y * a.sP : nullptr;
  const double* coef_v = a.coef_v + (int64_t)y * a.sI;
  const double* coef_d = a.coef_d + (int64_t)y * a.sI;

  if (tid == 0) mbar_init(sm.bar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t cb = (uint32_t)ncp * 8u, fb = (uint32_t)nfp * 8u;
    mbar_expect_tx(sm.bar, 7u * cb + 3u * fb + (uint32_t)nfp * 4u + (uint32_t)(T * NF) * 2u);
    bulk_g2s(sm.xi, a.Q + c0, cb, sm.bar);
    bulk_g2s(sm.u, a.Q + Ns + c0, cb, sm.bar);
    bulk_g2s(sm.v, a.Q + 2 * Ns + c0, cb, sm.bar);
    bulk_g2s(sm.P, a.hstill + c0, cb, sm.bar);
    bulk_g2s(sm.dxi, V + c0, cb, sm.bar);
    bulk_g2s(sm.du, V + Ns + c0, cb, sm.bar);
    bulk_g2s(sm.dv, V + 2 * Ns + c0, cb, sm.bar);
    bulk_g2s(sm.f[0], a.face_nx + fp, fb, sm.bar);
    bulk_g2s(sm.f[1], a.face_ny + fp, fb, sm.bar);
    bulk_g2s(sm.f[2], a.face_len + fp, fb, sm.bar);
    bulk_g2s(sm.lr, a.face_lr + fp, (uint32_t)nfp * 4u, sm.bar);
    bulk_g2s(sm.cf, a.cf_idx + (size_t)t * (T * NF), (uint32_t)(T * NF) * 2u, sm.bar);
  }
  // clamp (semi_discretize_swe_2D.jl:101-106: a clamped cell's h, q are constants; xi is not clamped) + derived duals
  auto stage_cell = [&](int32_t l, D1 xi, D1 qx, D1 qy, double hst) {
    SideD s;
    s.xi = xi;
    const D1 h0 = xi + hst;
    const bool dry = h0.v <= hs;
    s.h = dry ? mk(hs) : h0; s.hu = dry ? mk(0.0) : qx; s.hv = dry ? mk(0.0) : qy;
    derive_d(s, hst, g);
    sm.xi[l] = s.xi.v; sm.h[l] = s.h.v; sm.u[l] = s.u.v; sm.v[l] = s.v.v; sm.s[l] = s.s.v; sm.P[l] = s.P.v;
    sm.dxi[l] = s.xi.d; sm.dh[l] = s.h.d; sm.du[l] = s.u.d; sm.dv[l] = s.v.d; sm.ds[l] = s.s.d; sm.dP[l] = s.P.d;
  };
  for (int32_t k = tid; k < nh; k += TH) {
    const int32_t gi = __ldg(a.halo + hp + k);
    stage_cell(ncp + k, mk(a.Q[gi], V[gi]), mk(a.Q[Ns + gi], V[Ns + gi]), mk(a.Q[2 * Ns + gi], V[2 * Ns + gi]), a.hstill[gi]);
  }
  mbar_wait(sm.bar, 0);
  for (int32_t l = tid; l < nc; l += TH)
    stage_cell(l, mk(sm.xi[l], sm.dxi[l]), mk(sm.u[l], sm.du[l]), mk(sm.v[l], sm.dv[l]), sm.P[l]);
  __syncthreads();

  // ---- phase 2: every face once
  auto load_side = [&](int32_t l, SideD& S) {
    S.xi = mk(sm.xi[l], sm.dxi[l]); S.h = mk(sm.h[l], sm.dh[l]); S.u = mk(sm.u[l], sm.du[l]); S.v = mk(sm.v[l], sm.dv[l]);
    S.s = mk(sm.s[l], sm.ds[l]); S.P = mk(sm.P[l], sm.dP[l]);
    S.hu = S.h * S.u; S.hv = S.h * S.v;
  };
  auto zb_of = [&](int32_t l) { return a.zb[l < ncp ? c0 + l : __ldg(a.halo + hp + (l - ncp))]; };
  for (int32_t f = tid; f < nf; f += TH) {
    const uint32_t lr = sm.lr[f];
    const int32_t lL = lr & 0xFFFFu, lR = lr >> 16;
    const double nx = sm.f[0][f], ny = sm.f[1][f], len = sm.f[2][f];
    SideD L, R;
    load_side(lL, L);
    double zbl = 0.0, zbr = 0.0;
    const bool interior = f < nint;
    if (interior) {
      load_side(lR, R);
    } else {
      const int32_t gc = c0 + lL;                                  // boundary faces always touch an owned cell
      zbl = a.zb[gc];
      const int32_t e = __ldg(a.bface_e + bfp + (f - nint));
      const int32_t ty = a.bc_type[e], kgrp = a.bc_group[e];
      const double bnx = a.bc_nx[e], bny = a.bc_ny[e];
      const double hst = a.bc_hstill[e];
      if (ty == BC_INLETQ) {
        const double wet = L.h.v > hs ? 1.0 : 0.0;
        const D1 nc_ = mk(a.mann[gc], mann_d ? mann_d[gc] : 0.0);
        const D1 vn = (a.bc_l23[e] * mk(coef_v[kgrp], coef_d[kgrp])) * rcp(nc_);
        const D1 hv_ = L.h * vn;
        R.h = L.h; R.hu = (-bnx * wet) * hv_; R.hv = (-bny * wet) * hv_;
      } else if (ty == BC_EXITH) {
        const D1 hg = mk(a.wse[kgrp] - zbl, zb_d ? -zb_d[gc] : 0.0);
        R.h = hg.v > hs ? hg : mk(hs);                              // max(h_small, .): the clamp passes no tangent
        R.hu = L.hu; R.hv = L.hv;
      } else if (ty == BC_WALL) {
        R.h = L.h; R.hu = -L.hu; R.hv = -L.hv;
      } else {                                                      // symmetry
        const D1 vdn = bnx * L.hu + bny * L.hv;
        R.h = L.h; R.hu = L.hu - (2.0 * bnx) * vdn; R.hv = L.hv - (2.0 * bny) * vdn;
      }
      R.xi = R.h - hst;                                             // semi_discretize_swe_2D.jl:220
      zbr = a.bc_zb[e];
      derive_d(R, hst, g);
    }
    D1 o0, o1, o2;
    roe_flux_d(L, R, [&] { return interior ? zb_of(lL) : zbl; }, [&] { return interior ? zb_of(lR) : zbr; }, nx, ny, len, g, hs, o0, o1, o2);
    sm.f[0][f] = o0.v; sm.f[1][f] = o1.v; sm.f[2][f] = o2.v;
    sm.gt[0][f] = o0.d; sm.gt[1][f] = o1.d; sm.gt[2][f] = o2.d;
  }
  if (tid < 3) { sm.f[tid][nfp] = 0.0; sm.gt[tid][nfp] = 0.0; }      // the zero-flux slot of unused cf entries
  __syncthreads();

  // ---- phase 3: per-cell gather (reference face order) + dual sources
  const double kfr = g / (a.c.k_n * a.c.k_n);
  for (int32_t l = tid; l < nc; l += TH) {
    const int32_t gi = c0 + l;
    D1 s0 = mk(0.0), s1 = mk(0.0), s2 = mk(0.0);
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const uint32_t ix = sm.cf[l * NF + j];
      const int32_t f = ix & 0x7FFF;
      const double sg = (ix & 0x8000) ? -1.0 : 1.0;
      s0.v = fma(sg, sm.f[0][f], s0.v); s1.v = fma(sg, sm.f[1][f], s1.v); s2.v = fma(sg, sm.f[2][f], s2.v);
      s0.d = fma(sg, sm.gt[0][f], s0.d); s1.d = fma(sg, sm.gt[1][f], s1.d); s2.d = fma(sg, sm.gt[2][f], s2.d);
    }
    const double rA = -fast_rcp(a.area[gi]);
    const D1 xi = mk(sm.xi[l], sm.dxi[l]), h = mk(sm.h[l], sm.dh[l]);
    const D1 qx = h * mk(sm.u[l], sm.du[l]), qy = h * mk(sm.v[l], sm.dv[l]);
    const D1 n = mk(a.mann[gi], mann_d ? mann_d[gi] : 0.0);
    const D1 mag = dsqrt(qx * qx + qy * qy + EPS);
    const D1 coef = (kfr * (n * n)) * pow_m73_d(h + hs) * mag;      // g n^2/k_n^2/(h+hs)^(7/3) |q|
    const bool wet = h.v > hs;
    const D1 sx = mk(a.S0x[gi], S0x_d ? S0x_d[gi] : 0.0), sy = mk(a.S0y[gi], S0y_d ? S0y_d[gi] : 0.0);
    D1 r0 = rA * s0, r1 = rA * s1, r2 = rA * s2;
    if (wet) {
      r1 = r1 + (g * (xi * sx) - coef * qx);
      r2 = r2 + (g * (xi * sy) - coef * qy);
    }
    if (out) { out[gi] = r0.v; out[Ns + gi] = r1.v; out[2 * Ns + gi] = r2.v; }
    out_d[gi] = r0.d; out_d[Ns + gi] = r1.d; out_d[2 * Ns + gi] = r2.d;
  }
}

// per-cell n-dot from the zone tangents: process_ManningN_2D.jl:88 applied to pdot; K rows
__global__ void k_fjvp_expand(int32_t N, int64_t Ns, int32_t n_mat, const int32_t* __restrict__ matid, const double* __restrict__ pdot,
                              double* __restrict__ mann_d) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) mann_d[(int64_t)blockIdx.y * Ns + i] = pdot[(int64_t)blockIdx.y * n_mat + matid[i]];
}

struct FjvpKernel {
  const void* fn = nullptr;
  int threads = 0, smem = 0;
};
template <int T, int ML, int MF, int NF>
FjvpKernel fjvp_pick() {
  // one face per thread and trip; 128 threads x 3 CTAs/SM (T = 256: 72 KB of shared memory per CTA)
  constexpr int TH = T >= 512 ? 256 : 128;
  constexpr int MB = T >= 384 ? 1 : (T == 256 ? 3 : (T >= 192 ? 3 : 4));
  return FjvpKernel{(const void*)k_fused_jvp<T, ML, MF, NF, TH, MB>, TH, (int)sizeof(FjvpSmem<T, ML, MF, NF>)};
}
FjvpKernel fjvp_kernel(int cfg_id) {
  switch (cfg_id) {
#define X(id, T, ML, MF, NF, TH, MB) case id: return fjvp_pick<T, ML, MF, NF>();
    HG_TILE_CONFIGS(X)
#undef X
  }
  return FjvpKernel{};
}

}  // namespace

// K directions: d_V [K][3 Ns] tangents of the state (internal order), d_pdot [K][n_params] tangents of the active parameter
// (device, the caller's parameter order; NULL = zero), d_out [3 Ns] values or NULL, d_out_d [K][3 Ns].
int fused_jvp(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_V, const double* d_pdot, double* d_out, double* d_out_d, int64_t K) {
  FusedDev& d = ctx->fd;
  const FusedHost& fh = ctx->fh;
  if (ctx->n_halo > 0) { ctx->err = "forward mode: multi-rank contexts are not supported"; return HG_ERR_ARG; }
  if (ctx->mfn.type || ctx->active == HG_PARAM_UDE) { ctx->err = "forward mode: state-dependent Manning closures / the UDE network have no forward mode"; return HG_ERR_ARG; }
  if (K < 1 || K > 65535) { ctx->err = "forward mode: number of directions out of range"; return HG_ERR_ARG; }
  const FjvpKernel kk = fjvp_kernel(cfg_id);
  if (!kk.fn || kk.smem > 227 * 1024) { ctx->err = "forward mode: no tile configuration"; return HG_ERR_ARG; }
  if (!ctx->fjvp_ready) {
    cudaError_t e = cudaFuncSetAttribute(kk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kk.smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kk.fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute(jvp): ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
    ctx->fjvp_ready = true;
  }
  const size_t Ns = (size_t)fh.Ns, nI = (size_t)std::max<int64_t>(ctx->n_inletq, 1);
  const int th = 256;
  const bool with_p = d_pdot != nullptr && ctx->active != HG_PARAM_NONE;
  FjvpArgs a;
  a.N = (int32_t)ctx->N; a.n_tiles = fh.n_tiles; a.active = ctx->active; a.K = (int32_t)K; a.Ns = fh.Ns; a.c = ctx->c;
  a.tile_desc = d.tile_desc.p; a.halo = d.halo.p; a.bface_e = d.bface_e.p; a.face_lr = d.face_lr.p; a.cf_idx = d.cf_idx.p;
  a.face_nx = d.face_nx.p; a.face_ny = d.face_ny.p; a.face_len = d.face_len.p;
  a.area = d.area.p; a.hstill = d.hstill.p; a.zb = d.zb.p; a.S0x = d.S0x.p; a.S0y = d.S0y.p; a.mann = d.mann.p;
  a.bc_type = d.bc_type.p; a.bc_group = d.bc_group.p; a.bc_nx = d.bc_nx.p; a.bc_ny = d.bc_ny.p; a.bc_l23 = d.bc_l23.p;
  a.bc_hstill = d.bc_hstill.p; a.bc_zb = d.bc_zb.p; a.wse = d.wse.p;
  a.Q = d_Q; a.V = d_V; a.out = d_out; a.out_d = d_out_d;
  a.sV = 3 * (int64_t)Ns; a.sP = (int64_t)Ns; a.sI = (int64_t)nI;
  a.mann_d = a.zb_d = a.S0x_d = a.S0y_d = nullptr;
  const double* Qin_d = nullptr;
  // ---- per-cell parameter tangents of the K directions
  if (with_p && ctx->active == HG_PARAM_MANNING) {
    if (d.j_p0.n < (size_t)K * Ns && d.j_p0.alloc((size_t)K * Ns) != cudaSuccess) { ctx->err = "cudaMalloc(jvp)"; return HG_ERR_CUDA; }
    k_fjvp_expand<<<dim3((unsigned)((ctx->N + th - 1) / th), (unsigned)K), th, 0, ctx->stream>>>((int32_t)ctx->N, fh.Ns, (int32_t)ctx->n_mat,
                                                                                              d.matid.p, d_pdot, d.j_p0.p);
    ctx->launches++;
    a.mann_d = d.j_p0.p;
  } else if (with_p && ctx->active == HG_PARAM_ZB) {
    if (d.j_p0.n < (size_t)K * Ns) {
      if (d.j_p0.alloc((size_t)K * Ns) != cudaSuccess) { ctx->err = "cudaMalloc(jvp)"; return HG_ERR_CUDA; }
    }
    if (d.j_p1.n < (size_t)K * Ns && (d.j_p1.alloc((size_t)K * Ns) != cudaSuccess || d.j_p2.alloc((size_t)K * Ns) != cudaSuccess)) {
      ctx->err = "cudaMalloc(jvp)";
      return HG_ERR_CUDA;
    }
    for (int64_t k = 0; k < K; ++k) {     // update_bed_data is linear in zb: the binding kernel applied to the tangent
      const int rc = fused_bed_from(ctx, d_pdot + k * ctx->N, d.j_p0.p + k * Ns, d.j_p1.p + k * Ns, d.j_p2.p + k * Ns);
      if (rc != HG_OK) return rc;
    }
    a.zb_d = d.j_p0.p; a.S0x_d = d.j_p1.p; a.S0y_d = d.j_p2.p;
  } else if (with_p && ctx->active == HG_PARAM_Q) {
    Qin_d = d_pdot;       // [K][n_inletq] already
  }
  // ---- dual inlet coefficients
  if (d.j_coef.n < 2 * (size_t)K * nI && d.j_coef.alloc(2 * (size_t)K * nI) != cudaSuccess) { ctx->err = "cudaMalloc(jvp)"; return HG_ERR_CUDA; }
  a.coef_v = d.j_coef.p; a.coef_d = d.j_coef.p + (size_t)K * nI;
  if (ctx->n_inletq > 0) {
    k_fjvp_inlet<<<dim3((unsigned)ctx->n_inletq, (unsigned)K), 256, 0, ctx->stream>>>(
        ctx->c, (int32_t)ctx->n_inletq, fh.Ns, d.inlet_ptr.p, d.bc_cell.p, d.bc_l53.p, d_Q, d_V, a.sV, d.hstill.p, d.mann.p, a.mann_d, a.sP,
        d.Qin.p, Qin_d, d.j_coef.p, d.j_coef.p + (size_t)K * nI, d.err.p);
    ctx->launches++;
  }
  void* kargs[] = {(void*)&a};
  if ((int64_t)fh.n_tiles * K >= ((int64_t)1 << 31)) { ctx->err = "forward mode: too many directions for one launch"; return HG_ERR_ARG; }
  const cudaError_t le = cudaLaunchKernel(kk.fn, dim3((unsigned)(fh.n_tiles * K)), dim3((unsigned)kk.threads), kargs, (size_t)kk.smem, ctx->stream);
  if (le != cudaSuccess) { ctx->err = std::string("fused_jvp launch: ") + cudaGetErrorString(le); return HG_ERR_CUDA; }
  ctx->launches++;
  return HG_OK;
}

}  // namespace hg
