// Hand-written adjoint (vector-Jacobian product) of the fused RHS -- replaces Zygote on swe_2d_rhs for the
// inversion / UDE loss gradients (call sites: swe_2D_inversion.jl:339, swe_2D_UDE.jl:666; contract
// debug_AD.jl:60,75:  back(lambda) -> (Qbar, pbar)).
//
// Same tile structure and TMA staging as hg_fused.cu; forward intermediates are RECOMPUTED, not stored:
//   phase 1  cells + halo -> shared memory: clamped state, derived values (u, v, s = sqrt(h+eps)), the factors of the
//            derived map's transpose (s/h, 1/(2s), dP/dxi) and mu = lambda / area;
//   phase 2  every face once: Fbar = len * (mu_R - mu_L); reverse-mode sweep through Riemann_2D_Roe
//            (swe_2D_solvers.jl:4-164; branch predicates, clamps and wet flags are constants exactly as in
//            Zygote / ForwardDiff), composed with the transpose of the derived map and of the dry clamp, so that each
//            face parks only the adjoints of the raw state (xi, q_x, q_y) of its two sides in shared memory; two
//            faces per thread and trip where the launch shape has the registers (FPT = 2).  Boundary faces
//            additionally pull the ghost adjoint back through the boundary condition (bc_2D.jl:640-834) onto their
//            internal cell;
//   phase 3  each owned cell GATHERS the L- or R-side adjoints of its faces through the same slots used by the
//            forward pass (the adjoint of a gather done as a gather: no atomics, deterministic) and adds the adjoint
//            of the bed-slope and Manning-friction sources.
// Boundary-wide couplings (inlet conveyance split) and parameter reductions run in small follow-up kernels
// with fixed reduction trees.
#include "hg_device.cuh"

namespace hg {
namespace {
using namespace dev;

struct VjpArgs {
  int32_t N, n_tiles, want_s0;
  int32_t prefetch;            // > 0: CTA b pulls the blocks of tile b + prefetch into L2 (see hg_fused.cu, prefetch_work)
  const int32_t* tile_order;   // run the tiles tile_order[tile_base .. tile_base + n_run) (NULL: identity)
  int32_t tile_base, n_run;
  int64_t Ns;
  Consts c;
  const int32_t *tile_desc, *halo, *bface_e;
  const uint32_t* face_lr;
  const uint16_t* cf_idx;
  const double *face_nx, *face_ny, *face_len;
  const double *area, *hstill, *zb, *S0x, *S0y, *mann;
  const int32_t *bc_type, *bc_group;
  const double *bc_nx, *bc_ny, *bc_l23, *bc_hstill, *bc_zb, *inlet_coef, *wse;
  const int32_t *halo_off, *halo_cnt;
  const double* halo_recv;
  const double *Q, *lam;
  double *Qbar, *nbar, *s0bar;          // [3Ns], [Ns], [2Ns]
  double *ent_c, *ent_n, *ent_z;        // per boundary entry: inlet coef adjoint share, n adjoint, zb adjoint
  CommWait cw;                          // library-owned halo exchange: CTAs >= cw.from wait for the neighbours' pushes
};

struct Adj {
  double xi, h, u, v, s, P;   // adjoints of the staged per-cell variables (hu = h*u, hv = h*v folded in)
};

__device__ __forceinline__ void sabs_r(double x, double& A, double& r) {  // A = sqrt(x^2+eps), r = 1/A
  const double y = fma(x, x, EPS);
  r = fast_rsqrt(y);
  A = y * r;                                                              // same rounding as dev::smooth_abs
}

// Reverse sweep of the face flux, main (both sides wet) branch only: straight-line code.
// (f0b,f1b,f2b) = adjoint of the returned flux.  Outputs aL, aR = adjoints of (xi, h, u, v, s, P) per side.
__device__ __forceinline__ void roe_adj_wet(const Side& L, const Side& R, double nx, double ny, double g, double f0b,
                                            double f1b, double f2b, Adj& aL, Adj& aR) {
  // ---- forward recompute (same statements as dev::roe_flux)
  const double hRoe = 0.5 * (L.h + R.h);
  const double rs = fast_rcp(L.s + R.s);
  const double a_ = L.s * L.u + R.s * R.u, b_ = L.s * L.v + R.s * R.v;
  const double uRoe = a_ * rs, vRoe = b_ * rs;
  const double un = uRoe * nx + vRoe * ny;
  const double c2 = fma(g, hRoe, EPS);
  const double rc = fast_rsqrt(c2);
  const double c = c2 * rc, k = 0.5 * rc;
  const double d1 = R.xi - L.xi, d2 = R.hu - L.hu, d3 = R.hv - L.hv;
  const double t1 = uRoe * ny - vRoe * nx;
  const double w1 = -t1 * d1 + ny * d2 - nx * d3;
  const double e_ = un * d1 - (nx * d2 + ny * d3);
  const double m = k * e_;
  const double w2 = 0.5 * d1 + m, w3 = 0.5 * d1 - m;
  const double l1 = un, l2 = un - c, l3 = un + c;
  double A1, A2, A3, r1, r2, r3;
  sabs_r(l1, A1, r1); sabs_r(l2, A2, r2); sabs_r(l3, A3, r3);
  const double z2 = A2 * w2, z3 = A3 * w3;
  const double zs = z2 + z3;
  const double unL = L.u * nx + L.v * ny, unR = R.u * nx + R.v * ny;
  // ---- reverse
  const double b0 = 0.5 * f0b, b1 = 0.5 * f1b, b2 = 0.5 * f2b;
  const double y1b = -b0, y2b = -b1, y3b = -b2;
  double huLb = b0 * nx + b1 * unL, hvLb = b0 * ny + b2 * unL;
  double huRb = b0 * nx + b1 * unR, hvRb = b0 * ny + b2 * unR;
  const double unLb = b1 * L.hu + b2 * L.hv, unRb = b1 * R.hu + b2 * R.hv;
  const double Pb = b1 * nx + b2 * ny;
  double uLb = unLb * nx, vLb = unLb * ny, uRb = unRb * nx, vRb = unRb * ny;
  const double z1b = ny * y2b - nx * y3b;
  const double zsb = y1b + uRoe * y2b + vRoe * y3b;
  const double zdb = nx * y2b + ny * y3b;
  double uRoeb = zs * y2b, vRoeb = zs * y3b;
  double cb = zdb * (z3 - z2);
  const double z3b = zsb + c * zdb, z2b = zsb - c * zdb;
  const double A1b = z1b * w1, w1b = z1b * A1;
  const double A2b = z2b * w2, w2b = z2b * A2;
  const double A3b = z3b * w3, w3b = z3b * A3;
  const double l1b = A1b * l1 * r1, l2b = A2b * l2 * r2, l3b = A3b * l3 * r3;   // d sqrt(x^2+eps)/dx = x / sqrt(..)
  double unb = l1b + l2b + l3b;
  cb += l3b - l2b;
  double d1b = 0.5 * (w2b + w3b);
  const double mb = w2b - w3b;
  const double kb = mb * e_, eb = mb * k;
  unb += eb * d1;
  d1b += eb * un;
  double d2b = -eb * nx, d3b = -eb * ny;
  const double t1b = -w1b * d1;
  d1b -= w1b * t1;
  d2b += w1b * ny;
  d3b -= w1b * nx;
  uRoeb += t1b * ny;
  vRoeb -= t1b * nx;
  huRb += d2b; huLb -= d2b; hvRb += d3b; hvLb -= d3b;
  const double c2b = cb * k - 2.0 * kb * k * k * k;             // c = sqrt(c2), k = 1/(2 sqrt(c2))
  const double hRoeb = g * c2b;
  uRoeb += unb * nx;
  vRoeb += unb * ny;
  const double ab = uRoeb * rs, bb = vRoeb * rs;
  const double Sb = -(uRoeb * a_ + vRoeb * b_) * rs * rs;
  const double sLb = ab * L.u + bb * L.v + Sb, sRb = ab * R.u + bb * R.v + Sb;
  uLb += ab * L.s; vLb += bb * L.s; uRb += ab * R.s; vRb += bb * R.s;
  // ---- fold hu = h*u, hv = h*v into (h, u, v)
  aL.xi = -d1b; aL.P = Pb; aL.s = sLb;
  aL.u = uLb + huLb * L.h; aL.v = vLb + hvLb * L.h; aL.h = 0.5 * hRoeb + huLb * L.u + hvLb * L.v;
  aR.xi = d1b; aR.P = Pb; aR.s = sRb;
  aR.u = uRb + huRb * R.h; aR.v = vRb + hvRb * R.h; aR.h = 0.5 * hRoeb + huRb * R.u + hvRb * R.v;
}

// General face: the five wet/dry branches of Riemann_2D_Roe (swe_2D_solvers.jl:16-76) around the main sweep.
// Rare (dry fronts, boundary faces); the common path calls roe_adj_wet directly.
__device__ __forceinline__ void roe_flux_adj(Side L, Side R, double zbL, double zbR, double nx, double ny,
                                          double g, double hmin, double f0b, double f1b, double f2b, Adj* paL, Adj* paR) {
  Adj aL = Adj{0, 0, 0, 0, 0, 0}, aR = Adj{0, 0, 0, 0, 0, 0};
  const bool dryL = L.h <= hmin, dryR = R.h <= hmin;
  int mirror = 0;  // 1: R is the mirror image of L, 2: L is the mirror image of R
  if (dryL || dryR) {
    if (dryL && dryR) { *paL = aL; *paR = aR; return; }        // zero flux
    if ((L.h + zbL) < (zbR + hmin) && dryR) {
      mirror = 1;
      R.h = L.h; R.hu = -L.hu; R.hv = -L.hv; R.u = -L.u; R.v = -L.v; R.s = L.s;
    } else if ((R.h + zbR) < (zbL + hmin) && dryL) {
      mirror = 2;
      L.h = R.h; L.hu = -R.hu; L.hv = -R.hv; L.u = -R.u; L.v = -R.v; L.s = R.s;
    } else {
      // one-sided physical flux of the wet side W: o0 = hu nx + hv ny, o1 = hu un + p nx, o2 = hv un + p ny
      const Side& W = dryL ? R : L;
      Adj& aW = dryL ? aR : aL;
      const double un = W.u * nx + W.v * ny;
      const double hub = f0b * nx + f1b * un, hvb = f0b * ny + f2b * un;
      const double unb = f1b * W.hu + f2b * W.hv;
      const double pb = f1b * nx + f2b * ny;
      aW.u = unb * nx + hub * W.h;
      aW.v = unb * ny + hvb * W.h;
      aW.h = pb * g * (W.h + EPS) + hub * W.u + hvb * W.v;
      *paL = aL; *paR = aR;
      return;
    }
  }
  roe_adj_wet(L, R, nx, ny, g, f0b, f1b, f2b, aL, aR);
  if (mirror == 1) {         // R' = (L.h, -L.u, -L.v, L.s); xi_R, P_R stay R's own (swe_2D_solvers.jl:34-36)
    aL.h += aR.h; aL.u -= aR.u; aL.v -= aR.v; aL.s += aR.s;
    aR.h = 0.0; aR.u = 0.0; aR.v = 0.0; aR.s = 0.0;
  } else if (mirror == 2) {  // :49-51
    aR.h += aL.h; aR.u -= aL.u; aR.v -= aL.v; aR.s += aL.s;
    aL.h = 0.0; aL.u = 0.0; aL.v = 0.0; aL.s = 0.0;
  }
  *paL = aL; *paR = aR;
}

// ---- the common case, both sides wet: reverse sweep in two parts so that few values stay live.
// roe_adj_core: forward recompute + reverse through the Roe averages / eigen-decomposition, down to the adjoints of
// the face-level combinations (d1, d2, d3 = jumps of xi, hu, hv; hRoe; a = sL uL + sR uR, b likewise; S = sL + sR).
struct FaceCore {
  double b0, b1, b2;            // 0.5 * adjoint of the flux
  double d1b, d2b, d3b;         // adjoints of the jumps
  double hh;                    // 0.5 * adjoint of hRoe  (= adjoint of each side's h through hRoe)
  double ab, bb, Sb;            // adjoints of a, b, S
};
__device__ __forceinline__ void roe_adj_core(double xiL, double hL, double uL, double vL, double sL, double xiR, double hR,
                                             double uR, double vR, double sR, double nx, double ny, double g, double b0,
                                             double b1, double b2, FaceCore& o) {
  // (b0, b1, b2) = 0.5 * adjoint of the flux = (mu_R - mu_L) * (len/2), formed by the caller
  const double gh = 0.5 * g;
  const double rs = fast_rcp(sL + sR);
  const double a_ = fma(sL, uL, sR * uR), b_ = fma(sL, vL, sR * vR);
  const double uRoe = a_ * rs, vRoe = b_ * rs;
  const double un = fma(uRoe, nx, vRoe * ny);
  const double c2 = fma(gh, hL + hR, EPS);                        // g hRoe + eps
  const double rc = fast_rsqrt(c2);
  const double c = c2 * rc, k = 0.5 * rc;
  const double d1 = xiR - xiL, d2 = fma(hR, uR, -(hL * uL)), d3 = fma(hR, vR, -(hL * vL));
  const double t1 = fma(uRoe, ny, -(vRoe * nx));
  const double nd = fma(nx, d2, ny * d3);
  const double w1 = fma(ny, d2, -fma(nx, d3, t1 * d1));
  const double e_ = fma(un, d1, -nd);
  const double hd1 = 0.5 * d1, m = k * e_;
  const double w2 = hd1 + m, w3 = hd1 - m;
  const double l2 = un - c, l3 = un + c;
  double A1, A2, A3, r1, r2, r3;
  sabs_r(un, A1, r1); sabs_r(l2, A2, r2); sabs_r(l3, A3, r3);
  const double z2 = A2 * w2, z3 = A3 * w3;
  const double zs = z2 + z3;
  // ---- reverse
  const double z1b = fma(nx, b2, -(ny * b1));                 // y = R_mat z, ybar = -b
  const double zsb = -fma(vRoe, b2, fma(uRoe, b1, b0));
  const double zdb = -fma(nx, b1, ny * b2);
  const double czd = c * zdb;
  const double z3b = zsb + czd, z2b = zsb - czd;
  const double q1 = z1b * r1, q2 = z2b * r2, q3 = z3b * r3;   // d sqrt(x^2+eps)/dx = x / sqrt(..)
  const double w1b = z1b * A1, w2b = z2b * A2, w3b = z3b * A3;
  const double l1b = q1 * w1 * un, l2b = q2 * w2 * l2, l3b = q3 * w3 * l3;
  const double mb = w2b - w3b;
  const double kb = mb * e_, eb = mb * k;
  const double cb = fma(zdb, z3 - z2, l3b - l2b);
  const double unb = fma(eb, d1, l1b + l2b + l3b);
  const double t1b = -(w1b * d1);
  const double uRoeb = fma(unb, nx, fma(t1b, ny, -(zs * b1)));
  const double vRoeb = fma(unb, ny, -fma(t1b, nx, zs * b2));
  const double c2b = k * fma(-2.0 * kb, k * k, cb);            // c = sqrt(c2), k = 1/(2 sqrt(c2))
  o.b0 = b0; o.b1 = b1; o.b2 = b2;
  o.d1b = fma(0.5, w2b + w3b, fma(eb, un, -(w1b * t1)));
  o.d2b = fma(w1b, ny, -(eb * nx));
  o.d3b = -fma(w1b, nx, eb * ny);
  o.hh = gh * c2b;
  o.ab = uRoeb * rs; o.bb = vRoeb * rs;
  o.Sb = -fma(o.ab, uRoe, o.bb * vRoe);                        // -(uRoeb a + vRoeb b) rs^2 with a rs = uRoe
}
// One side's (xi, q_x, q_y) adjoints from the core: the per-side part of the sweep (physical flux, jumps, Roe
// weights) composed with the transpose of u = hu/h, v = hv/h, s = sqrt(h+eps), P(xi).  sgn = -1 for L, +1 for R.
// With sigma = s/h and bu = b1 u + b2 v:
//   qxb = [b0 nx + b1 un + sgn d2b] + bu nx + ab sigma          (likewise qyb)
//   xb  = Pb dP + sgn d1b + (ab u + bb v)(1/(2s) - sigma) + Sb/(2s) + hh - bu un
__device__ __forceinline__ void fold_core_side(const FaceCore& k, double sgn, double nx, double ny, double u, double v,
                                               double sg, double rs2, double dP, double& xb, double& qxb, double& qyb) {
  const double un = u * nx + v * ny;
  const double bu = k.b1 * u + k.b2 * v;
  const double abu = k.ab * u + k.bb * v;
  qxb = fma(k.ab, sg, fma(bu + k.b0, nx, fma(k.b1, un, sgn * k.d2b)));
  qyb = fma(k.bb, sg, fma(bu + k.b0, ny, fma(k.b2, un, sgn * k.d3b)));
  const double Pb = k.b1 * nx + k.b2 * ny;
  xb = fma(Pb, dP, sgn * k.d1b) + (fma(abu, rs2 - sg, fma(k.Sb, rs2, k.hh)) - bu * un);
}

template <int T, int ML, int MF, int NF>
struct __align__(16) VjpSmem {
  uint64_t bar[2];
  double xi[ML], h[ML], u[ML], v[ML], s[ML];              // staged forward state (clamped, derived); the pressure is not needed
  double sg[ML], rs2[ML], dP[ML];                         // s/h, 1/(2 s), dP/dxi = g (xi + eps + hstill): the derived map's transpose
  double m0[ML], m1[ML], m2[ML];                          // mu = lambda / area (lambda itself on arrival)
  double o[6][MF];                                        // rows 0..2: nx, ny, len on arrival; then (xi, q_x, q_y) adjoints, L side | R side
  double area[T], mann[T], sx[T], sy[T];
  uint32_t lr[MF];
  uint16_t cf[T * NF];
};

// adjoints of one side's staged variables (xi, h, u, v, s, P) -> adjoints of its raw state (xi, q_x, q_y): the
// transpose of the derived map u = hu/h, v = hv/h, s = sqrt(h+eps), P(xi) and of the dry clamp
// (semi_discretize_swe_2D.jl:104-106: a clamped cell's h, q are constants; xi itself is never clamped).
__device__ __forceinline__ void fold_side(const Adj& a, double h, double u, double v, double rh, double rs2, double dP,
                                          double hs, double& xb, double& qxb, double& qyb) {
  const bool wet = h > hs;                                 // clamped h == h_small  <=>  the cell was clamped
  const double ht = fma(a.s, rs2, fma(-(a.u * u + a.v * v), rh, a.h));
  const double x0 = fma(a.P, dP, a.xi);
  xb = wet ? x0 + ht : x0;
  qxb = wet ? a.u * rh : 0.0;
  qyb = wet ? a.v * rh : 0.0;
}

// ---- the rare faces
// interior face on a wet/dry front (exactly one dry side)
template <class Smem>
__device__ __forceinline__ void vjp_front_face(Smem& sm, const VjpArgs& a, int32_t f, double zbL, double zbR) {
    const uint32_t lr = sm.lr[f];
    const int32_t lL = lr & 0xFFFFu, lR = lr >> 16;
    const double hL = sm.h[lL], hR = sm.h[lR];
    const double g = a.c.g, hs = a.c.h_small;
    const double nx = sm.o[0][f], ny = sm.o[1][f], len = sm.o[2][f];
    const double f0b = (sm.m0[lR] - sm.m0[lL]) * len, f1b = (sm.m1[lR] - sm.m1[lL]) * len, f2b = (sm.m2[lR] - sm.m2[lL]) * len;
    Side L, R;
    L.xi = sm.xi[lL]; L.h = hL; L.u = sm.u[lL]; L.v = sm.v[lL]; L.s = sm.s[lL];
    L.hu = __dmul_rn(L.h, L.u); L.hv = __dmul_rn(L.h, L.v);
    R.xi = sm.xi[lR]; R.h = hR; R.u = sm.u[lR]; R.v = sm.v[lR]; R.s = sm.s[lR];
    R.hu = __dmul_rn(R.h, R.u); R.hv = __dmul_rn(R.h, R.v);
    Adj aL, aR;
    roe_flux_adj(L, R, zbL, zbR, nx, ny, g, hs, f0b, f1b, f2b, &aL, &aR);
    double xb, qxb, qyb;
    fold_side(aL, L.h, L.u, L.v, fast_rcp(L.h), sm.rs2[lL], sm.dP[lL], hs, xb, qxb, qyb);
    sm.o[0][f] = xb; sm.o[1][f] = qxb; sm.o[2][f] = qyb;
    fold_side(aR, R.h, R.u, R.v, fast_rcp(R.h), sm.rs2[lR], sm.dP[lR], hs, xb, qxb, qyb);
    sm.o[3][f] = xb; sm.o[4][f] = qxb; sm.o[5][f] = qyb;
}
// boundary face (physical boundary or halo face): rebuild the ghost state, sweep, pull back onto the owned cell.
// Its sweep is a second copy of the code: bits may differ from an interior face's in the last place, so a halo
// face is reproduced across rank counts to ~1e-15, not bit for bit (the RHS kernel is bit-identical).
template <class Smem>
__device__ __forceinline__ void vjp_boundary_face(Smem& sm, const VjpArgs& a, int32_t f, int32_t nint, int32_t bfp, int32_t c0) {
    const double g = a.c.g, hs = a.c.h_small;
    const uint32_t lr = sm.lr[f];
    const int32_t lL = lr & 0xFFFFu;
    double nx = sm.o[0][f], ny = sm.o[1][f];
    const double len = sm.o[2][f];
    Side L, R;
    L.xi = sm.xi[lL]; L.h = sm.h[lL]; L.u = sm.u[lL]; L.v = sm.v[lL]; L.s = sm.s[lL];
    L.hu = __dmul_rn(L.h, L.u); L.hv = __dmul_rn(L.h, L.v);   // never contracted into the flux FMAs
    double f0b = -sm.m0[lL], f1b = -sm.m1[lL], f2b = -sm.m2[lL];
    double zbl = a.zb[c0 + lL];                                // boundary faces always touch an owned cell
    // boundary face: rebuild the ghost state (bc_2D.jl:640-834)
    const int32_t e = __ldg(a.bface_e + bfp + (f - nint));
    const int32_t ty = a.bc_type[e];
    const int32_t kgrp = a.bc_group[e];
    const double bnx = a.bc_nx[e], bny = a.bc_ny[e];
    const double hstg = a.bc_hstill[e];
    double zbr = a.bc_zb[e];
    double vn = 0.0, wet = 0.0, mannc = 1.0;
    bool flip = false, exit_free = false;
    if (ty == BC_INLETQ) {
      wet = L.h > hs ? 1.0 : 0.0;
      mannc = sm.mann[lL];
      pdl_wait();                                // coef comes from k_inlet_coef, which may still be running (PDL)
      vn = __ldcg(a.inlet_coef + kgrp) * a.bc_l23[e] / mannc;
      R.h = L.h; R.hu = -L.h * vn * bnx * wet; R.hv = -L.h * vn * bny * wet;
    } else if (ty == BC_EXITH) {
      const double hg = a.wse[kgrp] - zbl;
      exit_free = hg > hs;                       // max(h_small, .) passes the derivative only when not clamped
      R.h = exit_free ? hg : hs; R.hu = L.hu; R.hv = L.hv;
    } else if (ty == BC_WALL) {
      R.h = L.h; R.hu = -L.hu; R.hv = -L.hv;
    } else if (ty == BC_SYMM) {
      const double vdn = L.hu * bnx + L.hv * bny;
      R.h = L.h; R.hu = L.hu - 2.0 * vdn * bnx; R.hv = L.hv - 2.0 * vdn * bny;
    } else {
      // remote cell: state and cotangent arrived through the halo buffers; its own adjoint is computed by its owner
      const int32_t off = a.halo_off[e], n = a.halo_cnt[e];
      const double xr = __ldcg(a.halo_recv + off), qxr = __ldcg(a.halo_recv + off + n), qyr = __ldcg(a.halo_recv + off + 2 * n);
      const double rAr = fast_rcp(a.bc_l23[e]);          // remote cell area rides in l23
      // mu_remote is a rounded product exactly like an in-tile cell's (no FMA contraction with the sum)
      f0b += __dmul_rn(__ldcg(a.halo_recv + off + 3 * n), rAr); f1b += __dmul_rn(__ldcg(a.halo_recv + off + 4 * n), rAr);
      f2b += __dmul_rn(__ldcg(a.halo_recv + off + 5 * n), rAr);
      const double hr = xr + hstg;
      const bool dryr = hr <= hs;
      R.h = dryr ? hs : hr; R.hu = dryr ? 0.0 : qxr; R.hv = dryr ? 0.0 : qyr; R.xi = xr;
      flip = kgrp != 0;
    }
    if (ty != BC_HALO) R.xi = R.h - hstg;
    derive(R, hstg, g);
    if (ty == BC_HALO) { R.hu = __dmul_rn(R.h, R.u); R.hv = __dmul_rn(R.h, R.v); }
    f0b *= len; f1b *= len; f2b *= len;
    if (flip) {   // evaluated with the remote cell as L (see hg_fused.cu): flux_out = -roe(R, L, -n)
      const Side tmp = L; L = R; R = tmp;
      const double tz = zbl; zbl = zbr; zbr = tz;
      nx = -nx; ny = -ny; f0b = -f0b; f1b = -f1b; f2b = -f2b;
    }
    Adj aL, aR;
    roe_flux_adj(L, R, zbl, zbr, nx, ny, g, hs, f0b, f1b, f2b, &aL, &aR);
    if (flip) { const Adj ta = aL; aL = aR; aR = ta; const Side tmp = L; L = R; R = tmp; }
    double ec = 0.0, en = 0.0, ez = 0.0;
    if (ty != BC_HALO) {
      // ghost (xi, h, u, v, s, P) adjoints -> ghost primitives (h_g, hu_g, hv_g); xi_g = h_g - hstill_g
      const double rhg = fast_rcp(R.h);
      const double hub = aR.u * rhg, hvb = aR.v * rhg;
      const double xib = aR.xi + aR.P * g * (R.xi + EPS + hstg);
      const double hgb = aR.h - (aR.u * R.u + aR.v * R.v) * rhg + aR.s * 0.5 * fast_rcp(R.s) + xib;
      // boundary condition transposed: adjoints of the internal cell's clamped (h, hu, hv)
      double hcb = 0.0, hucb = 0.0, hvcb = 0.0;
      if (ty == BC_INLETQ) {
        const double G = -(hub * bnx + hvb * bny) * wet;   // d(hu_g, hv_g) = -n wet d(h_c vn)
        hcb = hgb + G * vn;
        en = -G * L.h * vn / mannc;                        // vn = coef L^(2/3) / n_c
        ec = G * L.h * a.bc_l23[e] / mannc;                // share of the adjoint of coef_k = Q_k / A_k
      } else if (ty == BC_EXITH) {
        hucb = hub; hvcb = hvb;
        ez = exit_free ? -hgb : 0.0;                       // h_g = WSE - zb_c
      } else if (ty == BC_WALL) {
        hcb = hgb; hucb = -hub; hvcb = -hvb;
      } else {
        const double dn = hub * bnx + hvb * bny;
        hcb = hgb; hucb = hub - 2.0 * dn * bnx; hvcb = hvb - 2.0 * dn * bny;
      }
      aL.u += hucb * L.h; aL.v += hvcb * L.h; aL.h += hcb + hucb * L.u + hvcb * L.v;
    }
    a.ent_c[e] = ec; a.ent_n[e] = en; a.ent_z[e] = ez;
    double xb, qxb, qyb;
    fold_side(aL, L.h, L.u, L.v, fast_rcp(L.h), sm.rs2[lL], sm.dP[lL], hs, xb, qxb, qyb);
    sm.o[0][f] = xb; sm.o[1][f] = qxb; sm.o[2][f] = qyb;
    sm.o[3][f] = 0.0; sm.o[4][f] = 0.0; sm.o[5][f] = 0.0;
}

// The common path needs ~90 registers; TH x MB is picked per tile size so that MB CTAs fit next to each other in
// shared memory and TH*MB*regs <= 64K (vjp_pick below).
#ifdef HG_PHASE_CLOCKS   // experiment builds only: where a CTA's lifetime goes (thread 0's clock at the phase boundaries)
__device__ unsigned long long g_phase_clk[8];
#define PHASE_MARK(i) do { if (tid == 0) { const long long c_ = clock64(); atomicAdd(&g_phase_clk[i], (unsigned long long)(c_ - pc_)); pc_ = c_; } } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif
template <int T, int ML, int MF, int NF, int TH, int MB, int FPT>
__global__ void __launch_bounds__(TH, MB) k_fused_vjp(const __grid_constant__ VjpArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using Smem = VjpSmem<T, ML, MF, NF>;
  Smem& sm = *reinterpret_cast<Smem*>(smraw);
  constexpr int kThreads = TH;

  const int tid = threadIdx.x;
#ifdef HG_PHASE_CLOCKS
  long long pc_ = clock64();
#endif
  const int t = a.tile_order ? __ldg(a.tile_order + a.tile_base + (int)blockIdx.x) : (int)blockIdx.x;
  const int4 d0 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc));
  const int4 d1 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc) + 1);
  const int4 d2 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)t * kTileDesc) + 2);
  // Tiles are runs of T cells (build_tiles_T), so the cell range needs no descriptor: the per-cell rows are copied while
  // the descriptor is still on its way from L2 (~600 cycles); only the face rows and the halo list wait for it.
  const int32_t c0 = t * T, nc = min(T, a.N - c0);
  const int32_t ncp = (nc + 1) & ~1;
  const double g = a.c.g, hs = a.c.h_small;
  const int64_t Ns = a.Ns;
  if (tid == 0) mbar_init(sm.bar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t cb = (uint32_t)ncp * 8u;
    mbar_expect_tx_only(sm.bar, 11u * cb + (uint32_t)(T * NF) * 2u);
    bulk_g2s(sm.xi, a.Q + c0, cb, sm.bar);
    bulk_g2s(sm.u, a.Q + Ns + c0, cb, sm.bar);             // raw q_x; u replaces it in place
    bulk_g2s(sm.v, a.Q + 2 * Ns + c0, cb, sm.bar);         // raw q_y
    bulk_g2s(sm.dP, a.hstill + c0, cb, sm.bar);            // raw hstill; dP replaces it in place
    bulk_g2s(sm.m0, a.lam + c0, cb, sm.bar);
    bulk_g2s(sm.m1, a.lam + Ns + c0, cb, sm.bar);
    bulk_g2s(sm.m2, a.lam + 2 * Ns + c0, cb, sm.bar);
    bulk_g2s(sm.area, a.area + c0, cb, sm.bar);
    bulk_g2s(sm.mann, a.mann + c0, cb, sm.bar);
    bulk_g2s(sm.sx, a.S0x + c0, cb, sm.bar);
    bulk_g2s(sm.sy, a.S0y + c0, cb, sm.bar);
    bulk_g2s(sm.cf, a.cf_idx + (size_t)t * (T * NF), (uint32_t)(T * NF) * 2u, sm.bar);
  }
  const int32_t hp = d0.z, nh = d0.w;
  const int32_t fp = d1.x, nf = d1.y, nfp = d1.z;
  const int32_t nint = d2.y, bfp = d2.z;
  // bed elevation of a tile-local cell: only the (rare) faces with a dry side and boundary faces read it
  auto zb_local = [&](int32_t l) { return a.zb[l < ncp ? c0 + l : __ldg(a.halo + hp + (l - ncp))]; };

  // first trip of the halo gathers, part 1: the cell id
  const int32_t hgi = tid < nh ? __ldg(a.halo + hp + tid) : -1;
  PHASE_MARK(0);   // cell-row copies issued, descriptor here
  // The bulk copies are all issued by one thread.  Measured (scripts/micro/bulk_issue.cu, profiles/round2_bulk_issue_micro.txt):
  // a cp.async.bulk costs its thread ~48 cycles on an idle SM and ~100 under this kernel's load, the copy engine serialises
  // them whoever issues (sixteen lanes of one instruction: 4800 cycles; four warps with four copies each: 1200 cycles per
  // warp), and the engine's L2 prefetches queue behind the same port.
  if (tid == 0) {
    const uint32_t fb = (uint32_t)nfp * 8u;
    mbar_expect_tx(sm.bar, 3u * fb + (uint32_t)nfp * 4u);   // the barrier's only arrival
    bulk_g2s(sm.o[0], a.face_nx + fp, fb, sm.bar);
    bulk_g2s(sm.o[1], a.face_ny + fp, fb, sm.bar);
    bulk_g2s(sm.o[2], a.face_len + fp, fb, sm.bar);
    bulk_g2s(sm.lr, a.face_lr + fp, (uint32_t)nfp * 4u, sm.bar);
  }
  PHASE_MARK(7);   // issue of the bulk copies (thread 0)
  // clamp + derived values + the factors of the derived map's transpose, for local cell l
  auto stage_cell = [&](int32_t l, double xi, double qx, double qy, double hst) {
    const double h0 = xi + hst;
    const bool dry = h0 <= hs;
    const double h = dry ? hs : h0;
    const double rh = fast_rcp(h);
    const double y = h + EPS;
    const double r = fast_rsqrt(y);
    double s = y * r;
    s = fma(fma(-s, s, y), 0.5 * r, s);
    const double xe = xi + EPS;
    sm.h[l] = h; sm.u[l] = dry ? 0.0 : qx * rh; sm.v[l] = dry ? 0.0 : qy * rh; sm.s[l] = s;
    sm.sg[l] = s * rh; sm.rs2[l] = 0.5 * r; sm.dP[l] = g * (xe + hst);
  };
  // halo cells (state + lambda / area), the only indirect reads: two dependent trips to L2 / HBM.  The values of the first
  // trip (nh <= kThreads on all but ragged tilings) are only LOADED here; they are consumed after the owned cells of phase 1,
  // whose bulk copies need one trip, so the gather latency passes behind that work instead of idling the CTA.
  double hx = 0.0, hqx = 0.0, hqy = 0.0, hhst = 0.0, hA = 1.0, hl0 = 0.0, hl1 = 0.0, hl2 = 0.0;
  if (hgi >= 0) {
    hx = a.Q[hgi]; hqx = a.Q[Ns + hgi]; hqy = a.Q[2 * Ns + hgi];
    hhst = a.hstill[hgi]; hA = a.area[hgi];
    hl0 = a.lam[hgi]; hl1 = a.lam[Ns + hgi]; hl2 = a.lam[2 * Ns + hgi];
  }
  auto stage_halo = [&](int32_t k, double xi, double qx, double qy, double hst, double A, double l0, double l1, double l2) {
    const double rA = fast_rcp(A);
    const int32_t l = ncp + k;
    sm.xi[l] = xi;
    stage_cell(l, xi, qx, qy, hst);
    sm.m0[l] = l0 * rA; sm.m1[l] = l1 * rA; sm.m2[l] = l2 * rA;
  };
  if (a.prefetch > 0 && tid == kThreads - 1 && (int)blockIdx.x + a.prefetch < a.n_run) {
    // L2 prefetch of the tile one residency ahead.  (Through the load/store path instead -- a warp of prefetch.global.L2 per
    // row -- the kernel takes 1.13 ms instead of 0.81: the issuing warp stalls ~4500 cycles and holds up its CTA.)
    const int32_t wp = (int)blockIdx.x + a.prefetch;
    const int32_t tp = a.tile_order ? __ldg(a.tile_order + a.tile_base + wp) : wp;
    const int4 p0 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)tp * kTileDesc));
    const int4 p1 = __ldg(reinterpret_cast<const int4*>(a.tile_desc + (size_t)tp * kTileDesc) + 1);
    const uint32_t cb = (uint32_t)((p0.y + 1) & ~1) * 8u, fb = (uint32_t)p1.z * 8u;
    const int32_t pc = p0.x, pf = p1.x;
    bulk_prefetch_l2(a.Q + pc, cb); bulk_prefetch_l2(a.Q + Ns + pc, cb); bulk_prefetch_l2(a.Q + 2 * Ns + pc, cb);
    bulk_prefetch_l2(a.lam + pc, cb); bulk_prefetch_l2(a.lam + Ns + pc, cb); bulk_prefetch_l2(a.lam + 2 * Ns + pc, cb);
    bulk_prefetch_l2(a.hstill + pc, cb); bulk_prefetch_l2(a.area + pc, cb); bulk_prefetch_l2(a.mann + pc, cb);
    bulk_prefetch_l2(a.S0x + pc, cb); bulk_prefetch_l2(a.S0y + pc, cb);
    bulk_prefetch_l2(a.face_nx + pf, fb); bulk_prefetch_l2(a.face_ny + pf, fb); bulk_prefetch_l2(a.face_len + pf, fb);
    bulk_prefetch_l2(a.face_lr + pf, (uint32_t)p1.z * 4u);
    bulk_prefetch_l2(a.cf_idx + (size_t)tp * (T * NF), (uint32_t)(T * NF) * 2u);
    if (p0.w > 0) bulk_prefetch_l2(a.halo + p0.z, (uint32_t)((p0.w + 3) & ~3) * 4u);
    if (!a.tile_order && wp + a.prefetch < a.n_run) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.tile_desc + (size_t)(wp + a.prefetch) * kTileDesc));
  }
  PHASE_MARK(1);   // bulk copies and halo gathers issued (+ prefetch)
  if (a.cw.n > 0 && (int)blockIdx.x >= a.cw.from && (int)blockIdx.x < a.cw.to) comm_wait(a.cw, tid);   // band tile: phase 2b reads what the neighbours push
  mbar_wait(sm.bar, 0);
  PHASE_MARK(2);   // wait for the bulk copies

  // ---- phase 1: owned cells, in place
  auto own_cell = [&](int32_t l) {
    stage_cell(l, sm.xi[l], sm.u[l], sm.v[l], sm.dP[l]);
    const double rA = fast_rcp(sm.area[l]);
    sm.m0[l] *= rA; sm.m1[l] *= rA; sm.m2[l] *= rA;
  };
  if constexpr (FPT == 1) {
    for (int32_t l = tid; l < nc; l += kThreads) own_cell(l);
  } else {   // two cells per trip in one basic block (independent chains interleave)
    for (int32_t l = tid; l < nc; l += 2 * kThreads) {
      if (l + kThreads < nc) {
        const int32_t l2 = l + kThreads;
        const double x1 = sm.xi[l], qx1 = sm.u[l], qy1 = sm.v[l], hs1 = sm.dP[l], A1 = sm.area[l];
        const double x2 = sm.xi[l2], qx2 = sm.u[l2], qy2 = sm.v[l2], hs2 = sm.dP[l2], A2 = sm.area[l2];
        stage_cell(l, x1, qx1, qy1, hs1);
        stage_cell(l2, x2, qx2, qy2, hs2);
        const double rA1 = fast_rcp(A1), rA2 = fast_rcp(A2);
        sm.m0[l] *= rA1; sm.m1[l] *= rA1; sm.m2[l] *= rA1;
        sm.m0[l2] *= rA2; sm.m1[l2] *= rA2; sm.m2[l2] *= rA2;
      } else {
        own_cell(l);
      }
    }
  }
  // the halo cells loaded above
  if (hgi >= 0) stage_halo(tid, hx, hqx, hqy, hhst, hA, hl0, hl1, hl2);
  for (int32_t k = tid + kThreads; k < nh; k += kThreads) {   // ragged tilings: further trips
    const int32_t gi = __ldg(a.halo + hp + k);
    stage_halo(k, a.Q[gi], a.Q[Ns + gi], a.Q[2 * Ns + gi], a.hstill[gi], a.area[gi], a.lam[gi], a.lam[Ns + gi], a.lam[2 * Ns + gi]);
  }
  __syncthreads();
  PHASE_MARK(3);   // phase 1 + halo cells

  // ---- phase 2a: interior faces, common cases only.  Both sides wet: straight-line sweep with the derived map's
  // transpose folded in algebraically (roe_adj_core + fold_core_side, ~90 registers); both dry: zero flux, zero
  // adjoint.  Faces with exactly one dry side (wet/dry fronts) take the general routine.  (Measured: moving the
  // general routine out of line, or into a second scan over the faces, is 5-10 % slower.)
  // one face, any case
  auto one_face = [&](int32_t f, int32_t lL, int32_t lR, double hL, double hR) {
    const bool dL = hL <= hs, dR = hR <= hs;
    if (__builtin_expect(dL || dR, 0)) {
      if (dL && dR) {
        sm.o[0][f] = 0.0; sm.o[1][f] = 0.0; sm.o[2][f] = 0.0; sm.o[3][f] = 0.0; sm.o[4][f] = 0.0; sm.o[5][f] = 0.0;
      } else {
        vjp_front_face(sm, a, f, zb_local(lL), zb_local(lR));
      }
      return;
    }
    const double nx = sm.o[0][f], ny = sm.o[1][f], hl = 0.5 * sm.o[2][f];
    const double f0b = (sm.m0[lR] - sm.m0[lL]) * hl, f1b = (sm.m1[lR] - sm.m1[lL]) * hl, f2b = (sm.m2[lR] - sm.m2[lL]) * hl;
    FaceCore k;
    roe_adj_core(sm.xi[lL], hL, sm.u[lL], sm.v[lL], sm.s[lL], sm.xi[lR], hR, sm.u[lR], sm.v[lR], sm.s[lR], nx, ny, g,
                 f0b, f1b, f2b, k);
    double xb, qxb, qyb;
    fold_core_side(k, -1.0, nx, ny, sm.u[lL], sm.v[lL], sm.sg[lL], sm.rs2[lL], sm.dP[lL], xb, qxb, qyb);
    sm.o[0][f] = xb; sm.o[1][f] = qxb; sm.o[2][f] = qyb;
    fold_core_side(k, 1.0, nx, ny, sm.u[lR], sm.v[lR], sm.sg[lR], sm.rs2[lR], sm.dP[lR], xb, qxb, qyb);
    sm.o[3][f] = xb; sm.o[4][f] = qxb; sm.o[5][f] = qyb;
  };
  if constexpr (FPT == 1) {
    for (int32_t f = tid; f < nint; f += kThreads) {
      const uint32_t lr = sm.lr[f];
      const int32_t lL = lr & 0xFFFFu, lR = lr >> 16;
      one_face(f, lL, lR, sm.h[lL], sm.h[lR]);
    }
  } else {
    // two faces per thread and trip: when both are in the common case their sweeps sit in one basic block, so the two
    // independent dependency chains interleave (the sweep is latency-bound at the occupancy its registers allow)
    for (int32_t fA = tid; fA < nint; fA += 2 * kThreads) {
      const int32_t fB = fA + kThreads;
      const uint32_t lrA = sm.lr[fA];
      const int32_t aL = lrA & 0xFFFFu, aR = lrA >> 16;
      const double hAL = sm.h[aL], hAR = sm.h[aR];
      if (fB >= nint) { one_face(fA, aL, aR, hAL, hAR); break; }
      const uint32_t lrB = sm.lr[fB];
      const int32_t bL = lrB & 0xFFFFu, bR = lrB >> 16;
      const double hBL = sm.h[bL], hBR = sm.h[bR];
      if (__builtin_expect(hAL <= hs || hAR <= hs || hBL <= hs || hBR <= hs, 0)) {
        one_face(fA, aL, aR, hAL, hAR);
        one_face(fB, bL, bR, hBL, hBR);
        continue;
      }
      const double nxA = sm.o[0][fA], nyA = sm.o[1][fA], lenA = 0.5 * sm.o[2][fA];   // half lengths: the 1/2 of the Roe average
      const double nxB = sm.o[0][fB], nyB = sm.o[1][fB], lenB = 0.5 * sm.o[2][fB];
      FaceCore kA, kB;
      roe_adj_core(sm.xi[aL], hAL, sm.u[aL], sm.v[aL], sm.s[aL], sm.xi[aR], hAR, sm.u[aR], sm.v[aR], sm.s[aR], nxA, nyA, g,
                   (sm.m0[aR] - sm.m0[aL]) * lenA, (sm.m1[aR] - sm.m1[aL]) * lenA, (sm.m2[aR] - sm.m2[aL]) * lenA, kA);
      roe_adj_core(sm.xi[bL], hBL, sm.u[bL], sm.v[bL], sm.s[bL], sm.xi[bR], hBR, sm.u[bR], sm.v[bR], sm.s[bR], nxB, nyB, g,
                   (sm.m0[bR] - sm.m0[bL]) * lenB, (sm.m1[bR] - sm.m1[bL]) * lenB, (sm.m2[bR] - sm.m2[bL]) * lenB, kB);
      double xA, qxA, qyA, xB, qxB, qyB;
      fold_core_side(kA, -1.0, nxA, nyA, sm.u[aL], sm.v[aL], sm.sg[aL], sm.rs2[aL], sm.dP[aL], xA, qxA, qyA);
      fold_core_side(kB, -1.0, nxB, nyB, sm.u[bL], sm.v[bL], sm.sg[bL], sm.rs2[bL], sm.dP[bL], xB, qxB, qyB);
      sm.o[0][fA] = xA; sm.o[1][fA] = qxA; sm.o[2][fA] = qyA;
      sm.o[0][fB] = xB; sm.o[1][fB] = qxB; sm.o[2][fB] = qyB;
      fold_core_side(kA, 1.0, nxA, nyA, sm.u[aR], sm.v[aR], sm.sg[aR], sm.rs2[aR], sm.dP[aR], xA, qxA, qyA);
      fold_core_side(kB, 1.0, nxB, nyB, sm.u[bR], sm.v[bR], sm.sg[bR], sm.rs2[bR], sm.dP[bR], xB, qxB, qyB);
      sm.o[3][fA] = xA; sm.o[4][fA] = qxA; sm.o[5][fA] = qyA;
      sm.o[3][fB] = xB; sm.o[4][fB] = qxB; sm.o[5][fB] = qyB;
    }
  }
  PHASE_MARK(4);   // phase 2a (thread 0's own faces)
  // ---- phase 2b: boundary faces (physical boundaries and halo faces)
  for (int32_t f = nint + tid; f < nf; f += kThreads) vjp_boundary_face(sm, a, f, nint, bfp, c0);
  if (tid < 6) sm.o[tid][nfp] = 0.0;   // the zero slot of unused cf entries
  __syncthreads();
  PHASE_MARK(5);   // phase 2b + waiting for the slowest warp of phase 2

  // ---- phase 3: per-cell gather of the face adjoints (L or R side of each face) + source adjoint
  const double kfr = g / (a.c.k_n * a.c.k_n);
  struct CellOut { double xib, qxb, qyb, nb, s0x, s0y; };
  auto cell_adj = [&](int32_t l) {
    uint16_t slot[NF];
    if constexpr (NF == 4) {
      const uint2 w = *reinterpret_cast<const uint2*>(&sm.cf[l * 4]);
      slot[0] = (uint16_t)(w.x & 0xFFFFu); slot[1] = (uint16_t)(w.x >> 16);
      slot[2] = (uint16_t)(w.y & 0xFFFFu); slot[3] = (uint16_t)(w.y >> 16);
    } else {
      const uint4 w = *reinterpret_cast<const uint4*>(&sm.cf[l * 8]);
      slot[0] = (uint16_t)(w.x & 0xFFFFu); slot[1] = (uint16_t)(w.x >> 16);
      slot[2] = (uint16_t)(w.y & 0xFFFFu); slot[3] = (uint16_t)(w.y >> 16);
      slot[4] = (uint16_t)(w.z & 0xFFFFu); slot[5] = (uint16_t)(w.z >> 16);
      slot[6] = (uint16_t)(w.w & 0xFFFFu); slot[7] = (uint16_t)(w.w >> 16);
    }
    double xib = 0.0, qxb = 0.0, qyb = 0.0;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int32_t f = slot[j] & 0x7FFF;
      const int side = (slot[j] & 0x8000) ? 3 : 0;
      xib += sm.o[side + 0][f]; qxb += sm.o[side + 1][f]; qyb += sm.o[side + 2][f];
    }
    // sources (wet cells): r1 += g xi S0x - C m qx,  C = g n^2/k_n^2 (h+hs)^(-7/3),  m = sqrt(qx^2+qy^2+eps).
    // Evaluated branch-free (every operand is finite for a clamped cell too) and selected by the wet flag.
    const double h = sm.h[l];
    const bool wet = h > hs;
    const double xi = sm.xi[l], u = sm.u[l], v = sm.v[l];
    const double A = sm.area[l], n = sm.mann[l];
    const double lam1 = sm.m1[l] * A, lam2 = sm.m2[l] * A;   // mu * area = lambda
    const double qx = h * u, qy = h * v;
    const double y = fma(qx, qx, fma(qy, qy, EPS));
    const double rm = fast_rsqrt(y);
    const double mag = y * rm;
    const double w = rcbrt_pos(h + hs), w2 = w * w, w3 = w2 * w;   // (h+hs)^(-1/3): w^7 = (h+hs)^(-7/3), w^3 = 1/(h+hs)
    const double C = kfr * n * n * (w3 * w3 * w);
    const double fxb = -lam1, fyb = -lam2;
    const double Cm = C * mag, Cr = C * rm;
    const double lq = fma(fxb, qx, fyb * qy);             // fbar . q
    const double dqx = fma(fxb, Cm, Cr * qx * lq);        // fbar_x C m + C q_x (fbar . q) / m
    const double dqy = fma(fyb, Cm, Cr * qy * lq);
    const double D = lq * Cm;                              // fxb*fx + fyb*fy
    // -(7/3) D / (h+hs): through h = xi + hstill (a wet cell is unclamped); bed slope: g xi S0 . lambda
    const double dxi = fma(-(7.0 / 3.0) * D, w3, g * fma(sm.sx[l], lam1, sm.sy[l] * lam2));
    CellOut o;
    o.xib = wet ? xib + dxi : xib;
    o.qxb = wet ? qxb + dqx : qxb;
    o.qyb = wet ? qyb + dqy : qyb;
    o.nb = wet ? 2.0 * D * fast_rcp(n) : 0.0;
    o.s0x = wet ? g * xi * lam1 : 0.0;
    o.s0y = wet ? g * xi * lam2 : 0.0;
    return o;
  };
  auto cell_store = [&](int32_t l, const CellOut& o) {
    const int32_t gi = c0 + l;
    a.Qbar[gi] = o.xib;
    a.Qbar[Ns + gi] = o.qxb;
    a.Qbar[2 * Ns + gi] = o.qyb;
    a.nbar[gi] = o.nb;
    if (a.want_s0) { a.s0bar[gi] = o.s0x; a.s0bar[Ns + gi] = o.s0y; }
  };
  if constexpr (FPT == 1) {
    for (int32_t l = tid; l < nc; l += kThreads) cell_store(l, cell_adj(l));
  } else {
    for (int32_t l = tid; l < nc; l += 2 * kThreads) {
      if (l + kThreads < nc) {
        const CellOut o1 = cell_adj(l), o2 = cell_adj(l + kThreads);
        cell_store(l, o1); cell_store(l + kThreads, o2);
      } else {
        cell_store(l, cell_adj(l));
      }
    }
  }
  PHASE_MARK(6);   // phase 3
}

// ---------------------------------------------------------------- boundary-wide couplings of the inlet-q split
// coef_k = Q_k / A_k,  A_k = sum_e L^(5/3) h_c / n_c wet  (bc_2D.jl:674-691).  One CTA per inlet boundary:
// coefbar_k = sum_e ent_c[e] (fixed tree), Qbar_k = coefbar_k / A_k, Abar_k = -coefbar_k coef_k / A_k; then every
// wet entry receives  hbar_c += Abar L^(5/3)/n_c  and  nbar_c -= Abar L^(5/3) h_c / n_c^2  (stored per entry).
__global__ void __launch_bounds__(256) k_inlet_adj(Consts c, const int32_t* inlet_ptr, const int32_t* bc_cell,
                                                   const double* bc_l53, const double* Q, const double* hstill,
                                                   const double* mann, const double* Atot, const double* coef,
                                                   const double* ent_c, double* ent_h, double* ent_n, double* Qinbar) {
  __shared__ double red[256];
  __shared__ double sAbar;
  const int k = blockIdx.x;
  const int32_t e0 = inlet_ptr[k], e1 = inlet_ptr[k + 1];
  // the per-entry gathers (cell id -> depth, n) do not depend on the reduction: issue them first, so that their latency
  // (two dependent global accesses) overlaps the tree instead of following it
  constexpr int KM = 4;                          // entries per thread kept in registers (inlets of up to 1024 faces)
  double hh[KM], nn[KM], ll[KM];
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    const int32_t e = e0 + threadIdx.x + j * 256;
    hh[j] = 0.0; nn[j] = 1.0; ll[j] = 0.0;
    if (e < e1) {
      const int32_t ci = bc_cell[e];
      hh[j] = Q[ci] + hstill[ci]; nn[j] = mann[ci]; ll[j] = bc_l53[e];
      acc += ent_c[e];
    }
  }
  for (int32_t e = e0 + threadIdx.x + KM * 256; e < e1; e += 256) acc += ent_c[e];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cbar = red[0];
    const double rA = 1.0 / Atot[k];              // coef = Q / A
    Qinbar[k] = cbar * rA;                        // d coef / d Q = 1 / A
    sAbar = -cbar * coef[k] * rA;                 // d coef / d A = -Q / A^2
  }
  __syncthreads();
  const double Abar = sAbar;
#pragma unroll
  for (int j = 0; j < KM; ++j) {
    const int32_t e = e0 + threadIdx.x + j * 256;
    if (e < e1) {
      const bool wet = hh[j] > c.h_small;
      ent_h[e] = wet ? Abar * ll[j] / nn[j] : 0.0;
      ent_n[e] += wet ? -Abar * ll[j] * hh[j] / (nn[j] * nn[j]) : 0.0;
    }
  }
  for (int32_t e = e0 + threadIdx.x + KM * 256; e < e1; e += 256) {
    const int32_t ci = bc_cell[e];
    const double h = Q[ci] + hstill[ci];
    const bool wet = h > c.h_small;
    const double n = mann[ci];
    ent_h[e] = wet ? Abar * bc_l53[e] / n : 0.0;
    ent_n[e] += wet ? -Abar * bc_l53[e] * h / (n * n) : 0.0;
  }
}

// One thread per boundary-adjacent cell: add its entries' contributions in fixed order (deterministic).
__global__ void k_bc_scatter(int32_t nbcell, const int32_t* __restrict__ bcell, const int32_t* __restrict__ bcell_ptr,
                             const int32_t* __restrict__ bcell_ent, const int32_t* __restrict__ bc_type,
                             const double* __restrict__ ent_h, const double* __restrict__ ent_n, double* Qbar, double* nbar) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbcell) return;
  const int32_t c = bcell[i];
  double hb = 0.0, nb = 0.0;
  for (int32_t k = bcell_ptr[i]; k < bcell_ptr[i + 1]; ++k) {
    const int32_t e = bcell_ent[k];
    if (bc_type[e] == BC_INLETQ) { hb += ent_h[e]; nb += ent_n[e]; }
  }
  Qbar[c] += hb;   // inlet cells that receive this are wet, hence unclamped: xi_bar += h_bar
  nbar[c] += nb;
}

// pbar_z = sum_{cells of zone z} nbar  (process_ManningN_2D.jl:88 transposed): two fixed-shape stages
constexpr int kZoneBlock = 256, kZoneChunk = 4096;
__global__ void __launch_bounds__(kZoneBlock) k_zone_partial(int32_t N, int32_t n_mat, const int32_t* __restrict__ matid,
                                                             const double* __restrict__ nbar, double* __restrict__ part) {
  extern __shared__ double zs[];   // [kZoneBlock][n_mat] would be too big for many zones: loop zones instead
  const int32_t b0 = blockIdx.x * kZoneChunk;
  for (int32_t z = 0; z < n_mat; ++z) {
    double acc = 0.0;
    for (int32_t i = b0 + threadIdx.x; i < min(N, b0 + kZoneChunk); i += kZoneBlock)
      if (matid[i] == z) acc += nbar[i];
    zs[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kZoneBlock / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) zs[threadIdx.x] += zs[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * n_mat + z] = zs[0];
    __syncthreads();
  }
}
// (many zones: one CTA per zone over the chunk partials of k_zone_partial)
__global__ void __launch_bounds__(256) k_zone_final_many(int32_t nblocks, int32_t n_mat, const double* __restrict__ part, double* __restrict__ pbar) {
  __shared__ double red[256];
  const int32_t z = blockIdx.x;
  double acc = 0.0;
  for (int32_t b = threadIdx.x; b < nblocks; b += 256) acc += part[(size_t)b * n_mat + z];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) pbar[z] = red[0];
}
// Few zones (the reference's cases have 1-6): ONE launch straight from nbar + one-byte zone ids (9 B per cell, bandwidth-bound).
// A fixed grid of CTAs strides over the cells with four independent loads in flight per thread and one register
// accumulator per zone; per-CTA sums go to part2[cta][.]; the CTA that finishes last (ticket counter) adds part2 in a fixed
// shape (256 strided sums in CTA order, then a tree) and writes pbar.  Every order is fixed: bit-reproducible.
constexpr int kZoneFew = 8, kZoneGrid = 148 * 8;
template <int NZ>
__global__ void __launch_bounds__(256) k_zone_reduce(int32_t N, int32_t n_mat, const uint8_t* __restrict__ matid, const double* __restrict__ nbar,
                                                     double* __restrict__ part2, unsigned int* __restrict__ ticket, double* __restrict__ pbar) {
  __shared__ double ws[8][NZ];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[NZ];
#pragma unroll
  for (int z = 0; z < NZ; ++z) acc[z] = 0.0;
  const int32_t stride = gridDim.x * 256;
  for (int32_t i = blockIdx.x * 256 + threadIdx.x; i < N; i += 4 * stride) {
    int32_t m[4];
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int32_t j = i + k * stride;
      m[k] = j < N ? (int32_t)__ldg(matid + j) : -1;
      v[k] = j < N ? __ldg(nbar + j) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int z = 0; z < NZ; ++z) acc[z] += (m[k] == z) ? v[k] : 0.0;
  }
#pragma unroll
  for (int z = 0; z < NZ; ++z) {
    double v = acc[z];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) ws[warp][z] = v;
  }
  __syncthreads();
  if (threadIdx.x < NZ) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += ws[w][threadIdx.x];
    part2[(size_t)blockIdx.x * NZ + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA: every thread adds its strided share of the CTAs' sums for ALL zones (fixed order), then ONE shuffle +
  // shared-memory reduction for all zones at once
#pragma unroll
  for (int z = 0; z < NZ; ++z) acc[z] = 0.0;
  for (unsigned c = threadIdx.x; c < gridDim.x; c += 256) {
#pragma unroll
    for (int z = 0; z < NZ; ++z) acc[z] += __ldcg(part2 + (size_t)c * NZ + z);
  }
  __syncthreads();                      // ws is reused
#pragma unroll
  for (int z = 0; z < NZ; ++z) {
    double v = acc[z];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) ws[warp][z] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_mat) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += ws[w][threadIdx.x];
    pbar[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *ticket = 0;   // ready for the next launch
}

// zbar (reference order) = (update_bed_data)^T S0bar + exit-h entries.  Transposed Green-Gauss as a gather:
// cell i collects, per face, its own term and the neighbour's term through the shared face value
// zb_f = (zb_i + zb_nb)/2 (interior) or zb_i (boundary)  (fvm_schemes_2D.jl:89-105, 133-167).
__global__ void k_zb_bar(int32_t N, int64_t Ns, const int32_t* __restrict__ iperm, const int32_t* __restrict__ cf_ptr,
                         const int32_t* __restrict__ cf_nb, const int32_t* __restrict__ cf_rev,
                         const double* __restrict__ cf_nx, const double* __restrict__ cf_ny,
                         const double* __restrict__ cf_len, const double* __restrict__ area_ref,
                         const double* __restrict__ s0bar, double* __restrict__ zbar) {
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int32_t i = iperm[r];
  const double gx = -s0bar[i] / area_ref[r], gy = -s0bar[Ns + i] / area_ref[r];   // S0 = -1 * (sum / A)
  double acc = 0.0;
  for (int32_t k = cf_ptr[r]; k < cf_ptr[r + 1]; ++k) {
    const int32_t nb = cf_nb[k];
    const double own = (cf_nx[k] * gx + cf_ny[k] * gy) * cf_len[k];
    if (nb >= N) { acc += own; continue; }
    const int32_t j = iperm[nb], kk = cf_rev[k];
    const double hx = -s0bar[j] / area_ref[nb], hy = -s0bar[Ns + j] / area_ref[nb];
    acc += 0.5 * (own + (cf_nx[kk] * hx + cf_ny[kk] * hy) * cf_len[kk]);
  }
  zbar[r] = acc;
}
__global__ void k_zb_bar_exit(int32_t nbcell, const int32_t* __restrict__ bcell_ref, const int32_t* __restrict__ bcell_ptr,
                              const int32_t* __restrict__ bcell_ent, const int32_t* __restrict__ bc_type,
                              const double* __restrict__ ent_z, double* __restrict__ zbar) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbcell) return;
  double acc = 0.0;
  for (int32_t k = bcell_ptr[i]; k < bcell_ptr[i + 1]; ++k) {
    const int32_t e = bcell_ent[k];
    if (bc_type[e] == BC_EXITH) acc += ent_z[e];
  }
  zbar[bcell_ref[i]] += acc;
}

__global__ void k_gather1(int32_t N, const int32_t* __restrict__ map, const double* __restrict__ src, double* __restrict__ dst) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) dst[i] = src[map[i]];
}

// ---- launch shapes.  (threads, CTAs/SM) per tile size: CTAs/SM is what the shared-memory footprint allows,
// threads keeps threads*CTAs*registers <= 64K.  Variants 1, 2 (hg_options.reserved[2]) exist for tuning sweeps.
struct VjpKernel {
  const void* fn = nullptr;
  int threads = 0, smem = 0, ctas_per_sm = 1;
};
template <int T, int ML, int MF, int NF, int TH, int MB, int FPT>
VjpKernel vjp_mk() {
  return VjpKernel{(const void*)k_fused_vjp<T, ML, MF, NF, TH, MB, FPT>, TH, (int)sizeof(VjpSmem<T, ML, MF, NF>), MB};
}
template <int T, int ML, int MF, int NF>
VjpKernel vjp_pick(int v) {
  if constexpr (T == 256) {
    // measured at 16M cells (ms): 128x3 two faces per trip 0.92 | 192x3 one face 1.02 | 160x3 two faces 1.14 (spills)
    if (v == 1) return vjp_mk<T, ML, MF, NF, 192, 3, 1>();
    if (v == 2) return vjp_mk<T, ML, MF, NF, 160, 3, 2>();
    return vjp_mk<T, ML, MF, NF, 128, 3, 2>();
  } else if constexpr (T == 384) {
    // fewer, larger tiles (measured at 16M cells: 0.89 ms vs 0.82 ms for T = 256 -- two CTAs per SM hide less latency)
    if (v == 1) return vjp_mk<T, ML, MF, NF, 256, 2, 1>();
    return vjp_mk<T, ML, MF, NF, 192, 2, 2>();
  } else if constexpr (T == 224 || T == 240) {
    // <= 512 interior faces per tile: the face phase is exactly two two-face trips of 128 threads (T = 256 needs a third, mostly empty one)
    if (v == 1) return vjp_mk<T, ML, MF, NF, 160, 3, 1>();
    return vjp_mk<T, ML, MF, NF, 128, 3, 2>();
  } else if constexpr (T == 192) {
    // 96x4 two faces per trip 0.91 | 128x4 one face 0.96 | 128x4 two faces 1.07 (spills)
    if (v == 1) return vjp_mk<T, ML, MF, NF, 128, 4, 1>();
    if (v == 2) return vjp_mk<T, ML, MF, NF, 128, 4, 2>();
    return vjp_mk<T, ML, MF, NF, 96, 4, 2>();
  } else if constexpr (T == 128 && NF == 4) {
    return vjp_mk<T, ML, MF, NF, 128, 4, 1>();
  } else if constexpr (T == 128) {
    return vjp_mk<T, ML, MF, NF, 128, 3, 1>();
  } else {
    return vjp_mk<T, ML, MF, NF, 512, 1, 1>();
  }
}
VjpKernel vjp_kernel(int cfg_id, int variant) {
  switch (cfg_id) {
#define X(id, T, ML, MF, NF, TH, MB) case id: return vjp_pick<T, ML, MF, NF>(variant);
    HG_TILE_CONFIGS(X)
#undef X
  }
  return VjpKernel{};
}

}  // namespace

int fused_vjp_smem_bytes(int cfg_id) { return vjp_kernel(cfg_id, 0).smem; }

int fused_vjp_prepare(hg_ctx* ctx, int cfg_id) {
  const VjpKernel k = vjp_kernel(cfg_id, ctx->opt.reserved[2]);
  if (!k.fn) { ctx->err = "no VJP tile configuration"; return HG_ERR_ARG; }
  if (k.smem > 227 * 1024) { ctx->err = "VJP tile does not fit shared memory"; return HG_ERR_ARG; }
  cudaError_t e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k.fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { ctx->err = std::string("cudaFuncSetAttribute(vjp): ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}

// The tile kernel over the tiles tile_order[tile_base .. tile_base + n_run) (tile_order NULL: all tiles, identity).
// The inlet coefficients must be current (fused_inlet_coef).
int fused_vjp_tiles(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_lam, double* d_Qbar, const int32_t* tile_order,
                    int32_t tile_base, int32_t n_run, int comm_mode, bool pdl) {
  FusedDev& d = ctx->fd;
  const FusedHost& fh = ctx->fh;
  if (n_run == 0) return HG_OK;
  VjpArgs a;
  a.N = (int32_t)ctx->N; a.n_tiles = fh.n_tiles; a.want_s0 = ctx->active == HG_PARAM_ZB ? 1 : 0;
  a.Ns = fh.Ns; a.c = ctx->c;
  a.tile_desc = d.tile_desc.p; a.halo = d.halo.p; a.bface_e = d.bface_e.p; a.face_lr = d.face_lr.p;
  a.cf_idx = d.cf_idx.p; a.face_nx = d.face_nx.p; a.face_ny = d.face_ny.p; a.face_len = d.face_len.p;
  a.area = d.area.p; a.hstill = d.hstill.p; a.zb = d.zb.p; a.S0x = d.S0x.p; a.S0y = d.S0y.p; a.mann = d.mann.p;
  a.bc_type = d.bc_type.p; a.bc_group = d.bc_group.p; a.bc_nx = d.bc_nx.p; a.bc_ny = d.bc_ny.p;
  a.bc_l23 = d.bc_l23.p; a.bc_hstill = d.bc_hstill.p; a.bc_zb = d.bc_zb.p; a.inlet_coef = d.inlet_coef.p;
  a.halo_off = d.halo_off.p; a.halo_cnt = d.halo_cnt.p; a.halo_recv = d.halo_recv.p;
  a.wse = d.wse.p; a.Q = d_Q; a.lam = d_lam; a.Qbar = d_Qbar; a.nbar = d.nbar.p; a.s0bar = d.s0bar.p;
  a.ent_c = d.ent_c.p; a.ent_n = d.ent_n.p; a.ent_z = d.ent_z.p;
  a.tile_order = tile_order; a.tile_base = tile_base;
  if (comm_mode) {   // library-owned exchange (see launch_rhs): 1 = the whole mesh in one launch, 2 = a pipeline stage with the band
    const hg_comm* cm = ctx->comm;
    a.halo_recv = cm->recv[cm->epoch & 1];
    a.cw.flags = cm->flags; a.cw.epoch = cm->epoch; a.cw.n = cm->n; a.cw.err = d.err.p;
    if (comm_mode == 1) {
      a.tile_order = d.comm_order.p; a.tile_base = 0; n_run = fh.n_tiles;
      a.cw.from = fh.comm_band0; a.cw.to = fh.comm_band0 + (fh.n_tiles - fh.n_interior_tiles);
    } else {
      a.cw.from = 0; a.cw.to = n_run >= 0 ? n_run : fh.n_tiles;
    }
  }
  const unsigned grid = (unsigned)(n_run >= 0 ? n_run : fh.n_tiles);
  a.n_run = (int32_t)grid;
  const VjpKernel kk = vjp_kernel(cfg_id, ctx->opt.reserved[2]);
  if (!kk.fn) { ctx->err = "no VJP tile configuration"; return HG_ERR_ARG; }
  {
    const int r3 = ctx->opt.reserved[3];
    a.prefetch = r3 < 0 ? 0 : (r3 > 0 ? r3 : ctx->n_sm * kk.ctas_per_sm);
    void* kargs[] = {(void*)&a};
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3((unsigned)kk.threads); lc.dynamicSmemBytes = (size_t)kk.smem; lc.stream = ctx->stream;
    cudaLaunchAttribute pdl_attr[1];   // programmatic dependent of the k_inlet_coef launched just before (hg_device.cuh)
    pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = pdl_attr; lc.numAttrs = pdl ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelExC(&lc, kk.fn, kargs);
    if (le != cudaSuccess) { ctx->err = std::string("fused_vjp launch: ") + cudaGetErrorString(le); return HG_ERR_CUDA; }
  }
  ctx->launches++;
#ifdef HG_PHASE_CLOCKS
  {
    static int n_launch = 0;
    if (++n_launch % 40 == 0) {
      unsigned long long h[8];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpyFromSymbol(h, g_phase_clk, sizeof(h));
      const double den = 40.0 * (double)grid;
      fprintf(stderr, "[phase clocks / tile] desc %.0f  issue %.0f  copy wait %.0f  phase1+halo %.0f  phase2a %.0f  phase2b+sync %.0f  phase3 %.0f  (bulk-copy issue %.0f)\n",
              h[0] / den, h[1] / den, h[2] / den, h[3] / den, h[4] / den, h[5] / den, h[6] / den, h[7] / den);
      memset(h, 0, sizeof(h));
      cudaMemcpyToSymbol(g_phase_clk, h, sizeof(h));
    }
  }
#endif
  return HG_OK;
}

// Boundary-wide couplings (inlet conveyance split) and the parameter adjoint of the active parameter (pbar, device),
// after every tile has run.
int fused_vjp_finish(hg_ctx* ctx, const double* d_Q, double* d_Qbar) {
  FusedDev& d = ctx->fd;
  const FusedHost& fh = ctx->fh;
  const int th = 256;
  if (ctx->n_inletq > 0) {
    k_inlet_adj<<<(unsigned)ctx->n_inletq, 256, 0, ctx->stream>>>(ctx->c, d.inlet_ptr.p, d.bc_cell.p, d.bc_l53.p, d_Q,
                                                                 d.hstill.p, d.mann.p, d.inlet_A.p, d.inlet_coef.p, d.ent_c.p,
                                                                 d.ent_h.p, d.ent_n.p, d.Qinbar.p);
    ctx->launches++;
    if (ctx->nbcell > 0) {
      k_bc_scatter<<<(unsigned)((ctx->nbcell + th - 1) / th), th, 0, ctx->stream>>>((int32_t)ctx->nbcell, d.bcell.p, d.bcell_ptr.p,
                                                                                 d.bcell_ent.p, d.bc_type.p, d.ent_h.p,
                                                                                 d.ent_n.p, d_Qbar, d.nbar.p);
      ctx->launches++;
    }
  }
  // ---- parameter adjoints
  if (ctx->active == HG_PARAM_MANNING) {
    if (ctx->n_mat <= kZoneFew) {
      if (d.zone_part2.n < (size_t)kZoneGrid * kZoneFew + 2) {
        if (d.zone_part2.alloc((size_t)kZoneGrid * kZoneFew + 2) != cudaSuccess) { ctx->err = "cudaMalloc(zone_part2)"; return HG_ERR_CUDA; }
        cudaMemsetAsync(d.zone_part2.p + (size_t)kZoneGrid * kZoneFew, 0, 16, ctx->stream);   // the ticket counter
      }
      unsigned int* ticket = reinterpret_cast<unsigned int*>(d.zone_part2.p + (size_t)kZoneGrid * kZoneFew);
      const unsigned g2 = (unsigned)std::min<int64_t>(kZoneGrid, (ctx->N + 1023) / 1024);
      k_zone_reduce<kZoneFew><<<g2, 256, 0, ctx->stream>>>((int32_t)ctx->N, (int32_t)ctx->n_mat, d.matid8.p, d.nbar.p, d.zone_part2.p, ticket, d.pbar.p);
      ctx->launches++;
    } else {
      const int nblocks = (int)((ctx->N + kZoneChunk - 1) / kZoneChunk);
      if (d.zone_part.n < (size_t)nblocks * ctx->n_mat) {
        if (d.zone_part.alloc((size_t)nblocks * ctx->n_mat) != cudaSuccess) { ctx->err = "cudaMalloc(zone_part)"; return HG_ERR_CUDA; }
      }
      k_zone_partial<<<nblocks, kZoneBlock, kZoneBlock * sizeof(double), ctx->stream>>>((int32_t)ctx->N, (int32_t)ctx->n_mat, d.matid.p,
                                                                                        d.nbar.p, d.zone_part.p);
      k_zone_final_many<<<(unsigned)ctx->n_mat, 256, 0, ctx->stream>>>(nblocks, (int32_t)ctx->n_mat, d.zone_part.p, d.pbar.p);
      ctx->launches += 2;
    }
  } else if (ctx->active == HG_PARAM_ZB) {
    PlainDev& p = ctx->pd;
    k_zb_bar<<<(unsigned)((ctx->N + th - 1) / th), th, 0, ctx->stream>>>((int32_t)ctx->N, fh.Ns, d.iperm.p, p.cf_ptr.p, p.cf_nb.p,
                                                                       p.cf_rev.p, p.cf_nx.p, p.cf_ny.p, p.cf_len.p, p.area.p,
                                                                       d.s0bar.p, d.pbar.p);
    ctx->launches++;
    if (ctx->nbcell > 0) {
      k_zb_bar_exit<<<(unsigned)((ctx->nbcell + th - 1) / th), th, 0, ctx->stream>>>((int32_t)ctx->nbcell, d.bcell_ref.p,
                                                                                  d.bcell_ptr.p, d.bcell_ent.p, d.bc_type.p,
                                                                                  d.ent_z.p, d.pbar.p);
      ctx->launches++;
    }
  } else if (ctx->active == HG_PARAM_Q) {
    cudaMemcpyAsync(d.pbar.p, d.Qinbar.p, ctx->n_inletq * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
  } else if (ctx->active == HG_PARAM_UDE) {   // nbar is final here (friction + inlet conveyance): pull it back through the network
    const int rc = ude_adjoint(ctx, d_Q, d_Qbar);
    if (rc != HG_OK) return rc;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("fused_vjp launch: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}


// Qbar (internal order, [3Ns]) and the parameter adjoint for the active parameter (pbar, device).
int fused_vjp(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_lam, double* d_Qbar) {
  if (ctx->active == HG_PARAM_UDE) {
    const int rc0 = ude_eval_n(ctx, d_Q);
    if (rc0 != HG_OK) return rc0;
  }
  int use_comm = 0;
  if (hg_comm_ready(ctx)) {   // the adjoint of a cut face needs the remote cell's state AND cotangent
    hg_comm* cm = ctx->comm;
    if (cm->auto_exchange) {
      const int rc0 = comm_push(ctx, d_Q, d_lam);
      if (rc0 != HG_OK) return rc0;
    } else if (!cm->pushed) {
      ctx->err = "halo exchange: auto mode is off and hg_comm_exchange(with_lambda = 1) was not called before this evaluation";
      return HG_ERR_STATE;
    }
    cm->pushed = false;
    use_comm = 1;
  }
  // the conveyance sum last: the tile kernel is its (or the halo push's) programmatic dependent
  const bool pdl = ctx->n_inletq > 0 || use_comm == 1;
  if (ctx->n_inletq > 0) fused_inlet_coef(ctx, d_Q, use_comm == 1 && ctx->comm->auto_exchange);
  const int rc = fused_vjp_tiles(ctx, cfg_id, d_Q, d_lam, d_Qbar, nullptr, 0, -1, use_comm, pdl);
  return rc != HG_OK ? rc : fused_vjp_finish(ctx, d_Q, d_Qbar);
}

// ncell_bar in reference order (UDE hook): dst[r] = nbar[iperm[r]]
int fused_nbar_to_ref(hg_ctx* ctx, double* d_dst) {
  const int th = 256;
  k_gather1<<<(unsigned)((ctx->N + th - 1) / th), th, 0, ctx->stream>>>((int32_t)ctx->N, ctx->fd.iperm.p, ctx->fd.nbar.p, d_dst);
  ctx->launches++;
  return cudaGetLastError() == cudaSuccess ? HG_OK : HG_ERR_CUDA;
}

}  // namespace hg
