// Plain (strict) path: one thread per cell, the reference's own evaluation order -- every interior
// face evaluated from both of its cells, left-to-right accumulation over the cell's faces -- with
// FMA contraction disabled (this file is compiled with -fmad=false).  It exists as the parity
// vehicle: apart from libm's pow it performs the same IEEE operations as Hydrograd.jl's
//   swe_2d_rhs                semi_discretize_swe_2D.jl:18-277
//   process_all_boundaries_2d bc_2D.jl:575-875
//   Riemann_2D_Roe            swe_2D_solvers.jl:4-164
// The performance path is hg_fused.cu.
#include "hg_ctx.h"

namespace hg {
namespace {

__device__ __forceinline__ double smooth_abs(double x) { return sqrt(x * x + EPS); }
__device__ __forceinline__ double smooth_sqrt(double x) { return sqrt(x + EPS); }
__device__ __forceinline__ double smooth_pow2(double x) { double y = x + EPS; return y * y; }

// swe_2D_solvers.jl:4-164 in the reference's operation order
__device__ void roe_ref(double xiL, double hstL, double hL, double huL, double hvL, double zbL, double xiR,
                        double hstR, double hR, double huR, double hvR, double zbR, double g, double nx, double ny,
                        double hmin, double& o0, double& o1, double& o2) {
  if (hL <= hmin && hR <= hmin) { o0 = o1 = o2 = 0.0; return; }
  else if ((hL + zbL) < (zbR + hmin) && hR <= hmin) { hR = hL; huR = -huL; hvR = -hvL; }
  else if ((hR + zbR) < (zbL + hmin) && hL <= hmin) { hL = hR; huL = -huR; hvL = -hvR; }
  else if (hL <= hmin) {
    double p = (0.5 * g) * smooth_pow2(hR);
    o0 = huR * nx + hvR * ny;
    o1 = (huR * (huR / hR) + p) * nx + huR * (hvR / hR) * ny;
    o2 = (hvR * (huR / hR)) * nx + (hvR * (hvR / hR) + p) * ny;
    return;
  } else if (hR <= hmin) {
    double p = (0.5 * g) * smooth_pow2(hL);
    o0 = huL * nx + hvL * ny;
    o1 = (huL * (huL / hL) + p) * nx + huL * (hvL / hL) * ny;
    o2 = (hvL * (huL / hL)) * nx + (hvL * (hvL / hL) + p) * ny;
    return;
  }
  double uL = huL / hL, vL = hvL / hL, uR = huR / hR, vR = hvR / hR;
  double sL = smooth_sqrt(hL), sR = smooth_sqrt(hR);
  double hRoe = (hL + hR) / 2.0;
  double uRoe = (sL * uL + sR * uR) / (sL + sR);
  double vRoe = (sL * vL + sR * vR) / (sL + sR);
  double un = uRoe * nx + vRoe * ny;
  double c = smooth_sqrt(g * hRoe);
  double o2c = 1.0 / 2.0 / c;
  double R22 = uRoe - c * nx, R23 = uRoe + c * nx, R32 = vRoe - c * ny, R33 = vRoe + c * ny;
  double L11 = -(uRoe * ny - vRoe * nx), L12 = ny, L13 = -nx;
  double L21 = un * o2c + 0.5, L22 = -nx * o2c, L23 = -ny * o2c;
  double L31 = -un * o2c + 0.5, L32 = nx * o2c, L33 = ny * o2c;
  double a1 = smooth_abs(un), a2 = smooth_abs(un - c), a3 = smooth_abs(un + c);
  double d1 = xiR - xiL, d2 = huR - huL, d3 = hvR - hvL;
  double w1 = (L11 * d1 + L12 * d2) + L13 * d3;
  double w2 = (L21 * d1 + L22 * d2) + L23 * d3;
  double w3 = (L31 * d1 + L32 * d2) + L33 * d3;
  double z1 = a1 * w1, z2 = a2 * w2, z3 = a3 * w3;
  double y1 = z2 + z3;
  double y2 = (ny * z1 + R22 * z2) + R23 * z3;
  double y3 = (-nx * z1 + R32 * z2) + R33 * z3;
  double pL = (0.5 * g) * (smooth_pow2(xiL) + 2.0 * xiL * hstL);
  double pR = (0.5 * g) * (smooth_pow2(xiR) + 2.0 * xiR * hstR);
  double f1L = huL * nx + hvL * ny;
  double f2L = (huL * uL + pL) * nx + huL * vL * ny;
  double f3L = (hvL * uL) * nx + (hvL * vL + pL) * ny;
  double f1R = huR * nx + hvR * ny;
  double f2R = (huR * uR + pR) * nx + huR * vR * ny;
  double f3R = (hvR * uR) * nx + (hvR * vR + pR) * ny;
  o0 = (f1L + f1R - y1) / 2.0;
  o1 = (f2L + f2R - y2) / 2.0;
  o2 = (f3L + f3R - y3) / 2.0;
}

struct PlainArgs {
  int32_t N, B, n_inlet, active;
  Consts c;
  const int32_t *cf_ptr, *cf_nb;
  const double *cf_nx, *cf_ny, *cf_len;
  const double *area, *hstill, *zb, *S0x, *S0y, *mann, *ks;
  MannFn mfn;
  const int32_t* matid;
  const int32_t *bc_type, *bc_group, *bc_ghost, *bc_cell, *inlet_ptr;
  const double *bc_nx, *bc_ny, *bc_l53, *bc_l23, *hstill_g, *zb_g;
  double *gh, *gqx, *gqy, *gxi;
  const double *Qin, *wse, *Q, *params;
  double* dQ;
  int32_t* err;
};

__device__ __forceinline__ void load_cell(const PlainArgs& a, int32_t i, double& xi, double& h, double& qx, double& qy) {
  xi = a.Q[i];
  h = xi + a.hstill[i];
  bool dry = h <= a.c.h_small;  // semi_discretize_swe_2D.jl:104-106
  h = dry ? a.c.h_small : h;
  qx = dry ? 0.0 : a.Q[a.N + i];
  qy = dry ? 0.0 : a.Q[2 * a.N + i];
}
__device__ __forceinline__ double mann_of(const PlainArgs& a, int32_t i) {
  if (a.mfn.type)   // variable Manning's n of forward simulations (semi_discretize_swe_2D.jl:140-149)
    return manning_of_state(a.mfn, a.Q[i], a.Q[a.N + i], a.Q[2 * (int64_t)a.N + i], a.hstill[i], a.ks[i], a.c.h_small);
  return a.active == HG_PARAM_MANNING ? a.params[a.matid[i]] : a.mann[i];
}
__device__ __forceinline__ double zb_of(const PlainArgs& a, int32_t i) {
  return a.active == HG_PARAM_ZB ? a.params[i] : a.zb[i];
}

// process_all_boundaries_2d (bc_2D.jl:575-875): ghost states, written in GHOST order
__global__ void k_plain_ghost(PlainArgs a) {
  __shared__ double coef[64];
  const double hs = a.c.h_small;
  for (int32_t k = threadIdx.x; k < a.n_inlet; k += blockDim.x) {
    double tot = 0.0;  // sequential left fold like the reference's generator sum (:674-676)
    bool first = true;
    for (int32_t e = a.inlet_ptr[k]; e < a.inlet_ptr[k + 1]; ++e) {
      double xi, h, qx, qy;
      int32_t c = a.bc_cell[e];
      load_cell(a, c, xi, h, qx, qy);
      double wet = h > hs ? 1.0 : 0.0;
      double term = a.bc_l53[e] * h / mann_of(a, c) * wet;
      tot = first ? term : tot + term;
      first = false;
    }
    if (!(tot > 1e-10)) atomicExch(a.err, HG_ERR_CONVEYANCE);  // :678-680
    double Q = a.active == HG_PARAM_Q ? a.params[k] : a.Qin[k];
    if (k < 64) coef[k] = Q / tot;
  }
  __syncthreads();
  for (int32_t e = threadIdx.x; e < a.B; e += blockDim.x) {
    int32_t c = a.bc_cell[e], gi = a.bc_ghost[e], t = a.bc_type[e], k = a.bc_group[e];
    double xi, h, qx, qy, hg, gx, gy;
    load_cell(a, c, xi, h, qx, qy);
    double nx = a.bc_nx[e], ny = a.bc_ny[e];
    if (t == BC_INLETQ) {
      double wet = h > hs ? 1.0 : 0.0;
      double vn = coef[k] * a.bc_l23[e] / mann_of(a, c);  // :690-691
      hg = h;
      gx = -h * vn * nx * wet;                             // :693-694
      gy = -h * vn * ny * wet;
    } else if (t == BC_EXITH) {
      hg = fmax(hs, a.wse[k] - zb_of(a, c));               // :763-764
      gx = qx; gy = qy;
    } else if (t == BC_WALL) {
      hg = h; gx = -qx; gy = -qy;                          // :796
    } else {
      double vdn = qx * nx + qy * ny;                      // :819-827
      hg = h;
      gx = qx - 2.0 * vdn * nx;
      gy = qy - 2.0 * vdn * ny;
    }
    a.gh[gi] = hg; a.gqx[gi] = gx; a.gqy[gi] = gy;
    a.gxi[gi] = hg - a.hstill_g[gi];                       // semi_discretize_swe_2D.jl:220
  }
}

// compute_inviscid_fluxes + compute_source_terms, one thread per cell
__global__ void k_plain_cell(PlainArgs a) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const double g = a.c.g, hs = a.c.h_small;
  double xi, h, qx, qy;
  load_cell(a, i, xi, h, qx, qy);
  const double zbi = zb_of(a, i), hsti = a.hstill[i];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  double gx = 0.0, gy = 0.0;  // Green-Gauss bed gradient when zb is the active parameter
  for (int32_t k = a.cf_ptr[i]; k < a.cf_ptr[i + 1]; ++k) {
    const int32_t r = a.cf_nb[k];
    const double nx = a.cf_nx[k], ny = a.cf_ny[k], L = a.cf_len[k];
    double xr, hr, qxr, qyr, zbr, hstr;
    if (r < a.N) {
      load_cell(a, r, xr, hr, qxr, qyr);
      zbr = zb_of(a, r); hstr = a.hstill[r];
    } else {
      const int32_t gi = r - a.N;
      xr = a.gxi[gi]; hr = a.gh[gi]; qxr = a.gqx[gi]; qyr = a.gqy[gi]; hstr = a.hstill_g[gi];
      // update_ghost_cells_scalar (fvm_schemes_2D.jl:3-30): zb_ghost = zb of the internal cell
      zbr = a.active == HG_PARAM_ZB ? zbi : a.zb_g[gi];
    }
    double f0, f1, f2;
    roe_ref(xi, hsti, h, qx, qy, zbi, xr, hstr, hr, qxr, qyr, zbr, g, nx, ny, hs, f0, f1, f2);
    s0 = s0 + f0 * L; s1 = s1 + f1 * L; s2 = s2 + f2 * L;
    if (a.active == HG_PARAM_ZB) {
      double zf = (r < a.N) ? (zbi + zbr) / 2.0 : zbi;  // cells_to_faces_scalar, fvm_schemes_2D.jl:89-105
      gx = gx + nx * zf * L; gy = gy + ny * zf * L;
    }
  }
  const double A = a.area[i];
  double S0x, S0y;
  if (a.active == HG_PARAM_ZB) { S0x = -1.0 * (gx / A); S0y = -1.0 * (gy / A); }
  else { S0x = a.S0x[i]; S0y = a.S0y[i]; }
  const double n = mann_of(a, i);
  const double mag = smooth_sqrt(qx * qx + qy * qy);
  const double coef = g * (n * n) / (a.c.k_n * a.c.k_n) / pow(h + hs, 7.0 / 3.0);
  const double frx = coef * mag * qx, fry = coef * mag * qy;
  const double wet = h > hs ? 1.0 : 0.0;
  a.dQ[i] = -s0 / A + 0.0;
  a.dQ[a.N + i] = -s1 / A + wet * (g * xi * S0x - frx);
  a.dQ[2 * a.N + i] = -s2 / A + wet * (g * xi * S0y - fry);
}

}  // namespace

int plain_rhs(hg_ctx* ctx, const double* d_Q, double* d_out) {
  PlainDev& p = ctx->pd;
  PlainArgs a;
  a.N = (int32_t)ctx->N; a.B = (int32_t)ctx->B; a.n_inlet = (int32_t)ctx->n_inletq; a.active = ctx->active;
  a.c = ctx->c;
  a.cf_ptr = p.cf_ptr.p; a.cf_nb = p.cf_nb.p; a.cf_nx = p.cf_nx.p; a.cf_ny = p.cf_ny.p; a.cf_len = p.cf_len.p;
  a.area = p.area.p; a.hstill = p.hstill.p; a.zb = p.zb.p; a.S0x = p.S0x.p; a.S0y = p.S0y.p; a.mann = p.mann.p;
  a.ks = p.ks.p; a.mfn = ctx->mfn;
  a.matid = p.matid.p;
  a.bc_type = p.bc_type.p; a.bc_group = p.bc_group.p; a.bc_ghost = p.bc_ghost.p; a.bc_cell = p.bc_cell.p;
  a.inlet_ptr = p.inlet_ptr.p;
  a.bc_nx = p.bc_nx.p; a.bc_ny = p.bc_ny.p; a.bc_l53 = p.bc_l53.p; a.bc_l23 = p.bc_l23.p;
  a.hstill_g = p.hstill_g.p; a.zb_g = p.zb_g.p;
  a.gh = p.gh.p; a.gqx = p.gqx.p; a.gqy = p.gqy.p; a.gxi = p.gxi.p;
  a.Qin = p.Qin.p; a.wse = p.wse.p; a.Q = d_Q; a.params = p.params.p; a.dQ = d_out; a.err = p.err.p;
  if (ctx->n_inletq > 64) { ctx->err = "plain path supports at most 64 inlet-q boundaries"; return HG_ERR_ARG; }
  if (ctx->B > 0) {
    k_plain_ghost<<<1, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
  }
  const int threads = 128;
  k_plain_cell<<<(unsigned)((ctx->N + threads - 1) / threads), threads, 0, ctx->stream>>>(a);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("plain_rhs launch: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}

}  // namespace hg
