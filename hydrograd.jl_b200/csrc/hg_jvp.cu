// Forward mode of the RHS on the device (hg_rhs_jvp): dQ/dt and J_Q v + J_p pdot in one sweep of dual-number arithmetic -- the
// counterpart of ForwardDiff.Dual flowing through swe_2d_rhs (swe_2D_sensitivity.jl:34-80; the ForwardDiffSensitivity /
// ForwardSensitivity inversion options, solve_swe_2D.jl:230-235).  The arithmetic is hg_jvp_impl.h (shared with the g++ CPU
// check); this file is the launch structure, the same two launches as the strict path (hg_plain.cu), whose tables it uses:
//   k_jvp_ghost   one block: inlet-q conveyance coefficients (values and tangents) into shared memory, then one boundary
//                 entry per thread -> ghost states and their tangents
//   k_jvp_cell    one thread per cell: every face of the cell from its own side, reference summation order
// Compiled with -fmad=false like hg_plain.cu, so the values it returns are those of the strict path.
#include "hg_ctx.h"
#include "hg_jvp_impl.h"

namespace hg {
namespace {
static_assert((int)jvp::kInletQ == (int)BC_INLETQ && (int)jvp::kExitH == (int)BC_EXITH && (int)jvp::kWall == (int)BC_WALL &&
                  (int)jvp::kSymm == (int)BC_SYMM, "BcType");
static_assert((int)jvp::kParamZb == (int)HG_PARAM_ZB && (int)jvp::kParamManning == (int)HG_PARAM_MANNING &&
                  (int)jvp::kParamQ == (int)HG_PARAM_Q, "HG_PARAM_*");
constexpr int kMaxInlets = 64;

// K directions per launch: blockIdx.y selects the direction, i.e. the rows of V / pdot / the outputs and its own copy of the
// ghost buffers (strides in doubles; every direction recomputes the values -- the path is launch-bound on the meshes it
// serves, so K directions cost two launches instead of 2 K)
struct Batch {
  int64_t sV, sP, sB;
};
__device__ __forceinline__ void select_direction(jvp::Args& a, const Batch& b) {
  const int64_t y = blockIdx.y;
  if (a.V) a.V += y * b.sV;
  if (a.pdot) a.pdot += y * b.sP;
  a.dQ_d += y * b.sV;
  a.gh += y * b.sB; a.gqx += y * b.sB; a.gqy += y * b.sB; a.gxi += y * b.sB;
  a.gh_d += y * b.sB; a.gqx_d += y * b.sB; a.gqy_d += y * b.sB; a.gxi_d += y * b.sB;
  if (y > 0) a.dQ = nullptr;     // the values are delivered once
}

__global__ void __launch_bounds__(256) k_jvp_ghost(jvp::Args a, const Batch b) {
  select_direction(a, b);
  __shared__ double coef_v[kMaxInlets], coef_d[kMaxInlets];
  for (int32_t k = threadIdx.x; k < a.n_inlet; k += blockDim.x) {
    const jvp::Dual c = jvp::inlet_coef<jvp::Dual>(a, k);
    coef_v[k] = c.v; coef_d[k] = c.d;
  }
  __syncthreads();
  for (int32_t e = threadIdx.x; e < a.B; e += blockDim.x) {
    const bool inlet = a.bc_type[e] == jvp::kInletQ;
    const int32_t k = a.bc_group[e];
    jvp::ghost_entry<jvp::Dual>(a, e, inlet ? jvp::mk(coef_v[k], coef_d[k]) : jvp::mk(0.0, 0.0));
  }
}

__global__ void __launch_bounds__(128) k_jvp_cell(jvp::Args a, const Batch b) {
  select_direction(a, b);
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.N) jvp::cell<jvp::Dual>(a, i);
}

// ---- small kernels of the forward-sensitivity solve (hg_solve_tsit5_sens): the augmented state is U[(1 + K)][3N],
// row 0 the values, row k the partial with respect to parameter k
struct Comb {
  const double* k[7];
  double c[7];
  int n;
};
__global__ void k_sens_lincomb(int64_t len, double* __restrict__ y, const double* __restrict__ x, Comb cb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double v = x[i];
  for (int m = 0; m < cb.n; ++m) v += cb.c[m] * cb.k[m][i];
  y[i] = v;
}
constexpr int kErrBlock = 256;
// DiffEqBase's norm for Dual states: residual of entry i over abstol + max(|u_i|, |unew_i|) reltol with |.| the norm of the
// Dual (values and partials), squares summed over values AND partials.  One entry per thread, fixed-shape tree per block.
__global__ void __launch_bounds__(kErrBlock) k_sens_err_partial(int64_t n3, int rows, const double* __restrict__ u,
                                                                const double* __restrict__ unew, Comb cb, double abstol,
                                                                double reltol, double* __restrict__ part) {
  __shared__ double sh[kErrBlock];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (i < n3) {
    double nu = 0.0, nn = 0.0;
    for (int r = 0; r < rows; ++r) {
      const double a = u[r * n3 + i], b = unew[r * n3 + i];
      nu += a * a; nn += b * b;
    }
    const double den = abstol + fmax(sqrt(nu), sqrt(nn)) * reltol;
    for (int r = 0; r < rows; ++r) {
      double ut = 0.0;
      for (int m = 0; m < cb.n; ++m) ut += cb.c[m] * cb.k[m][r * n3 + i];
      const double q = ut / den;
      acc += q * q;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kErrBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(kErrBlock) k_sens_err_final(int nblocks, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double sh[kErrBlock];
  double acc = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += kErrBlock) acc += part[b];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kErrBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}
int launch_ok(hg_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}
}  // namespace

int sens_lincomb(hg_ctx* ctx, int64_t len, double* y, const double* x, int n, const double* const* k, const double* coef) {
  Comb cb;
  cb.n = n;
  for (int m = 0; m < 7; ++m) { cb.k[m] = m < n ? k[m] : nullptr; cb.c[m] = m < n ? coef[m] : 0.0; }
  k_sens_lincomb<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(len, y, x, cb);
  ctx->launches++;
  return launch_ok(ctx, "sens_lincomb");
}
int sens_err_blocks(int64_t n3) { return (int)((n3 + kErrBlock - 1) / kErrBlock); }
// sum over values and partials of the squared scaled residuals -> d_sum[0]
int sens_err_norm(hg_ctx* ctx, int64_t n3, int rows, const double* u, const double* unew, int n, const double* const* k,
                  const double* coef, double abstol, double reltol, double* d_part, double* d_sum) {
  Comb cb;
  cb.n = n;
  for (int m = 0; m < 7; ++m) { cb.k[m] = m < n ? k[m] : nullptr; cb.c[m] = m < n ? coef[m] : 0.0; }
  const int nb = sens_err_blocks(n3);
  k_sens_err_partial<<<nb, kErrBlock, 0, ctx->stream>>>(n3, rows, u, unew, cb, abstol, reltol, d_part);
  k_sens_err_final<<<1, kErrBlock, 0, ctx->stream>>>(nb, d_part, d_sum);
  ctx->launches += 2;
  return launch_ok(ctx, "sens_err_norm");
}

// K directions in one pair of launches: d_V [K][3N] (row stride sV; NULL = zero), d_pdot [K][n_params] (row stride sP) or
// nullptr, d_out [3N] or nullptr (values, once), d_out_dot [K][3N] (row stride sV); reference cell order throughout
int plain_jvp_batch(hg_ctx* ctx, const double* d_Q, const double* d_V, int64_t sV, const double* d_pdot, int64_t sP, double* d_out,
                    double* d_out_dot, int64_t K) {
  PlainDev& p = ctx->pd;
  if (ctx->n_inletq > kMaxInlets) { ctx->err = "plain path supports at most 64 inlet-q boundaries"; return HG_ERR_ARG; }
  if (K < 1 || K > 65535) { ctx->err = "plain_jvp: number of directions out of range"; return HG_ERR_ARG; }
  const size_t B = (size_t)std::max<int64_t>(ctx->B, 1);
  if (p.gh_d.n < B * (size_t)K) {     // one copy of the ghost buffers per direction
    cudaError_t e = p.gh_d.alloc(B * K);
    if (e == cudaSuccess) e = p.gqx_d.alloc(B * K);
    if (e == cudaSuccess) e = p.gqy_d.alloc(B * K);
    if (e == cudaSuccess) e = p.gxi_d.alloc(B * K);
    if (e == cudaSuccess) e = p.jgh.alloc(B * K);
    if (e == cudaSuccess) e = p.jgqx.alloc(B * K);
    if (e == cudaSuccess) e = p.jgqy.alloc(B * K);
    if (e == cudaSuccess) e = p.jgxi.alloc(B * K);
    if (e != cudaSuccess) { ctx->err = std::string("plain_jvp: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
  }
  jvp::Args a;
  a.N = (int32_t)ctx->N; a.B = (int32_t)ctx->B; a.n_inlet = (int32_t)ctx->n_inletq; a.active = ctx->active;
  a.g = ctx->c.g; a.k_n = ctx->c.k_n; a.h_small = ctx->c.h_small;
  a.cf_ptr = p.cf_ptr.p; a.cf_nb = p.cf_nb.p; a.cf_nx = p.cf_nx.p; a.cf_ny = p.cf_ny.p; a.cf_len = p.cf_len.p;
  a.area = p.area.p; a.hstill = p.hstill.p; a.zb = p.zb.p; a.S0x = p.S0x.p; a.S0y = p.S0y.p; a.mann = p.mann.p;
  a.matid = p.matid.p;
  a.bc_type = p.bc_type.p; a.bc_group = p.bc_group.p; a.bc_ghost = p.bc_ghost.p; a.bc_cell = p.bc_cell.p;
  a.inlet_ptr = p.inlet_ptr.p;
  a.bc_nx = p.bc_nx.p; a.bc_ny = p.bc_ny.p; a.bc_l53 = p.bc_l53.p; a.bc_l23 = p.bc_l23.p;
  a.hstill_g = p.hstill_g.p; a.zb_g = p.zb_g.p;
  a.gh = p.jgh.p; a.gqx = p.jgqx.p; a.gqy = p.jgqy.p; a.gxi = p.jgxi.p;
  a.gh_d = p.gh_d.p; a.gqx_d = p.gqx_d.p; a.gqy_d = p.gqy_d.p; a.gxi_d = p.gxi_d.p;
  a.Qin = p.Qin.p; a.wse = p.wse.p; a.Q = d_Q; a.V = d_V; a.params = p.params.p; a.pdot = d_pdot;
  a.dQ = d_out; a.dQ_d = d_out_dot; a.err = p.err.p;
  const Batch b{sV, sP, (int64_t)B};
  if (ctx->B > 0) {
    k_jvp_ghost<<<dim3(1, (unsigned)K), 256, 0, ctx->stream>>>(a, b);
    ctx->launches++;
  }
  const int threads = 128;
  k_jvp_cell<<<dim3((unsigned)((ctx->N + threads - 1) / threads), (unsigned)K), threads, 0, ctx->stream>>>(a, b);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string("plain_jvp launch: ") + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}

int plain_jvp(hg_ctx* ctx, const double* d_Q, const double* d_V, const double* d_pdot, double* d_out, double* d_out_dot) {
  return plain_jvp_batch(ctx, d_Q, d_V, 0, d_pdot, 0, d_out, d_out_dot, 1);
}

}  // namespace hg
