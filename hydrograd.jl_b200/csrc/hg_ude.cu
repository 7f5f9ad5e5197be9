// UDE closure on the device: Manning's n of every cell from the neural network of hg_set_ude_model (forward) and the
// pull-back of nbar through it onto the state and the network parameters (reverse).  The per-cell arithmetic lives in
// hg_ude.h (shared with the CPU check of tests/); this file is the launch structure around it:
//
//   forward   [whole-array LayerNorm only: per hidden layer l = 1..L one pass that recomputes the network up to the
//             activations of layer l and reduces their (count, mean, M2) -> (mean, 1/sqrt(var + eps)) of that layer]
//             k_ude_n: n of every cell -> the Manning block the fused RHS / VJP kernels stage
//   reverse   [whole-array LayerNorm only: per hidden layer l = L..1 one pass reducing mean(g), mean(g xhat)]
//             k_ude_adj: Qbar += (dn/dQ)^T nbar per cell, thetabar partial sums per block -> k_ude_colsum
//
// The kernels are templates on the network shape: for the shapes the reference ships (examples/SWE_2D/UDE/*/run_control.json:
// 1 or 3 inputs, two hidden layers of 3 tanh units) every loop bound, activation and parameter offset is a compile-time
// constant, so the per-cell tape and the thetabar accumulators live in registers; any other shape runs the generic
// instantiation (same source, run-time bounds, tape in local memory).  theta is staged in shared memory in the canonical
// order of hg_ude.h once per block.
//
// Every reduction has a fixed shape (per-thread sequential over a fixed cell range, warp butterfly / shared-memory tree,
// per-block partials combined in block order), so results are bit-reproducible run to run.  This first version is NOT fused
// into phase 1 of k_fused_rhs (the forward-simulation closures are): one extra pass over the state per RHS (40 B read +
// 8 B written per cell) plus the statistics passes; DESIGN.md lists the fusion as the next step.
#include "hg_ctx.h"
#include "hg_ude.h"

namespace hg {
namespace {

constexpr int kUB = 128;         // threads per block
constexpr int kUChunk = 2048;    // cells per block (16 per thread)

struct UdeArgs {
  ude::Model m;
  ude::ThetaMap map;
  int64_t N, Ns;
  double hs;
  const double *Q, *hstill, *ks, *theta, *stats, *bstats;
};

// theta -> shared memory, canonical order
__device__ __forceinline__ void load_theta(const UdeArgs& a, int P, double* th) {
  for (int k = threadIdx.x; k < P; k += kUB) th[k] = a.theta[a.map.to_user[k]];
  __syncthreads();
}
template <class S>
__device__ __forceinline__ void cell_inputs(const UdeArgs& a, const ude::Model& m, int64_t i, ude::Inputs& in) {
  ude::inputs<S>(m, a.Q[i], a.Q[a.Ns + i], a.Q[2 * a.Ns + i], a.hstill[i], ude::s_nin<S>(m) == 3 ? a.ks[i] : 1.0, a.hs, in);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;   // lane 0
}

// ---- (count, mean, M2) of the activations of hidden layer `layer`: one partial triple per block
// (the layer is a template parameter: with a run-time early exit inside forward() the tape does not stay in registers)
template <class S, int LAYER>
__global__ void __launch_bounds__(kUB) k_ude_stats(const UdeArgs a, double* __restrict__ part) {
  constexpr int layer = LAYER;
  __shared__ double th[S::P];
  __shared__ ude::Moments red[kUB];
  const ude::Model& m = a.m;
  load_theta(a, ude::s_np<S>(m), th);
  ude::Moments acc{0.0, 0.0, 0.0};
  const int64_t b0 = (int64_t)blockIdx.x * kUChunk, b1 = min(a.N, b0 + (int64_t)kUChunk);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kUB) {
    ude::Inputs in;
    ude::Tape t;
    cell_inputs<S>(a, m, i, in);
    ude::forward<S>(m, th, in.x, a.stats, layer, t);
#pragma unroll
    for (int l = 0; l < ude::MAXH; ++l)
      if (l == layer) {
#pragma unroll
        for (int j = 0; j < ude::MAXW; ++j)
          if (j < ude::s_width<S>(m, l)) ude::moments_push(acc, t.y[l][j]);
      }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kUB / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = ude::moments_merge(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[3 * blockIdx.x] = red[0].n;
    part[3 * blockIdx.x + 1] = red[0].mean;
    part[3 * blockIdx.x + 2] = red[0].m2;
  }
}
// one block: thread t merges the partials [t c, (t + 1) c) in order, then a tree over the threads
__global__ void __launch_bounds__(256) k_ude_stats_final(int nblk, const double* __restrict__ part, double eps, double* __restrict__ out2) {
  __shared__ ude::Moments red[256];
  const int c = (nblk + 255) / 256;
  ude::Moments acc{0.0, 0.0, 0.0};
  for (int b = threadIdx.x * c; b < min(nblk, (threadIdx.x + 1) * c); ++b)
    acc = ude::moments_merge(acc, ude::Moments{part[3 * b], part[3 * b + 1], part[3 * b + 2]});
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = ude::moments_merge(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out2[0] = red[0].mean;
    out2[1] = 1.0 / sqrt(red[0].m2 / red[0].n + eps);
  }
}

// ---- n of every cell
template <class S>
__global__ void __launch_bounds__(kUB) k_ude_n(const UdeArgs a, double* __restrict__ mann) {
  __shared__ double th[S::P];
  const ude::Model& m = a.m;
  load_theta(a, ude::s_np<S>(m), th);
  const int64_t b0 = (int64_t)blockIdx.x * kUChunk, b1 = min(a.N, b0 + (int64_t)kUChunk);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kUB) {
    ude::Inputs in;
    ude::Tape t;
    cell_inputs<S>(a, m, i, in);
    mann[i] = ude::forward<S>(m, th, in.x, a.stats, -1, t);
  }
}

// ---- sums of g and g xhat of hidden layer `layer` (g = abar * scale): two partial sums per block
template <class S, int LAYER>
__global__ void __launch_bounds__(kUB) k_ude_bstats(const UdeArgs a, const double* __restrict__ nbar, double* __restrict__ part) {
  constexpr int layer = LAYER;
  __shared__ double th[S::P];
  __shared__ double red[2][kUB / 32];
  const ude::Model& m = a.m;
  load_theta(a, ude::s_np<S>(m), th);
  double s1 = 0.0, s2 = 0.0;
  const int64_t b0 = (int64_t)blockIdx.x * kUChunk, b1 = min(a.N, b0 + (int64_t)kUChunk);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kUB) {
    ude::Inputs in;
    ude::Tape t;
    double g[ude::MAXW];
    cell_inputs<S>(a, m, i, in);
    ude::forward<S>(m, th, in.x, a.stats, -1, t);
    ude::backward<S>(m, th, in.x, t, a.bstats, nbar[i], layer, g, nullptr, nullptr);
#pragma unroll
    for (int l = 0; l < ude::MAXH; ++l)
      if (l == layer) {
#pragma unroll
        for (int j = 0; j < ude::MAXW; ++j)
          if (j < ude::s_width<S>(m, l)) {
            s1 += g[j];
            s2 += g[j] * t.xh[l][j];
          }
      }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < kUB / 32; ++w) s += red[threadIdx.x][w];
    part[2 * blockIdx.x + threadIdx.x] = s;
  }
}

// ---- the full reverse sweep: Qbar += (dn/dQ)^T nbar; per-block partial sums of thetabar (canonical order)
template <class S>
__global__ void __launch_bounds__(kUB) k_ude_adj(const UdeArgs a, const double* __restrict__ nbar, double* __restrict__ Qbar,
                                                 double* __restrict__ part) {
  __shared__ double th[S::P];
  __shared__ double red[S::P][kUB / 32];
  const ude::Model& m = a.m;
  load_theta(a, ude::s_np<S>(m), th);
  double acc[S::P];
  const int P = ude::s_np<S>(m);
#pragma unroll
  for (int k = 0; k < S::P; ++k) acc[k] = 0.0;
  const int64_t b0 = (int64_t)blockIdx.x * kUChunk, b1 = min(a.N, b0 + (int64_t)kUChunk);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kUB) {
    ude::Inputs in;
    ude::Tape t;
    double xbar[3] = {0.0, 0.0, 0.0};
    cell_inputs<S>(a, m, i, in);
    ude::forward<S>(m, th, in.x, a.stats, -1, t);
    ude::backward<S>(m, th, in.x, t, a.bstats, nbar[i], -1, nullptr, acc, xbar);
    double xib, qxb, qyb;
    ude::inputs_adj<S>(m, in, xbar, xib, qxb, qyb);
    Qbar[i] += xib;
    Qbar[a.Ns + i] += qxb;
    Qbar[2 * a.Ns + i] += qyb;
  }
#pragma unroll
  for (int k = 0; k < S::P; ++k) {
    if (k < P) {
      const double s = warp_sum(acc[k]);
      if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P; k += kUB) {
    double s = 0.0;
    for (int w = 0; w < kUB / 32; ++w) s += red[k][w];
    part[(int64_t)blockIdx.x * P + k] = s;
  }
}

// out[k] = scale * sum over the blocks (in block order) of part[b][k]; with a map, out[map[k]] (canonical -> caller's theta)
__global__ void k_ude_colsum(int nblk, int P, const double* __restrict__ part, double scale, double* __restrict__ out,
                             const ude::ThetaMap map, int use_map) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[(int64_t)b * P + k];
  out[use_map ? map.to_user[k] : k] = scale * s;
}

UdeArgs make_args(hg_ctx* ctx, const double* d_Q) {
  FusedDev& d = ctx->fd;
  UdeArgs a;
  a.m = ctx->ude;
  a.map = ctx->ude_map;
  a.N = ctx->N;
  a.Ns = ctx->fh.Ns;
  a.hs = ctx->c.h_small;
  a.Q = d_Q;
  a.hstill = d.hstill.p;
  a.ks = d.ks.p;
  a.theta = d.ude_theta.p;
  a.stats = d.ude_stats.p;
  a.bstats = d.ude_stats.p + 2 * ude::MAXH;
  return a;
}
int n_blocks(const hg_ctx* ctx) { return (int)((ctx->N + kUChunk - 1) / kUChunk); }
int launched(hg_ctx* ctx, int n, const char* what) {
  ctx->launches += n;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return HG_ERR_CUDA; }
  return HG_OK;
}
// which instantiation serves the model (opt.reserved[5] = 1 forces the generic one)
int spec_of(const hg_ctx* ctx) { return ctx->opt.reserved[5] == 1 ? 0 : ude::spec_of(ctx->ude); }
}  // namespace

int ude_spec_id(const hg_ctx* ctx) { return spec_of(ctx); }

// buffers of the UDE closure (called by hg_set_ude_model); theta is uploaded by bind_params
int ude_prepare(hg_ctx* ctx) {
  FusedDev& d = ctx->fd;
  const size_t nblk = (size_t)n_blocks(ctx);
  const size_t np_user = (size_t)std::max<int64_t>(ctx->ude_user_params, 1);
  if (d.ude_theta.n < np_user && d.ude_theta.alloc(np_user) != cudaSuccess) { ctx->err = "cudaMalloc(ude_theta)"; return HG_ERR_CUDA; }
  if (d.ude_stats.n < (size_t)(4 * ude::MAXH) && d.ude_stats.alloc(4 * ude::MAXH) != cudaSuccess) { ctx->err = "cudaMalloc(ude_stats)"; return HG_ERR_CUDA; }
  const size_t need = nblk * (size_t)std::max(ctx->ude.n_params, 3);
  if (d.ude_part.n < need && d.ude_part.alloc(need) != cudaSuccess) { ctx->err = "cudaMalloc(ude_part)"; return HG_ERR_CUDA; }
  if (d.pbar.n < np_user && d.pbar.alloc(np_user) != cudaSuccess) { ctx->err = "cudaMalloc(pbar)"; return HG_ERR_CUDA; }
  cudaMemsetAsync(d.ude_stats.p, 0, d.ude_stats.bytes(), ctx->stream);
  cudaMemsetAsync(d.pbar.p, 0, d.pbar.bytes(), ctx->stream);   // entries of theta the network does not use keep a zero adjoint
  return launched(ctx, 0, "ude_prepare");
}

// ManningN of every own cell from the state d_Q and the bound theta -> fd.mann (what the RHS / VJP kernels stage)
int ude_eval_n(hg_ctx* ctx, const double* d_Q) {
  FusedDev& d = ctx->fd;
  const UdeArgs a = make_args(ctx, d_Q);
  const int nblk = n_blocks(ctx);
  int n = 0;
  switch (spec_of(ctx)) {
#define X(id, S)                                                                                                       \
  case id: {                                                                                                           \
    using SS = HG_UDE_UNPAREN S;                                                                                       \
    if (a.m.ln_mode == HG_LN_WHOLE_ARRAY) {                                                                            \
      for (int l = 0; l < a.m.n_hidden; ++l) {                                                                         \
        if (l == 0) k_ude_stats<SS, 0><<<nblk, kUB, 0, ctx->stream>>>(a, d.ude_part.p);                                \
        else if (l == 1) k_ude_stats<SS, 1><<<nblk, kUB, 0, ctx->stream>>>(a, d.ude_part.p);                           \
        else k_ude_stats<SS, 2><<<nblk, kUB, 0, ctx->stream>>>(a, d.ude_part.p);                                       \
        k_ude_stats_final<<<1, 256, 0, ctx->stream>>>(nblk, d.ude_part.p, a.m.eps, d.ude_stats.p + 2 * l);             \
        n += 2;                                                                                                        \
      }                                                                                                                \
    }                                                                                                                  \
    k_ude_n<SS><<<nblk, kUB, 0, ctx->stream>>>(a, d.mann.p);                                                           \
  } break;
    HG_UDE_SPECS(X)
#undef X
  }
  return launched(ctx, n + 1, "ude_eval_n");
}

// after the VJP tile kernel and the inlet coupling (fd.nbar final): Qbar += (dn/dQ)^T nbar, fd.pbar = (dn/dtheta)^T nbar.
// The forward statistics of d_Q must be current (ude_eval_n(d_Q) ran before the tile kernel).
int ude_adjoint(hg_ctx* ctx, const double* d_Q, double* d_Qbar) {
  FusedDev& d = ctx->fd;
  const UdeArgs a = make_args(ctx, d_Q);
  const int nblk = n_blocks(ctx);
  const int P = a.m.n_params;
  int n = 0;
  switch (spec_of(ctx)) {
#define X(id, S)                                                                                                       \
  case id: {                                                                                                           \
    using SS = HG_UDE_UNPAREN S;                                                                                       \
    if (a.m.ln_mode == HG_LN_WHOLE_ARRAY) {                                                                            \
      for (int l = a.m.n_hidden - 1; l >= 0; --l) {                                                                    \
        if (l == 0) k_ude_bstats<SS, 0><<<nblk, kUB, 0, ctx->stream>>>(a, d.nbar.p, d.ude_part.p);                     \
        else if (l == 1) k_ude_bstats<SS, 1><<<nblk, kUB, 0, ctx->stream>>>(a, d.nbar.p, d.ude_part.p);                \
        else k_ude_bstats<SS, 2><<<nblk, kUB, 0, ctx->stream>>>(a, d.nbar.p, d.ude_part.p);                            \
        k_ude_colsum<<<1, 32, 0, ctx->stream>>>(nblk, 2, d.ude_part.p, 1.0 / ((double)ctx->N * a.m.width[l]),          \
                                                d.ude_stats.p + 2 * ude::MAXH + 2 * l, a.map, 0);                      \
        n += 2;                                                                                                        \
      }                                                                                                                \
    }                                                                                                                  \
    k_ude_adj<SS><<<nblk, kUB, 0, ctx->stream>>>(a, d.nbar.p, d_Qbar, d.ude_part.p);                                   \
  } break;
    HG_UDE_SPECS(X)
#undef X
  }
  k_ude_colsum<<<(P + 63) / 64, 64, 0, ctx->stream>>>(nblk, P, d.ude_part.p, 1.0, d.pbar.p, a.map, 1);
  return launched(ctx, n + 2, "ude_adjoint");
}

}  // namespace hg
