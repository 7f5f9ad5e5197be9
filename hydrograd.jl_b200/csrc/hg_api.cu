// extern "C" entry points of include/hydrograd_b200.h.  Owns the context, device memory and the
// stream; there is no CPU fallback: without a CUDA device hg_create fails with HG_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <utility>
#include <thread>
#include <vector>
#include <memory>
#include <mutex>

#include "hg_ctx.h"

namespace {
std::mutex g_err_mu;
std::string g_err;  // error of the last failed hg_create

void set_global_err(const std::string& s) {
  std::lock_guard<std::mutex> l(g_err_mu);
  g_err = s;
}

#define CK(ctx, expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
      return HG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

template <class T>
int up(hg_ctx* ctx, hg::DBuf<T>& b, const std::vector<T>& h) {
  CK(ctx, b.upload(h, ctx->stream));
  ctx->device_bytes += (int64_t)b.bytes();
  return HG_OK;
}
template <class T>
int al(hg_ctx* ctx, hg::DBuf<T>& b, size_t n) {
  CK(ctx, b.alloc(n));
  ctx->device_bytes += (int64_t)b.bytes();
  if (n) CK(ctx, cudaMemsetAsync(b.p, 0, b.bytes(), ctx->stream));
  return HG_OK;
}
// cell-indexed arrays of the fused path are padded to Ns elements (TMA copies may over-read the last tile)
template <class T>
int upN(hg_ctx* ctx, hg::DBuf<T>& b, const std::vector<T>& h, size_t padded) {
  CK(ctx, b.alloc(padded));
  ctx->device_bytes += (int64_t)b.bytes();
  CK(ctx, cudaMemsetAsync(b.p, 0, b.bytes(), ctx->stream));
  if (!h.empty()) CK(ctx, cudaMemcpyAsync(b.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}
#define TRY(x)                 \
  do {                         \
    int _rc = (x);             \
    if (_rc != HG_OK) return _rc; \
  } while (0)

template <class T>
std::vector<T> permuted(const T* src, const std::vector<int32_t>& perm) {
  std::vector<T> o(perm.size());
  const int64_t n = (int64_t)perm.size();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) o[i] = src[perm[i]];
  return o;
}

int check_err_flag(hg_ctx* ctx) {
  int32_t* flag = ctx->opt.path == 1 ? ctx->pd.err.p : ctx->fd.err.p;
  int32_t h = 0;
  CK(ctx, cudaMemcpyAsync(&h, flag, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (h != 0) {
    CK(ctx, cudaMemsetAsync(flag, 0, sizeof(int32_t), ctx->stream));
    if (h == HG_ERR_CONVEYANCE) ctx->err = "Total cross-sectional conveyance for an inlet-q boundary is not positive";
    else if (h == HG_ERR_COMM) ctx->err = "halo exchange: a neighbour's push did not arrive within the time-out (exchanges are collective: every rank must issue the same sequence of RHS / VJP calls)";
    else ctx->err = "device-side error flag " + std::to_string(h);
    return h;
  }
  return HG_OK;
}

using hg::Frozen;
hg_ctx* ext(hg_ctx* c) { return c; }
}  // namespace

static int upload_fields(hg_ctx* ctx) {
  hg_ctx* x = ext(ctx);
  const int64_t N = ctx->N, B = ctx->B;
  const Frozen& fr = x->fr;
  if (ctx->opt.path == 1) {
    hg::PlainDev& p = ctx->pd;
    CK(ctx, cudaMemcpyAsync(p.mann.p, fr.mann_ref.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(p.zb.p, fr.zb_ref.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(p.S0x.p, fr.S0_ref.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(p.S0y.p, fr.S0_ref.data() + N, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (B) CK(ctx, cudaMemcpyAsync(p.zb_g.p, fr.zbg_ghost.data(), B * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_inletq) CK(ctx, cudaMemcpyAsync(p.Qin.p, fr.Qin.data(), ctx->n_inletq * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_exith) CK(ctx, cudaMemcpyAsync(p.wse.p, fr.wse.data(), ctx->n_exith * 8, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    hg::FusedDev& d = ctx->fd;
    const auto& perm = ctx->fh.perm;
    auto mann = permuted(fr.mann_ref.data(), perm), zb = permuted(fr.zb_ref.data(), perm);
    auto s0x = permuted(fr.S0_ref.data(), perm), s0y = permuted(fr.S0_ref.data() + N, perm);
    std::vector<double> bzb(B);
    for (int64_t e = 0; e < B; ++e) bzb[e] = fr.zbg_ghost[ctx->bch.ghost[e]];
    CK(ctx, cudaMemcpyAsync(d.mann.p, mann.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d.zb.p, zb.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d.S0x.p, s0x.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d.S0y.p, s0y.data(), N * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (B) CK(ctx, cudaMemcpyAsync(d.bc_zb.p, bzb.data(), B * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_inletq) CK(ctx, cudaMemcpyAsync(d.Qin.p, fr.Qin.data(), ctx->n_inletq * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->n_exith) CK(ctx, cudaMemcpyAsync(d.wse.p, fr.wse.data(), ctx->n_exith * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));  // the temporaries above die here
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}

// Bind params_vector to the frozen fields (semi_discretize_swe_2D.jl:114-126, 153-161, 190-199).
// Parameters change once per optimiser iteration but the RHS runs thousands of times in between,
// so the binding is done here, once, and the RHS kernels stay parameter-agnostic.
// The reference evaluates the closures in forward simulations only (semi_discretize_swe_2D.jl:140-149): no derivative path
static int no_closure(hg_ctx* ctx, const char* what) {
  if (ctx->mfn.type == HG_MANNING_CONSTANT) return HG_OK;
  ctx->err = std::string(what) + ": not available while a variable Manning's n closure is set (forward simulation only)";
  return HG_ERR_ARG;
}

static int bind_params(hg_ctx* ctx, const double* params, int64_t np, int32_t active) {
  hg_ctx* x = ext(ctx);
  if (active < HG_PARAM_NONE || active > HG_PARAM_UDE) { ctx->err = "bad active_param"; return HG_ERR_ARG; }
  if (active == HG_PARAM_UDE) {
    if (!ctx->ude_set) { ctx->err = "active parameter UDE: no model (call hg_set_ude_model first)"; return HG_ERR_ARG; }
    if (ctx->opt.path == 1) { ctx->err = "active parameter UDE needs the fused path (strict = 0)"; return HG_ERR_ARG; }
  } else if (ctx->ude_set && active != HG_PARAM_NONE) {
    ctx->err = "a UDE model is set: the active parameter must be UDE or NONE";
    return HG_ERR_ARG;
  }
  const int64_t want = active == HG_PARAM_ZB ? ctx->N : active == HG_PARAM_MANNING ? ctx->n_mat : active == HG_PARAM_Q ? ctx->n_inletq :
                       active == HG_PARAM_UDE ? ctx->ude_user_params : 0;
  if (active != HG_PARAM_NONE && (np != want || !params)) {
    ctx->err = "params_vector has length " + std::to_string(np) + ", expected " + std::to_string(want);
    return HG_ERR_ARG;
  }
  if (active == HG_PARAM_MANNING && x->matid_ref.empty()) { ctx->err = "matID_cells was not provided at hg_create"; return HG_ERR_ARG; }
  if (active == HG_PARAM_MANNING) TRY(no_closure(ctx, "active parameter ManningN"));
  if (active == HG_PARAM_NONE) np = 0;
  const bool same = (active == x->last_active) && (np == (int64_t)x->last_params.size()) &&
                    (np == 0 || std::memcmp(params, x->last_params.data(), np * 8) == 0);
  if (same) return HG_OK;
  // restore pristine fields if the previous binding overwrote them
  // (a new theta over an old theta has nothing to restore: the network rewrites ManningN_cells in every RHS anyway)
  if (x->last_active > HG_PARAM_NONE && ctx->opt.path != 1 && !(active == HG_PARAM_UDE && x->last_active == HG_PARAM_UDE))
    TRY(upload_fields(ctx));
  x->last_active = active;
  x->last_params.assign(params, params + np);
  ctx->active = active;
  ctx->n_params = np;
  if (active == HG_PARAM_NONE) return HG_OK;
  if (ctx->opt.path == 1) {
    CK(ctx, cudaMemcpyAsync(ctx->pd.params.p, params, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return HG_OK;
  }
  hg::FusedDev& d = ctx->fd;
  if (active == HG_PARAM_UDE) {   // theta stays in its own small buffer; n of every cell is re-evaluated from the state by each RHS
    CK(ctx, cudaMemcpyAsync(d.ude_theta.p, x->last_params.data(), np * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return HG_OK;
  }
  CK(ctx, cudaMemcpyAsync(d.params.p, params, np * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (active == HG_PARAM_Q) {
    CK(ctx, cudaMemcpyAsync(d.Qin.p, d.params.p, np * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (active == HG_PARAM_MANNING) {
    TRY(hg::fused_bind_manning(ctx, d.params.p));
  } else {
    TRY(hg::fused_bind_zb(ctx, d.params.p));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}

extern "C" {

void hg_default_options(hg_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->device = 0; o->tile_cells = 256; o->reorder = 1; o->strict = 0; o->path = 0;
}
int hg_abi_version(void) { return HG_ABI_VERSION; }

const char* hg_last_error(const hg_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> l(g_err_mu);
  static thread_local std::string copy;
  copy = g_err;
  return copy.c_str();
}
int64_t hg_n_cells(const hg_ctx* ctx) { return ctx ? ctx->N : 0; }
int64_t hg_kernel_launches(const hg_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t hg_state_generation(const hg_ctx* ctx) { return ctx ? (int64_t)ctx->state_gen : -1; }

void hg_destroy(hg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->opt.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  hg_comm_disconnect(ctx);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_cmp) cudaEventDestroy(e);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  cudaStream_t s = ctx->own_stream;
  delete ctx;  // frees the device buffers
  if (s) cudaStreamDestroy(s);
}

static int create_impl(hg_ctx* ctx, const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f) {
  hg::StageTimer whole_timer("hg_create (all of the above)");
  // Driver initialisation (~0.6 s in a cold process) and the device's primary context are brought up on a helper thread
  // while this one validates and preprocesses the mesh; the guard joins it on every way out.  Descriptor errors are
  // therefore reported before a missing device is.
  struct Warm {
    int ndev = 0;
    cudaError_t status = cudaSuccess;
    std::thread th;
    explicit Warm(int dev) {
      th = std::thread([this, dev] {
        status = cudaGetDeviceCount(&ndev);
        if (status == cudaSuccess && dev >= 0 && dev < ndev && cudaSetDevice(dev) == cudaSuccess) cudaFree(nullptr);
      });
    }
    void join() { if (th.joinable()) th.join(); }
    ~Warm() { join(); }
  } warm(ctx->opt.device);

  std::vector<int32_t> cf_ptr, cf_nb, cf_face;
  std::vector<double> cf_nx, cf_ny, cf_len;
  TRY(hg::build_host(ctx, m, b, f, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face));
  {
    hg::StageTimer cuda_timer("cuda init (what the mesh checks did not hide)");
    warm.join();
  }
  if (warm.status != cudaSuccess || warm.ndev == 0) {
    ctx->err = "no CUDA device available (this library has no CPU fallback)";
    return HG_ERR_CUDA;
  }
  if (ctx->opt.device < 0 || ctx->opt.device >= warm.ndev) { ctx->err = "bad device ordinal"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  cudaDeviceProp prop;
  CK(ctx, cudaGetDeviceProperties(&prop, ctx->opt.device));
  if (prop.major != 10) {
    ctx->err = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this build is sm_100a only";
    return HG_ERR_CUDA;
  }
  CK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = ctx->stream;
  CK(ctx, cudaEventCreate(&ctx->ev0));
  CK(ctx, cudaEventCreate(&ctx->ev1));

  const int64_t N = ctx->N, B = ctx->B;
  hg_ctx* x = ext(ctx);
  Frozen& fr = x->fr;
  auto frozen_timer = std::make_unique<hg::StageTimer>("frozen field copies");
#pragma omp parallel sections   // ~0.8 GB of host copies on a 16M-cell mesh: one thread per field
  {
#pragma omp section
    fr.mann_ref.assign(f->ManningN_cells, f->ManningN_cells + N);
#pragma omp section
    fr.zb_ref.assign(f->zb_cells, f->zb_cells + N);
#pragma omp section
    fr.S0_ref.assign(f->S0_cells, f->S0_cells + 2 * N);
#pragma omp section
    fr.zbg_ghost.assign(f->zb_ghost, f->zb_ghost + B);
  }
  fr.Qin.assign(f->inletQ_TotalQ, f->inletQ_TotalQ + ctx->n_inletq);
  fr.wse.assign(f->exitH_WSE, f->exitH_WSE + ctx->n_exith);
  if (f->matID_cells) {
    x->matid_ref.resize(N);
    const int64_t nmat = std::max<int64_t>(f->n_mat, 1);
    int bad_id = 0;
#pragma omp parallel for schedule(static) reduction(| : bad_id)
    for (int64_t i = 0; i < N; ++i) {
      bad_id |= (f->matID_cells[i] < 0 || f->matID_cells[i] >= nmat) ? 1 : 0;
      x->matid_ref[i] = (int32_t)f->matID_cells[i];
    }
    if (bad_id) { ctx->err = "matID_cells out of range"; return HG_ERR_ARG; }
  }
  frozen_timer.reset();
  const hg::BcHost& h = ctx->bch;
  const size_t npar = (size_t)std::max<int64_t>(std::max<int64_t>(N, ctx->n_mat), std::max<int64_t>(ctx->n_inletq, 1));
  std::vector<double> hstill, area;   // host copies only where a table is uploaded in reference order
  if (ctx->opt.path == 1) hstill.assign(f->hstill, f->hstill + N);
  area.assign(m->cell_areas, m->cell_areas + N);

  if (ctx->opt.path == 1) {
    hg::PlainDev& p = ctx->pd;
    TRY(up(ctx, p.cf_ptr, cf_ptr)); TRY(up(ctx, p.cf_nb, cf_nb));
    TRY(up(ctx, p.cf_nx, cf_nx)); TRY(up(ctx, p.cf_ny, cf_ny)); TRY(up(ctx, p.cf_len, cf_len));
    TRY(up(ctx, p.area, area)); TRY(up(ctx, p.hstill, hstill));
    TRY(al(ctx, p.zb, N)); TRY(al(ctx, p.S0x, N)); TRY(al(ctx, p.S0y, N)); TRY(al(ctx, p.mann, N));
    TRY(up(ctx, p.matid, x->matid_ref));
    TRY(up(ctx, p.bc_type, h.type)); TRY(up(ctx, p.bc_group, h.group)); TRY(up(ctx, p.bc_ghost, h.ghost));
    TRY(up(ctx, p.bc_cell, h.cell_ref)); TRY(up(ctx, p.inlet_ptr, h.inlet_ptr));
    TRY(up(ctx, p.bc_nx, h.nx)); TRY(up(ctx, p.bc_ny, h.ny)); TRY(up(ctx, p.bc_l53, h.l53)); TRY(up(ctx, p.bc_l23, h.l23));
    std::vector<double> hg_g(f->hstill_ghost, f->hstill_ghost + B);
    TRY(up(ctx, p.hstill_g, hg_g)); TRY(al(ctx, p.zb_g, B));
    TRY(al(ctx, p.gh, B)); TRY(al(ctx, p.gqx, B)); TRY(al(ctx, p.gqy, B)); TRY(al(ctx, p.gxi, B));
    TRY(al(ctx, p.Qin, ctx->n_inletq)); TRY(al(ctx, p.wse, ctx->n_exith));
    TRY(al(ctx, p.Q, 3 * N)); TRY(al(ctx, p.dQ, 3 * N)); TRY(al(ctx, p.params, npar)); TRY(al(ctx, p.err, 1));
  } else {
    TRY(hg::build_tiles(ctx, m, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face));
    hg::FusedHost& fh = ctx->fh;
    hg::FusedDev& d = ctx->fd;
    const size_t Ns = (size_t)fh.Ns;
    hg::StageTimer up_timer("device tables (alloc+H2D)");
    TRY(up(ctx, d.perm, fh.perm)); TRY(up(ctx, d.iperm, fh.iperm)); TRY(up(ctx, d.tile_desc, fh.tile_desc));
    TRY(up(ctx, d.halo, fh.halo)); TRY(up(ctx, d.bface_e, fh.bface_e));
    TRY(up(ctx, d.face_lr, fh.face_lr)); TRY(up(ctx, d.cf_idx, fh.cf_idx));
    TRY(up(ctx, d.face_nx, fh.face_nx)); TRY(up(ctx, d.face_ny, fh.face_ny)); TRY(up(ctx, d.face_len, fh.face_len));
    TRY(upN(ctx, d.area, permuted(m->cell_areas, fh.perm), Ns)); TRY(upN(ctx, d.hstill, permuted(f->hstill, fh.perm), Ns));
    TRY(al(ctx, d.zb, Ns)); TRY(al(ctx, d.S0x, Ns)); TRY(al(ctx, d.S0y, Ns)); TRY(al(ctx, d.mann, Ns));
    if (!x->matid_ref.empty()) {
      auto mid = permuted(x->matid_ref.data(), fh.perm);
      mid.resize((size_t)fh.n_tiles * fh.T, 0);   // whole tiles: the VJP kernel prefetches the next tile's block into L2
      TRY(up(ctx, d.matid, mid));
      if (ctx->n_mat <= 255) {
        std::vector<uint8_t> m8(mid.size());
        for (size_t q = 0; q < mid.size(); ++q) m8[q] = (uint8_t)mid[q];
        TRY(up(ctx, d.matid8, m8));
      }
    }
    std::vector<int32_t> bc_cell(B);
    std::vector<double> bhst(B);
    for (int64_t e = 0; e < B; ++e) { bc_cell[e] = fh.iperm[h.cell_ref[e]]; bhst[e] = h.hstill_g[e]; }
    TRY(up(ctx, d.bc_type, h.type)); TRY(up(ctx, d.bc_group, h.group)); TRY(up(ctx, d.bc_cell, bc_cell));
    TRY(up(ctx, d.bc_nx, h.nx)); TRY(up(ctx, d.bc_ny, h.ny)); TRY(up(ctx, d.bc_l53, h.l53)); TRY(up(ctx, d.bc_l23, h.l23));
    TRY(up(ctx, d.bc_hstill, bhst)); TRY(al(ctx, d.bc_zb, B)); TRY(up(ctx, d.inlet_ptr, h.inlet_ptr));
    TRY(al(ctx, d.inlet_coef, std::max<int64_t>(ctx->n_inletq, 1)));
    TRY(al(ctx, d.Qin, ctx->n_inletq)); TRY(al(ctx, d.wse, ctx->n_exith));
    TRY(al(ctx, d.Q, 3 * Ns)); TRY(al(ctx, d.Q2, 3 * Ns)); TRY(al(ctx, d.dQ, 3 * Ns)); TRY(al(ctx, d.stage, 3 * N));
    TRY(al(ctx, d.params, npar)); TRY(al(ctx, d.err, 1));
    TRY(up(ctx, d.tile_order, fh.tile_order));
    TRY(up(ctx, d.band_order, fh.band_order)); TRY(up(ctx, d.comm_order, fh.comm_order));
    if (fh.n_chunks > 1) {
      TRY(al(ctx, d.stage_out, 3 * N));
      CK(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
      CK(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
      ctx->ev_in.resize(fh.n_chunks); ctx->ev_cmp.resize(fh.n_chunks);
      for (int k = 0; k < fh.n_chunks; ++k) {
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&ctx->ev_cmp[k], cudaEventDisableTiming));
      }
    }
    TRY(up(ctx, d.halo_off, h.halo_off)); TRY(up(ctx, d.halo_cnt, h.halo_cnt));
    TRY(al(ctx, d.halo_send, std::max<int64_t>(6 * ctx->n_halo_entries, 1)));
    TRY(al(ctx, d.halo_recv, std::max<int64_t>(6 * ctx->n_halo_entries, 1)));
    // adjoint buffers
    TRY(al(ctx, d.lam, 3 * Ns)); TRY(al(ctx, d.Qbar, 3 * Ns)); TRY(al(ctx, d.nbar, Ns)); TRY(al(ctx, d.s0bar, 2 * Ns));
    TRY(al(ctx, d.pbar, npar)); TRY(al(ctx, d.ent_c, std::max<int64_t>(B, 1))); TRY(al(ctx, d.ent_n, std::max<int64_t>(B, 1)));
    TRY(al(ctx, d.ent_z, std::max<int64_t>(B, 1))); TRY(al(ctx, d.ent_h, std::max<int64_t>(B, 1)));
    TRY(al(ctx, d.Qinbar, std::max<int64_t>(ctx->n_inletq, 1))); TRY(al(ctx, d.inlet_A, std::max<int64_t>(ctx->n_inletq, 1)));
    {
      std::vector<int32_t> bcell_int(h.bcell_ref.size());
      for (size_t q = 0; q < bcell_int.size(); ++q) bcell_int[q] = fh.iperm[h.bcell_ref[q]];
      TRY(up(ctx, d.bcell, bcell_int)); TRY(up(ctx, d.bcell_ref, h.bcell_ref));
      TRY(up(ctx, d.bcell_ptr, h.bcell_ptr)); TRY(up(ctx, d.bcell_ent, h.bcell_ent));
    }
    // the plain CSR in reference order is kept for update_bed_data when zb is the active parameter
    hg::StageTimer csr_timer("plain CSR (alloc+H2D)");
    hg::PlainDev& p = ctx->pd;
    TRY(up(ctx, p.cf_ptr, cf_ptr)); TRY(up(ctx, p.cf_nb, cf_nb));
    TRY(up(ctx, p.cf_nx, cf_nx)); TRY(up(ctx, p.cf_ny, cf_ny)); TRY(up(ctx, p.cf_len, cf_len));
    TRY(up(ctx, p.area, area)); TRY(up(ctx, p.bc_cell, h.cell_ref)); TRY(up(ctx, p.cf_rev, h.cf_rev));
    TRY(hg::fused_prepare(ctx));
    TRY(hg::fused_vjp_prepare(ctx, hg::fused_cfg_id(ctx)));
  }
  {
    hg::StageTimer fields_timer("fields (permute+H2D)");
    TRY(upload_fields(ctx));
  }
  return HG_OK;
}

int hg_create(hg_ctx** out, const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f, const hg_options* opt) {
  if (!out) { set_global_err("hg_create: out is NULL"); return HG_ERR_ARG; }
  *out = nullptr;
  hg_ctx* ctx = new hg_ctx();
  if (opt) ctx->opt = *opt; else hg_default_options(&ctx->opt);
  if (ctx->opt.strict) ctx->opt.path = 1;  // the strict build IS the plain path
  int rc = create_impl(ctx, m, b, f);
  if (rc != HG_OK) {
    set_global_err(ctx->err);
    hg_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return HG_OK;
}

int hg_set_fields(hg_ctx* ctx, const double* mann, const double* zb, const double* zbg, const double* S0,
                  const double* Qin, const double* wse) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg_ctx* x = ext(ctx);
  Frozen& fr = x->fr;
  if (mann) fr.mann_ref.assign(mann, mann + ctx->N);
  if (zb) fr.zb_ref.assign(zb, zb + ctx->N);
  if (zbg) fr.zbg_ghost.assign(zbg, zbg + ctx->B);
  if (S0) fr.S0_ref.assign(S0, S0 + 2 * ctx->N);
  if (Qin) fr.Qin.assign(Qin, Qin + ctx->n_inletq);
  if (wse) fr.wse.assign(wse, wse + ctx->n_exith);
  x->last_active = -1;  // force re-binding on the next call
  x->last_params.clear();
  return upload_fields(ctx);
}

int hg_set_manning_function(hg_ctx* ctx, int32_t type, const double* params, const double* ks_cells) {
  if (!ctx) return HG_ERR_ARG;
  if (type < HG_MANNING_CONSTANT || type > HG_MANNING_H_UMAG_KS) { ctx->err = "hg_set_manning_function: unknown type"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::MannFn m;
  m.type = type;
  if (type != HG_MANNING_CONSTANT) {
    if (!params) { ctx->err = "hg_set_manning_function: params is NULL"; return HG_ERR_ARG; }
    m.n_lower = params[0]; m.n_upper = params[1]; m.k = params[2]; m.h_mid = params[3];
    if (type != HG_MANNING_H_UMAG_KS && !(m.k > 0.0)) { ctx->err = "hg_set_manning_function: k must be positive"; return HG_ERR_ARG; }   // process_ManningN_2D.jl:141,158,175
    if (type == HG_MANNING_SIGMOID && !(m.h_mid > 0.0)) { ctx->err = "hg_set_manning_function: h_mid must be positive"; return HG_ERR_ARG; }   // :176
    if (type == HG_MANNING_H_UMAG_KS && !ks_cells) { ctx->err = "hg_set_manning_function: ks_cells is NULL"; return HG_ERR_ARG; }
    if (ctx->active == HG_PARAM_MANNING) { ctx->err = "hg_set_manning_function: ManningN is the active parameter"; return HG_ERR_ARG; }
    if (ctx->ude_set) { ctx->err = "hg_set_manning_function: a UDE model is set"; return HG_ERR_ARG; }
    const int64_t N = ctx->N;
    std::vector<double> ks(N, 1.0);
    if (ks_cells) ks.assign(ks_cells, ks_cells + N);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->opt.path == 1) {
      TRY(up(ctx, ctx->pd.ks, ks));
    } else {
      TRY(upN(ctx, ctx->fd.ks, permuted(ks.data(), ctx->fh.perm), (size_t)ctx->fh.Ns));
    }
  }
  ctx->mfn = m;
  return HG_OK;
}

int hg_set_ude_model(hg_ctx* ctx, const hg_ude_desc* desc, const double* ks_cells) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (!desc) {   // clear: un-bind theta (restores the frozen ManningN_cells)
    if (ctx->active == HG_PARAM_UDE) TRY(bind_params(ctx, nullptr, 0, HG_PARAM_NONE));
    ctx->ude_set = false;
    return HG_OK;
  }
  if (ctx->opt.path == 1) { ctx->err = "hg_set_ude_model needs the fused path (strict = 0)"; return HG_ERR_ARG; }
  if (ctx->mfn.type != HG_MANNING_CONSTANT) { ctx->err = "hg_set_ude_model: a variable Manning's n closure is set"; return HG_ERR_ARG; }
  if (ctx->active != HG_PARAM_NONE && ctx->active != HG_PARAM_UDE) { ctx->err = "hg_set_ude_model: another parameter is active"; return HG_ERR_ARG; }
  if (ctx->ens_members > 0) { ctx->err = "hg_set_ude_model: not available for ensembles"; return HG_ERR_ARG; }
  hg::ude::Model m;
  hg::ude::ThetaMap map;
  if (const char* why = hg::ude::make_model(desc, m, map)) { ctx->err = std::string("hg_set_ude_model: ") + why; return HG_ERR_ARG; }
  if (m.ln_mode == HG_LN_WHOLE_ARRAY && ctx->n_halo > 0) {
    ctx->err = "hg_set_ude_model: HG_LN_WHOLE_ARRAY needs statistics over all ranks; not available on multi-rank contexts";
    return HG_ERR_ARG;
  }
  if (m.n_in == 3 && !ks_cells) { ctx->err = "hg_set_ude_model: ks_cells is NULL"; return HG_ERR_ARG; }
  if (ctx->active == HG_PARAM_UDE) TRY(bind_params(ctx, nullptr, 0, HG_PARAM_NONE));   // theta of the previous model is void
  if (m.n_in == 3) {
    std::vector<double> ks(ks_cells, ks_cells + ctx->N);
    TRY(upN(ctx, ctx->fd.ks, permuted(ks.data(), ctx->fh.perm), (size_t)ctx->fh.Ns));
  }
  ctx->ude = m;
  ctx->ude_map = map;
  ctx->ude_user_params = desc->n_params;
  TRY(hg::ude_prepare(ctx));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ude_set = true;
  return HG_OK;
}

int hg_sync(hg_ctx* ctx) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

int hg_set_params(hg_ctx* ctx, const double* params, int64_t np, int32_t active) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return bind_params(ctx, params, np, active);
}

int hg_set_state(hg_ctx* ctx, const double* Q) {
  if (!ctx || !Q) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  const int64_t N = ctx->N;
  if (ctx->opt.path == 1) {
    CK(ctx, cudaMemcpyAsync(ctx->pd.Q.p, Q, 3 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CK(ctx, cudaMemcpyAsync(ctx->fd.stage.p, Q, 3 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRY(hg::fused_permute(ctx, true, ctx->fd.stage.p, ctx->fd.Q.p));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->state_set = true;
  ctx->state_gen++;
  ctx->ab3_step = 1;   // a new state voids the multistep history of hg_step_ab3
  return HG_OK;
}

static int download3(hg_ctx* ctx, const double* d_internal_or_ref, double* host) {
  const int64_t N = ctx->N;
  if (ctx->opt.path == 1) {
    CK(ctx, cudaMemcpyAsync(host, d_internal_or_ref, 3 * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    TRY(hg::fused_permute(ctx, false, d_internal_or_ref, ctx->fd.stage.p));
    CK(ctx, cudaMemcpyAsync(host, ctx->fd.stage.p, 3 * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}

int hg_get_state(hg_ctx* ctx, double* Q) {
  if (!ctx || !Q) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_get_state: no resident state (call hg_set_state first)"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return download3(ctx, ctx->opt.path == 1 ? ctx->pd.Q.p : ctx->fd.Q.p, Q);
}

int hg_rhs_resident(hg_ctx* ctx) {
  if (!ctx) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_rhs_resident: no resident state"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  if (ctx->opt.path == 1) return hg::plain_rhs(ctx, ctx->pd.Q.p, ctx->pd.dQ.p);
  return hg::fused_rhs(ctx, ctx->fd.Q.p, ctx->fd.dQ.p, false, 0.0);
}

int hg_rhs_resident_phase(hg_ctx* ctx, int32_t phase) {
  if (!ctx || phase < 0 || phase > 2) return HG_ERR_ARG;
  if (phase == 0) return hg_rhs_resident(ctx);
  if (ctx->opt.path == 1) { ctx->err = "hg_rhs_resident_phase needs the fused path"; return HG_ERR_ARG; }
  if (!ctx->state_set) { ctx->err = "hg_rhs_resident_phase: no resident state"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return hg::fused_rhs_phase(ctx, ctx->fd.Q.p, ctx->fd.dQ.p, phase);
}

int hg_get_rhs(hg_ctx* ctx, double* dQ) {
  if (!ctx || !dQ) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(download3(ctx, ctx->opt.path == 1 ? ctx->pd.dQ.p : ctx->fd.dQ.p, dQ));
  return check_err_flag(ctx);
}

// ---- chunk geometry of the host-buffer pipeline (tables: build_tiles in hg_host.cpp, fh.chunk_cells = nominal rows per chunk)
// Rows [r0, r1) of chunk c for the component whose HOST address is hrow: the boundaries are shifted per component so that every
// host address a copy starts at is a multiple of 256 bytes (the caller's [3N] layout gives the components arbitrary offsets).
static inline void chunk_rows(const double* hrow, int c, int K, int64_t csz, int64_t N, int64_t& r0, int64_t& r1) {
  const int64_t sh = (int64_t)((reinterpret_cast<uintptr_t>(hrow) >> 3) & (uintptr_t)(hg::kPipeAlign - 1));
  r0 = c == 0 ? 0 : std::min<int64_t>(N, std::max<int64_t>(0, (int64_t)c * csz - sh));
  r1 = c == K - 1 ? N : std::min<int64_t>(N, std::max<int64_t>(0, (int64_t)(c + 1) * csz - sh));
  if (r0 > r1) r0 = r1;
}
// chunk c of a [3][N] vector between `host` and the device staging buffer (same row offsets on both sides)
static cudaError_t copy3_chunk(double* dst, const double* src, const double* host, int64_t N, int c, int K, int64_t csz,
                               cudaMemcpyKind kind, cudaStream_t s) {
  for (int q = 0; q < 3; ++q) {
    int64_t r0, r1;
    chunk_rows(host + q * N, c, K, csz, N, r0, r1);
    if (r1 <= r0) continue;
    const cudaError_t e = cudaMemcpyAsync(dst + q * N + r0, src + q * N + r0, (size_t)(r1 - r0) * 8, kind, s);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
// rows every component has landed once chunk c has arrived, whatever the shifts: [0, landed_rows(c))
static inline int64_t landed_rows(int c, int K, int64_t csz, int64_t N) {
  return c < 0 ? 0 : c >= K - 1 ? N : std::max<int64_t>(0, std::min<int64_t>(N, (int64_t)(c + 1) * csz - hg::kPipeAlign));
}
// rows result chunk c may take, whatever the shifts: [result_lo(c), result_hi(c)) (neighbouring chunks overlap by kPipeAlign rows;
// those are gathered twice, with the same finished values)
static inline int64_t result_lo(int c, int64_t csz, int64_t N) { return std::min<int64_t>(N, std::max<int64_t>(0, (int64_t)c * csz - hg::kPipeAlign)); }
static inline int64_t result_hi(int c, int K, int64_t csz, int64_t N) { return c == K - 1 ? N : std::min<int64_t>(N, (int64_t)(c + 1) * csz); }

// Host-buffer RHS as a three-stream pipeline: the state arrives over PCIe in reference-order chunks (s_in); after each
// chunk the compute stream scatters it into the internal order and runs every tile whose cells and halo have landed;
// finished chunks of dQdt are gathered back and leave on s_out while later chunks are still arriving.  PCIe is full
// duplex, so a call costs ~max(H2D, D2H) instead of their sum.  (Pinned host buffers are needed for the copies to be
// asynchronous; pageable buffers still work, serialised by the driver.)
static int rhs_pipelined(hg_ctx* ctx, const double* Q, double* dQdt) {
  hg::FusedDev& d = ctx->fd;
  const hg::FusedHost& fh = ctx->fh;
  const int K = fh.n_chunks;
  const int64_t N = ctx->N, csz = fh.chunk_cells;
  cudaStream_t sc = ctx->stream;
  CK(ctx, cudaStreamSynchronize(sc));
  // HG_DEBUG_PIPE=1: when did every chunk land, every stage finish, every result chunk leave (ms after the first copy started)
  static const bool trace = std::getenv("HG_DEBUG_PIPE") != nullptr;
  std::vector<cudaEvent_t> tev;
  if (trace) {
    tev.resize(3 * (size_t)K + 1);
    for (auto& e : tev) CK(ctx, cudaEventCreate(&e));
    CK(ctx, cudaEventRecord(tev[3 * (size_t)K], ctx->s_in));
  }
  for (int c = 0; c < K; ++c) {
    CK(ctx, copy3_chunk(d.stage.p, Q, Q, N, c, K, csz, cudaMemcpyHostToDevice, ctx->s_in));
    CK(ctx, cudaEventRecord(ctx->ev_in[c], ctx->s_in));
    if (trace) CK(ctx, cudaEventRecord(tev[c], ctx->s_in));
  }
  for (int s = 0; s < K; ++s) {
    const int64_t r0 = landed_rows(s - 1, K, csz, N), r1 = landed_rows(s, K, csz, N);
    CK(ctx, cudaStreamWaitEvent(sc, ctx->ev_in[s], 0));
    TRY(hg::fused_permute_range(ctx, true, d.stage.p, d.Q.p, r0, r1));
    const bool band = s == K - 1 && hg_comm_ready(ctx);   // multi-rank: the whole state has landed -- push the cut cells, the last stage holds the band
    if (band) { TRY(hg::comm_push(ctx, d.Q.p, nullptr)); ctx->comm->pushed = false; }
    if (s == K - 1 && ctx->n_inletq > 0) hg::fused_inlet_coef(ctx, d.Q.p);
    TRY(hg::fused_rhs_tiles(ctx, d.Q.p, d.dQ.p, fh.stage_ptr[s], fh.stage_ptr[s + 1] - fh.stage_ptr[s], band));
    if (trace) CK(ctx, cudaEventRecord(tev[K + s], sc));
    for (int c = 0; c < K; ++c) {
      if (fh.chunk_done[c] != s) continue;
      const int64_t q0 = result_lo(c, csz, N), q1 = result_hi(c, K, csz, N);
      if (q1 <= q0) continue;
      TRY(hg::fused_permute_range(ctx, false, d.dQ.p, d.stage_out.p, q0, q1));
      CK(ctx, cudaEventRecord(ctx->ev_cmp[c], sc));
      CK(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[c], 0));
      CK(ctx, copy3_chunk(dQdt, d.stage_out.p, dQdt, N, c, K, csz, cudaMemcpyDeviceToHost, ctx->s_out));
      if (trace) CK(ctx, cudaEventRecord(tev[2 * (size_t)K + c], ctx->s_out));
    }
  }
  CK(ctx, cudaStreamSynchronize(ctx->s_out));
  CK(ctx, cudaStreamSynchronize(sc));
  if (trace) {
    fprintf(stderr, "[hg pipe] rhs, %d chunks: chunk / landed / stage done / result left (after stage) [ms]\n", K);
    for (int c = 0; c < K; ++c) {
      float a = 0, b = 0, o = 0;
      cudaEventElapsedTime(&a, tev[3 * (size_t)K], tev[c]);
      cudaEventElapsedTime(&b, tev[3 * (size_t)K], tev[K + c]);
      if (result_hi(c, K, csz, N) > result_lo(c, csz, N)) cudaEventElapsedTime(&o, tev[3 * (size_t)K], tev[2 * (size_t)K + c]);
      fprintf(stderr, "[hg pipe] %3d %8.3f %8.3f %8.3f (%d)\n", c, a, b, o, fh.chunk_done[c]);
    }
    for (auto& e : tev) cudaEventDestroy(e);
  }
  ctx->state_set = true;
  ctx->state_gen++;
  ctx->ab3_step = 1;
  return check_err_flag(ctx);
}

int hg_rhs(hg_ctx* ctx, const double* Q, const double* params, int64_t np, int32_t active, double t, double* dQdt) {
  (void)t;  // unused, exactly like the reference (semi_discretize_swe_2D.jl:18)
  if (!ctx || !Q || !dQdt) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  // (the UDE network may need whole-array statistics of the state before any tile can run: no chunked pipeline)
  if (ctx->opt.path != 1 && ctx->fh.n_chunks > 1 && (ctx->n_halo == 0 || hg_comm_ready(ctx)) && ctx->active != HG_PARAM_UDE) return rhs_pipelined(ctx, Q, dQdt);
  TRY(hg_set_state(ctx, Q));
  TRY(hg_rhs_resident(ctx));
  return hg_get_rhs(ctx, dQdt);
}

static int set_lambda(hg_ctx* ctx, const double* lam) {
  const int64_t N = ctx->N;
  CK(ctx, cudaMemcpyAsync(ctx->fd.stage.p, lam, 3 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
  TRY(hg::fused_permute(ctx, true, ctx->fd.stage.p, ctx->fd.lam.p));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->lam_set = true;
  return HG_OK;
}

// Host-buffer VJP through the same three-stream pipeline as rhs_pipelined: chunks of Q and lambda arrive on s_in, the
// tiles whose cells and halo have landed run on the compute stream, finished chunks of Qbar leave on s_out.  The
// boundary-wide inlet coupling and the parameter reductions run after the last stage (the tiles they touch are
// last-stage tiles, see build_tiles), before the chunks holding those cells leave.
static int vjp_pipelined(hg_ctx* ctx, const double* Q, const double* lambda, double* Qbar) {
  hg::FusedDev& d = ctx->fd;
  const hg::FusedHost& fh = ctx->fh;
  const int K = fh.n_chunks;
  const int64_t N = ctx->N, csz = fh.chunk_cells;
  cudaStream_t sc = ctx->stream;
  if (d.stage_lam.n < (size_t)(3 * N)) CK(ctx, d.stage_lam.alloc(3 * N));
  const int cfg = hg::fused_cfg_id(ctx);
  CK(ctx, cudaStreamSynchronize(sc));
  for (int c = 0; c < K; ++c) {
    if (Q) CK(ctx, copy3_chunk(d.stage.p, Q, Q, N, c, K, csz, cudaMemcpyHostToDevice, ctx->s_in));   // NULL: the resident state
    CK(ctx, copy3_chunk(d.stage_lam.p, lambda, lambda, N, c, K, csz, cudaMemcpyHostToDevice, ctx->s_in));
    CK(ctx, cudaEventRecord(ctx->ev_in[c], ctx->s_in));
  }
  for (int s = 0; s < K; ++s) {
    const int64_t r0 = landed_rows(s - 1, K, csz, N), r1 = landed_rows(s, K, csz, N);
    CK(ctx, cudaStreamWaitEvent(sc, ctx->ev_in[s], 0));
    if (Q) TRY(hg::fused_permute_range(ctx, true, d.stage.p, d.Q.p, r0, r1));
    TRY(hg::fused_permute_range(ctx, true, d.stage_lam.p, d.lam.p, r0, r1));
    const bool band = s == K - 1 && hg_comm_ready(ctx);
    if (band) { TRY(hg::comm_push(ctx, d.Q.p, d.lam.p)); ctx->comm->pushed = false; }
    if (s == K - 1 && ctx->n_inletq > 0) hg::fused_inlet_coef(ctx, d.Q.p);
    TRY(hg::fused_vjp_tiles(ctx, cfg, d.Q.p, d.lam.p, d.Qbar.p, d.tile_order.p, fh.stage_ptr[s], fh.stage_ptr[s + 1] - fh.stage_ptr[s], band ? 2 : 0));
    if (s == K - 1) TRY(hg::fused_vjp_finish(ctx, d.Q.p, d.Qbar.p));
    for (int c = 0; c < K; ++c) {
      if (fh.chunk_done[c] != s) continue;
      const int64_t q0 = result_lo(c, csz, N), q1 = result_hi(c, K, csz, N);
      if (q1 <= q0) continue;
      TRY(hg::fused_permute_range(ctx, false, d.Qbar.p, d.stage_out.p, q0, q1));
      CK(ctx, cudaEventRecord(ctx->ev_cmp[c], sc));
      CK(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_cmp[c], 0));
      CK(ctx, copy3_chunk(Qbar, d.stage_out.p, Qbar, N, c, K, csz, cudaMemcpyDeviceToHost, ctx->s_out));
    }
  }
  CK(ctx, cudaStreamSynchronize(ctx->s_out));
  if (Q) { ctx->state_set = true; ctx->state_gen++; ctx->ab3_step = 1; }
  ctx->lam_set = true;
  return HG_OK;
}

int hg_rhs_vjp(hg_ctx* ctx, const double* Q, const double* params, int64_t np, int32_t active, double t, const double* lambda,
               double* Qbar, double* pbar, double* ncell_bar) {
  (void)t;
  if (!ctx || !lambda || !Qbar) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_rhs_vjp needs the fused path (strict = 0)"; return HG_ERR_ARG; }
  // Q = NULL: the pullback of a forward call whose state is still on the device (Zygote's pullback closes over the primal
  // of its forward pass; hg_state_generation tells the caller whether anything has moved the state since)
  if (!Q && !ctx->state_set) { ctx->err = "hg_rhs_vjp: Q is NULL and no state is resident (call hg_rhs or hg_set_state first)"; return HG_ERR_STATE; }
  TRY(no_closure(ctx, "hg_rhs_vjp"));
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  if (ctx->active != HG_PARAM_NONE && !pbar) { ctx->err = "hg_rhs_vjp: pbar is NULL"; return HG_ERR_ARG; }
  hg::FusedDev& d = ctx->fd;
  if (ctx->fh.n_chunks > 1 && (ctx->n_halo == 0 || hg_comm_ready(ctx)) && ctx->active != HG_PARAM_UDE) {
    TRY(vjp_pipelined(ctx, Q, lambda, Qbar));
  } else {
    if (Q) TRY(hg_set_state(ctx, Q));
    TRY(set_lambda(ctx, lambda));
    TRY(hg::fused_vjp(ctx, hg::fused_cfg_id(ctx), d.Q.p, d.lam.p, d.Qbar.p));
    TRY(download3(ctx, d.Qbar.p, Qbar));
  }
  if (ctx->active != HG_PARAM_NONE)
    CK(ctx, cudaMemcpyAsync(pbar, d.pbar.p, ctx->n_params * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (ncell_bar) {
    TRY(hg::fused_nbar_to_ref(ctx, d.stage.p));
    CK(ctx, cudaMemcpyAsync(ncell_bar, d.stage.p, ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

// Forward mode: dQdt (optional) and dQdt_dot = J_Q(Q, p) v + J_p(Q, p) pdot, what a ForwardDiff.Dual pass through swe_2d_rhs
// carries (one partial per call).  Runs on the plain tables (reference evaluation order, hg_jvp.cu): the context must be
// created with strict = 1 / path = 1.  The state-dependent Manning closures and the UDE network have no forward mode here.
// Forward mode on a fused (non-strict) context: the tile kernel of hg_fjvp.cu, K directions per launch.  Host vectors in the
// reference's order; V [K][3N], Pdot [K][np] or NULL, dQdt [3N] or NULL, JV [K][3N].
static int fused_jvp_host(hg_ctx* ctx, const double* Q, int64_t K, const double* V, const double* Pdot, double* dQdt, double* JV) {
  hg::FusedDev& d = ctx->fd;
  const int64_t N = ctx->N;
  const size_t Ns3 = 3 * (size_t)ctx->fh.Ns;
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  if (d.j_V.n < (size_t)K * Ns3) { TRY(al(ctx, d.j_V, (size_t)K * Ns3)); TRY(al(ctx, d.j_out, (size_t)K * Ns3)); }
  TRY(hg_set_state(ctx, Q));
  for (int64_t k = 0; k < K; ++k) {
    CK(ctx, cudaMemcpyAsync(d.stage.p, V + k * 3 * N, 3 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRY(hg::fused_permute(ctx, true, d.stage.p, d.j_V.p + k * Ns3));
  }
  const double* d_pdot = nullptr;
  if (Pdot && npar > 0) {
    if (d.j_pdot.n < (size_t)(K * npar)) TRY(al(ctx, d.j_pdot, (size_t)(K * npar)));
    CK(ctx, cudaMemcpyAsync(d.j_pdot.p, Pdot, (size_t)(K * npar) * 8, cudaMemcpyHostToDevice, ctx->stream));
    d_pdot = d.j_pdot.p;
  }
  TRY(hg::fused_jvp(ctx, hg::fused_cfg_id(ctx), d.Q.p, d.j_V.p, d_pdot, dQdt ? d.dQ.p : nullptr, d.j_out.p, K));
  if (dQdt) TRY(download3(ctx, d.dQ.p, dQdt));
  for (int64_t k = 0; k < K; ++k) TRY(download3(ctx, d.j_out.p + k * Ns3, JV + k * 3 * N));
  return check_err_flag(ctx);
}

int hg_rhs_jvp(hg_ctx* ctx, const double* Q, const double* params, int64_t np, int32_t active, double t, const double* v,
               const double* pdot, double* dQdt, double* dQdt_dot) {
  (void)t;
  if (!ctx || !Q || !v || !dQdt_dot) return HG_ERR_ARG;
  TRY(no_closure(ctx, "hg_rhs_jvp"));
  if (active == HG_PARAM_UDE) { ctx->err = "hg_rhs_jvp: the UDE network has no forward mode"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  if (ctx->opt.path != 1) return fused_jvp_host(ctx, Q, 1, v, pdot, dQdt, dQdt_dot);   // fused tile kernel (hg_fjvp.cu)
  hg::PlainDev& p = ctx->pd;
  const size_t n3 = 3 * (size_t)ctx->N;
  if (p.V.n != n3) { TRY(al(ctx, p.V, n3)); TRY(al(ctx, p.dQd, n3)); }
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  const double* d_pdot = nullptr;
  if (pdot && npar > 0) {
    if (p.pdot.n < (size_t)npar) TRY(al(ctx, p.pdot, (size_t)npar));
    CK(ctx, cudaMemcpyAsync(p.pdot.p, pdot, npar * 8, cudaMemcpyHostToDevice, ctx->stream));
    d_pdot = p.pdot.p;
  }
  CK(ctx, cudaMemcpyAsync(p.Q.p, Q, n3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(p.V.p, v, n3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  ctx->state_set = true;
  ctx->state_gen++;
  TRY(hg::plain_jvp(ctx, p.Q.p, p.V.p, d_pdot, dQdt ? p.dQ.p : nullptr, p.dQd.p));
  if (dQdt) CK(ctx, cudaMemcpyAsync(dQdt, p.dQ.p, n3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dQdt_dot, p.dQd.p, n3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

// K directions in one call (a ForwardDiff chunk: Dual{Tag, Float64, K}): Q is uploaded once, the K sweeps run back to back
// on the stream, one download.  V[K][3N], Pdot[K][n_params] or NULL, JV[K][3N].
int hg_rhs_jvp_multi(hg_ctx* ctx, const double* Q, const double* params, int64_t np, int32_t active, double t, int64_t K,
                     const double* V, const double* Pdot, double* dQdt, double* JV) {
  (void)t;
  if (!ctx || !Q || !V || !JV || K < 1) return HG_ERR_ARG;
  TRY(no_closure(ctx, "hg_rhs_jvp_multi"));
  if (active == HG_PARAM_UDE) { ctx->err = "hg_rhs_jvp_multi: the UDE network has no forward mode"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  if (ctx->opt.path != 1) return fused_jvp_host(ctx, Q, K, V, Pdot, dQdt, JV);   // fused tile kernel (hg_fjvp.cu), K directions per launch
  hg::PlainDev& p = ctx->pd;
  const size_t n3 = 3 * (size_t)ctx->N;
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  hg::DBuf<double> dV, dJ, dP;
  CK(ctx, dV.alloc((size_t)K * n3)); CK(ctx, dJ.alloc((size_t)K * n3));
  const bool with_p = Pdot && npar > 0;
  if (with_p) {
    CK(ctx, dP.alloc((size_t)(K * npar)));
    CK(ctx, cudaMemcpyAsync(dP.p, Pdot, (size_t)(K * npar) * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(ctx, cudaMemcpyAsync(p.Q.p, Q, n3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dV.p, V, (size_t)K * n3 * 8, cudaMemcpyHostToDevice, ctx->stream));
  ctx->state_set = true;
  ctx->state_gen++;
  // all K directions in one pair of launches
  TRY(hg::plain_jvp_batch(ctx, p.Q.p, dV.p, (int64_t)n3, with_p ? dP.p : nullptr, npar, dQdt ? p.dQ.p : nullptr, dJ.p, K));
  if (dQdt) CK(ctx, cudaMemcpyAsync(dQdt, p.dQ.p, n3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaMemcpyAsync(JV, dJ.p, (size_t)K * n3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

// Forward sensitivity solve: the reference's sensitivity driver on the device.  swe_2D_sensitivity.jl:34-80 wraps
// solve(prob, Tsit5(), adaptive=..., dt=dt; abstol, reltol) in ForwardDiff.jacobian, i.e. the state is a vector of Duals with
// one partial per parameter: values and partials advance together, the error estimate -- hence every step size -- includes
// the partials (DiffEqBase's norm of Dual numbers).  Here: U = [Q; dQ/dp_1; ...; dQ/dp_K] resident on the device, each Tsit5
// stage = K forward-mode sweeps (plain_jvp, the first one also delivers the values), the Dual-aware norm reduced on the
// device, the PI controller on the host (powers per hg_set_controller_pow).  Plain tables: the context needs strict = 1.
int hg_solve_tsit5_sens(hg_ctx* ctx, const double* Q0, const double* params, int64_t np, int32_t active, double t0, double t1,
                        double dt, int32_t adaptive, double abstol, double reltol, const double* t_save, int64_t n_save,
                        double* Q_save, double* Q_T, double* S, int64_t* stats) {
  if (!ctx || !Q0 || !S || !(t1 > t0) || !(dt > 0.0) || n_save < 0 || (n_save > 0 && (!t_save || !Q_save))) return HG_ERR_ARG;
  if (adaptive && (!(abstol > 0.0) || !(reltol > 0.0))) { ctx->err = "hg_solve_tsit5_sens: tolerances must be positive"; return HG_ERR_ARG; }
  // strict contexts: forward mode on the plain tables (reference operation order); fused contexts: the tile kernel of
  // hg_fjvp.cu -- the same solver around either, the augmented state in the context's own cell order
  const bool fused = ctx->opt.path != 1;
  TRY(no_closure(ctx, "hg_solve_tsit5_sens"));
  if (active != HG_PARAM_ZB && active != HG_PARAM_MANNING && active != HG_PARAM_Q) {
    ctx->err = "hg_solve_tsit5_sens: the active parameter must be zb, ManningN or Q";
    return HG_ERR_ARG;
  }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  const int64_t K = ctx->n_params;
  if (K < 1) { ctx->err = "hg_solve_tsit5_sens: no parameters"; return HG_ERR_ARG; }
  static const double A[7][6] = {
      {0, 0, 0, 0, 0, 0},
      {0.161, 0, 0, 0, 0, 0},
      {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
      {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
      {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
      {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
      {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
  static const double BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                               0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
  const double beta2 = 2.0 / 25.0, beta1 = 7.0 / 50.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1e-4;
  const int64_t n3h = 3 * ctx->N;                                   // a row on the host (reference order)
  const int64_t n3 = fused ? 3 * ctx->fh.Ns : n3h, rows = 1 + K, len = rows * n3;   // a row on the device (fused: padded internal order; the padding stays zero)
  const int cfg = fused ? hg::fused_cfg_id(ctx) : 0;
  auto fetch3 = [&](const double* d_row, double* host) -> int {     // one row of the augmented state -> host, reference order
    if (fused) return download3(ctx, d_row, host);
    CK(ctx, cudaMemcpyAsync(host, d_row, (size_t)n3h * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return HG_OK;
  };
  hg::DBuf<double> U, Unew, Y, kb[7], E, part, sum;
  CK(ctx, U.alloc((size_t)len)); CK(ctx, Unew.alloc((size_t)len)); CK(ctx, Y.alloc((size_t)len));
  for (int m = 0; m < 7; ++m) CK(ctx, kb[m].alloc((size_t)len));
  CK(ctx, cudaMemsetAsync(Unew.p, 0, (size_t)len * 8, ctx->stream)); CK(ctx, cudaMemsetAsync(Y.p, 0, (size_t)len * 8, ctx->stream));
  for (int m = 0; m < 7; ++m) CK(ctx, cudaMemsetAsync(kb[m].p, 0, (size_t)len * 8, ctx->stream));
  CK(ctx, E.alloc((size_t)(K * K)));
  CK(ctx, part.alloc((size_t)hg::sens_err_blocks(n3))); CK(ctx, sum.alloc(1));
  {
    std::vector<double> eye((size_t)(K * K), 0.0);
    for (int64_t k = 0; k < K; ++k) eye[(size_t)(k * K + k)] = 1.0;
    CK(ctx, cudaMemcpyAsync(E.p, eye.data(), eye.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  // saveat = t_save of the driver (swe_2D_sensitivity.jl:38-43): the VALUES at the save times by Tsit5's dense output, steps
  // independent of the save times (what forward_simulation_results.json holds)
  std::vector<std::pair<double, int64_t>> pending;
  for (int64_t i = 0; i < n_save; ++i) {
    if (t_save[i] == t0) std::memcpy(Q_save + (size_t)i * n3h, Q0, (size_t)n3h * 8);
    else if (t_save[i] > t0 && t_save[i] <= t1) pending.push_back({t_save[i], i});
    else { ctx->err = "hg_solve_tsit5_sens: save time outside [t0, t1]"; return HG_ERR_ARG; }
  }
  std::sort(pending.begin(), pending.end());
  size_t next_save = 0;
  static const double RI[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                  {0.0, 0.13169999999999998, -0.2234, 0.1017},
                                  {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                                  {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                                  {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                                  {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                                  {0.0, 1.5, -4.0, 2.5}};
  CK(ctx, cudaMemsetAsync(U.p, 0, (size_t)len * 8, ctx->stream));           // the partials start at zero (Q0 does not depend on p)
  if (fused) {
    CK(ctx, cudaMemcpyAsync(ctx->fd.stage.p, Q0, (size_t)n3h * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRY(hg::fused_permute(ctx, true, ctx->fd.stage.p, U.p));
    ctx->state_set = false;     // the resident state buffer is not what this solve integrates
    ctx->state_gen++;
  } else {
    CK(ctx, cudaMemcpyAsync(U.p, Q0, (size_t)n3h * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  // d/dt of the augmented state: row 0 = f(Q, p), row k = J_Q U_k + J_p e_k
  auto rhs_aug = [&](const double* u, double* du) -> int {
    // the K partials in one launch (pair): row k of the augmented state is direction k; pdot = e_k
    if (fused) return hg::fused_jvp(ctx, cfg, u, u + n3, E.p, du, du + n3, K);
    return hg::plain_jvp_batch(ctx, u, u + n3, (int64_t)n3, E.p, K, du, du + n3, K);
  };
  double* u = U.p;
  double* unew = Unew.p;
  double* k[7];
  for (int m = 0; m < 7; ++m) k[m] = kb[m].p;
  int64_t n_acc = 0, n_rej = 0, n_rhs = 0;
  double t = t0, dt_ctrl = dt, qold = qoldinit;
  ctx->last_steps.clear();
  TRY(rhs_aug(u, k[0]));
  ++n_rhs;
  while (t < t1) {
    double h = std::min(dt_ctrl, t1 - t);
    const bool clipped = h < dt_ctrl;
    if (t1 - (t + h) < 1e-12 * std::max(1.0, std::fabs(t1))) h = t1 - t;
    double coef[7];
    for (int i = 1; i < 7; ++i) {
      for (int j = 0; j < i; ++j) coef[j] = h * A[i][j];
      double* y = i < 6 ? Y.p : unew;
      TRY(hg::sens_lincomb(ctx, len, y, u, i, k, coef));
      TRY(rhs_aug(y, k[i]));
      ++n_rhs;
    }
    bool accept = true;
    double q = 1.0, q11 = 0.0, eest = 0.0;
    if (adaptive) {
      for (int m = 0; m < 7; ++m) coef[m] = h * BT[m];
      TRY(hg::sens_err_norm(ctx, n3, (int)rows, u, unew, 7, k, coef, abstol, reltol, part.p, sum.p));
      double ssum = 0.0;
      CK(ctx, cudaMemcpyAsync(&ssum, sum.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
      eest = std::sqrt(ssum / (double)(rows * n3h));
      if (!(eest == eest)) { ctx->err = "hg_solve_tsit5_sens: the error estimate is NaN"; return HG_ERR_STATE; }
      if (eest == 0.0) { q11 = 0.0; q = 1.0 / qmax; }
      else {
        q11 = ctx->controller_fastpow ? hg_fastpow(eest, beta1) : std::pow(eest, beta1);
        q = q11 / (ctx->controller_fastpow ? hg_fastpow(qold, beta2) : std::pow(qold, beta2));
        q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / gamma));
      }
      accept = eest <= 1.0;
    }
    if (accept) {
      const double tnew = (h == t1 - t) ? t1 : t + h;
      for (; next_save < pending.size() && pending[next_save].first <= tnew; ++next_save) {
        double* out = Q_save + (size_t)pending[next_save].second * n3h;
        if (pending[next_save].first == tnew) {
          TRY(fetch3(unew, out));
        } else {                                   // row 0 of u + h sum_i b_i(theta) k_i (Y is free between steps)
          const double th = (pending[next_save].first - t) / h;
          for (int m = 0; m < 7; ++m) coef[m] = h * (th * (RI[m][0] + th * (RI[m][1] + th * (RI[m][2] + th * RI[m][3]))));
          TRY(hg::sens_lincomb(ctx, n3, Y.p, u, 7, k, coef));
          TRY(fetch3(Y.p, out));
        }
      }
      std::swap(u, unew);
      std::swap(k[0], k[6]);                       // FSAL
      t = tnew;
      ++n_acc;
      ctx->last_steps.push_back(h);
      if (adaptive) {
        qold = std::max(eest, qoldinit);
        const double prop = h / q;
        dt_ctrl = clipped ? std::max(prop, dt_ctrl) : prop;
      }
    } else {
      ++n_rej;
      dt_ctrl = h / std::min(1.0 / qmin, q11 / gamma);
      if (!(dt_ctrl > 1e-14 * std::max(1.0, std::fabs(t)))) { ctx->err = "hg_solve_tsit5_sens: step size underflow"; return HG_ERR_STATE; }
    }
  }
  if (Q_T) TRY(fetch3(u, Q_T));
  for (int64_t kk = 0; kk < K; ++kk) TRY(fetch3(u + (1 + kk) * n3, S + (size_t)kk * n3h));
  if (stats) { stats[0] = n_acc; stats[1] = n_rej; stats[2] = n_rhs; }
  return check_err_flag(ctx);
}

int hg_set_lambda(hg_ctx* ctx, const double* lambda) {
  if (!ctx || !lambda) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_set_lambda needs the fused path"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return set_lambda(ctx, lambda);
}

int hg_vjp_resident(hg_ctx* ctx) {
  if (!ctx) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_vjp_resident needs the fused path"; return HG_ERR_ARG; }
  if (!ctx->state_set || !ctx->lam_set) { ctx->err = "hg_vjp_resident: state or lambda not set"; return HG_ERR_STATE; }
  TRY(no_closure(ctx, "hg_vjp_resident"));
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return hg::fused_vjp(ctx, hg::fused_cfg_id(ctx), ctx->fd.Q.p, ctx->fd.lam.p, ctx->fd.Qbar.p);
}

int hg_vjp_resident_phase(hg_ctx* ctx, int32_t phase) {
  if (!ctx || phase < 0 || phase > 2) return HG_ERR_ARG;
  if (phase == 0) return hg_vjp_resident(ctx);
  if (ctx->opt.path == 1) { ctx->err = "hg_vjp_resident_phase needs the fused path"; return HG_ERR_ARG; }
  if (!ctx->state_set || !ctx->lam_set) { ctx->err = "hg_vjp_resident_phase: state or lambda not set"; return HG_ERR_STATE; }
  TRY(no_closure(ctx, "hg_vjp_resident_phase"));
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const hg::FusedHost& fh = ctx->fh;
  const int cfg = hg::fused_cfg_id(ctx);
  if (phase == 1) {
    if (ctx->active == HG_PARAM_UDE) TRY(hg::ude_eval_n(ctx, d.Q.p));
    if (ctx->n_inletq > 0) hg::fused_inlet_coef(ctx, d.Q.p);
    return hg::fused_vjp_tiles(ctx, cfg, d.Q.p, d.lam.p, d.Qbar.p, d.band_order.p, 0, fh.n_interior_tiles);
  }
  TRY(hg::fused_vjp_tiles(ctx, cfg, d.Q.p, d.lam.p, d.Qbar.p, d.band_order.p, fh.n_interior_tiles, fh.n_tiles - fh.n_interior_tiles));
  return hg::fused_vjp_finish(ctx, d.Q.p, d.Qbar.p);
}

int hg_get_vjp(hg_ctx* ctx, double* Qbar, double* pbar, double* ncell_bar) {
  if (!ctx || !Qbar) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_get_vjp needs the fused path"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  TRY(download3(ctx, d.Qbar.p, Qbar));
  if (ctx->active != HG_PARAM_NONE && pbar)
    CK(ctx, cudaMemcpyAsync(pbar, d.pbar.p, ctx->n_params * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (ncell_bar) {
    TRY(hg::fused_nbar_to_ref(ctx, d.stage.p));
    CK(ctx, cudaMemcpyAsync(ncell_bar, d.stage.p, ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

int hg_halo_info(const hg_ctx* ctx, int64_t* n_neighbors, int64_t* n_entries) {
  if (!ctx) return HG_ERR_ARG;
  if (n_neighbors) *n_neighbors = ctx->n_halo;
  if (n_entries) *n_entries = ctx->n_halo_entries;
  return HG_OK;
}
int hg_halo_counts(const hg_ctx* ctx, int64_t* counts) {
  if (!ctx || !counts) return HG_ERR_ARG;
  for (size_t k = 0; k < ctx->bch.halo_counts.size(); ++k) counts[k] = ctx->bch.halo_counts[k];
  return HG_OK;
}
int hg_halo_buffers(hg_ctx* ctx, double** d_send, double** d_recv, int64_t* n_doubles) {
  if (!ctx) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "halo exchange needs the fused path"; return HG_ERR_ARG; }
  if (d_send) *d_send = ctx->fd.halo_send.p;
  if (d_recv) *d_recv = ctx->fd.halo_recv.p;
  if (n_doubles) *n_doubles = 6 * ctx->n_halo_entries;
  return HG_OK;
}
int hg_halo_pack(hg_ctx* ctx, int32_t with_lambda) {
  if (!ctx) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "halo exchange needs the fused path"; return HG_ERR_ARG; }
  if (!ctx->state_set || (with_lambda && !ctx->lam_set)) { ctx->err = "hg_halo_pack: state or lambda not set"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return hg::fused_halo_pack(ctx, with_lambda != 0);
}
int hg_set_stream(hg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return HG_OK;
}

// ---------------------------------------------------------------- parameter ensembles
// Multi-rank contexts (n_halo > 0): every RHS needs fresh halo states.  An entry point that loops the RHS without an
// exchange between steps / stages would read stale halo data, so it refuses unless the library owns the exchange.
static int need_exchange_guard(hg_ctx* ctx, const char* who, bool loops) {
  if (ctx->n_halo > 0 && loops && !hg_comm_ready(ctx)) {
    ctx->err = std::string(who) + ": multi-rank context without a library-owned halo exchange (hg_comm_connect) -- "
               "this entry point evaluates the RHS more than once per call and would read stale halo states";
    return HG_ERR_ARG;
  }
  return HG_OK;
}

int hg_ensemble_alloc(hg_ctx* ctx, int64_t M, int32_t per_member_manning) {
  if (!ctx || M <= 0) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "ensembles need the fused path"; return HG_ERR_ARG; }
  if (ctx->n_halo > 0) { ctx->err = "hg_ensemble_alloc: ensembles shard by member, not by domain -- multi-rank (halo) contexts are not supported"; return HG_ERR_ARG; }
  TRY(no_closure(ctx, "hg_ensemble_alloc"));
  if (ctx->ude_set) { ctx->err = "hg_ensemble_alloc: not available while a UDE model is set"; return HG_ERR_ARG; }
  if ((int64_t)ctx->fh.n_tiles * M >= ((int64_t)1 << 31)) { ctx->err = "too many members for one launch"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const size_t Ns = (size_t)ctx->fh.Ns, nI = (size_t)std::max<int64_t>(ctx->n_inletq, 1);
  TRY(al(ctx, d.ens_Q, (size_t)M * 3 * Ns)); TRY(al(ctx, d.ens_Q2, (size_t)M * 3 * Ns));
  if (per_member_manning) TRY(al(ctx, d.ens_mann, (size_t)M * Ns));
  TRY(al(ctx, d.ens_Qin, (size_t)M * nI)); TRY(al(ctx, d.ens_coef, (size_t)M * nI)); TRY(al(ctx, d.ens_A, (size_t)M * nI));
  for (int64_t m = 0; m < M && ctx->n_inletq > 0; ++m)
    CK(ctx, cudaMemcpyAsync(d.ens_Qin.p + m * ctx->n_inletq, d.Qin.p, ctx->n_inletq * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  for (int64_t m = 0; m < M && per_member_manning; ++m)
    CK(ctx, cudaMemcpyAsync(d.ens_mann.p + m * Ns, d.mann.p, Ns * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ens_members = M;
  ctx->ens_per_member_mann = per_member_manning != 0;
  return HG_OK;
}

int hg_ensemble_set_member(hg_ctx* ctx, int64_t m, const double* Q, const double* params, int64_t np, int32_t active) {
  if (!ctx || !Q || m < 0 || m >= ctx->ens_members) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const int64_t N = ctx->N, Ns = ctx->fh.Ns;
  CK(ctx, cudaMemcpyAsync(d.stage.p, Q, 3 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
  TRY(hg::fused_permute(ctx, true, d.stage.p, d.ens_Q.p + m * 3 * Ns));
  if (active == HG_PARAM_MANNING) {
    if (!ctx->ens_per_member_mann || np != ctx->n_mat || !params || ctx->matid_ref.empty()) {
      ctx->err = "hg_ensemble_set_member: ManningN members need per_member_manning, matID_cells and n_mat values";
      return HG_ERR_ARG;
    }
    CK(ctx, cudaMemcpyAsync(d.params.p, params, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* keep = d.mann.p;
    d.mann.p = d.ens_mann.p + m * Ns;            // expand the zone values straight into the member's field
    int rc = hg::fused_bind_manning(ctx, d.params.p);
    d.mann.p = keep;
    if (rc != HG_OK) return rc;
  } else if (active == HG_PARAM_Q) {
    if (np != ctx->n_inletq || !params) { ctx->err = "hg_ensemble_set_member: wrong number of inlet discharges"; return HG_ERR_ARG; }
    CK(ctx, cudaMemcpyAsync(d.ens_Qin.p + m * ctx->n_inletq, params, np * 8, cudaMemcpyHostToDevice, ctx->stream));
  } else if (active != HG_PARAM_NONE) {
    ctx->err = "hg_ensemble_set_member: only ManningN and Q vary per member";
    return HG_ERR_ARG;
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}

int hg_ensemble_step_euler(hg_ctx* ctx, double dt, int64_t nsteps) {
  if (!ctx || nsteps < 0 || ctx->ens_members <= 0) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  for (int64_t s = 0; s < nsteps; ++s) {
    TRY(hg::fused_rhs_ensemble(ctx, d.ens_Q.p, d.ens_Q2.p, true, dt));
    std::swap(d.ens_Q.p, d.ens_Q2.p);
  }
  return HG_OK;
}

int hg_ensemble_rhs(hg_ctx* ctx) {
  if (!ctx || ctx->ens_members <= 0) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  return hg::fused_rhs_ensemble(ctx, ctx->fd.ens_Q.p, ctx->fd.ens_Q2.p, false, 0.0);
}

int hg_ensemble_get_member(hg_ctx* ctx, int64_t m, int32_t what, double* out) {
  if (!ctx || !out || m < 0 || m >= ctx->ens_members) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  TRY(download3(ctx, (what ? d.ens_Q2.p : d.ens_Q.p) + m * 3 * ctx->fh.Ns, out));
  return check_err_flag(ctx);
}

int hg_time_ensemble(hg_ctx* ctx, int32_t n, double dt, float* ms) {
  if (!ctx || !ms || n <= 0 || ctx->ens_members <= 0) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  TRY(hg_ensemble_step_euler(ctx, dt, n));
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return check_err_flag(ctx);
}

int hg_step_euler(hg_ctx* ctx, double dt, int64_t nsteps) {
  if (!ctx || nsteps < 0) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_step_euler: no resident state"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  if (ctx->opt.path == 1) { ctx->err = "hg_step_euler needs the fused path (path=0)"; return HG_ERR_ARG; }
  TRY(need_exchange_guard(ctx, "hg_step_euler", nsteps > 1));
  ctx->ab3_step = 1;   // another stepper moves the state: hg_step_ab3 starts again
  ctx->state_gen++;
  hg::FusedDev& d = ctx->fd;
  for (int64_t s = 0; s < nsteps; ++s) {
    TRY(hg::fused_rhs(ctx, d.Q.p, d.Q2.p, true, dt));
    std::swap(d.Q.p, d.Q2.p);
  }
  return HG_OK;
}

int hg_step_rk4(hg_ctx* ctx, double dt, int64_t nsteps) {
  if (!ctx || nsteps < 0) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_step_rk4: no resident state"; return HG_ERR_STATE; }
  if (ctx->opt.path == 1) { ctx->err = "hg_step_rk4 needs the fused path (path=0)"; return HG_ERR_ARG; }
  TRY(need_exchange_guard(ctx, "hg_step_rk4", true));
  ctx->ab3_step = 1;   // another stepper moves the state: hg_step_ab3 starts again
  ctx->state_gen++;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const size_t n = 3 * (size_t)ctx->fh.Ns;
  if (d.rk_k.n != n) { TRY(al(ctx, d.rk_k, n)); TRY(al(ctx, d.rk_acc, n)); TRY(al(ctx, d.rk_tmp, n)); }
  for (int64_t s = 0; s < nsteps; ++s) {
    // k1 = f(Q);            tmp = Q + dt/2 k1;   acc  = dt/6 k1
    TRY(hg::fused_rhs(ctx, d.Q.p, d.rk_k.p, false, 0.0));
    TRY(hg::fused_axpy(ctx, d.rk_tmp.p, d.Q.p, d.rk_k.p, 0.5 * dt, nullptr, d.rk_acc.p, dt / 6.0));
    // k2 = f(tmp);          tmp = Q + dt/2 k2;   acc += dt/3 k2
    TRY(hg::fused_rhs(ctx, d.rk_tmp.p, d.rk_k.p, false, 0.0));
    TRY(hg::fused_axpy(ctx, d.rk_tmp.p, d.Q.p, d.rk_k.p, 0.5 * dt, d.rk_acc.p, d.rk_acc.p, dt / 3.0));
    // k3 = f(tmp);          tmp = Q + dt k3;     acc += dt/3 k3
    TRY(hg::fused_rhs(ctx, d.rk_tmp.p, d.rk_k.p, false, 0.0));
    TRY(hg::fused_axpy(ctx, d.rk_tmp.p, d.Q.p, d.rk_k.p, dt, d.rk_acc.p, d.rk_acc.p, dt / 3.0));
    // k4 = f(tmp);          Q += acc + dt/6 k4
    TRY(hg::fused_rhs(ctx, d.rk_tmp.p, d.rk_k.p, false, 0.0));
    TRY(hg::fused_axpy(ctx, nullptr, nullptr, d.rk_k.p, 0.0, d.rk_acc.p, d.rk_acc.p, dt / 6.0));
    TRY(hg::fused_axpy(ctx, d.Q.p, d.Q.p, d.rk_acc.p, 1.0, nullptr, nullptr, 0.0));
  }
  return HG_OK;
}

// OrdinaryDiffEq's `Euler()` (swe_2D_forward_simulation.jl:41, swe_2D_inversion.jl:272-273): u+ = u + dt f(u), WITHOUT the
// dry mask of the customized stepper (that one is hg_step_euler).
int hg_step_ode_euler(hg_ctx* ctx, double dt, int64_t nsteps) {
  if (!ctx || nsteps < 0) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_step_ode_euler: no resident state"; return HG_ERR_STATE; }
  if (ctx->opt.path == 1) { ctx->err = "hg_step_ode_euler needs the fused path (path=0)"; return HG_ERR_ARG; }
  TRY(need_exchange_guard(ctx, "hg_step_ode_euler", nsteps > 1));
  ctx->ab3_step = 1;   // another stepper moves the state: hg_step_ab3 starts again
  ctx->state_gen++;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  for (int64_t s = 0; s < nsteps; ++s) {
    TRY(hg::fused_rhs(ctx, d.Q.p, d.dQ.p, false, 0.0));
    TRY(hg::fused_axpy(ctx, d.Q.p, d.Q.p, d.dQ.p, dt, nullptr, nullptr, 0.0));
  }
  return HG_OK;
}

// OrdinaryDiffEq's `AB3()` (swe_2D_forward_simulation.jl:46-47, swe_2D_inversion.jl:274-275): three-step Adams-Bashforth
//   u+ = u + dt/12 (23 f_n - 16 f_{n-1} + 5 f_{n-2}),
// the first two steps by Ralston's second-order method  u+ = u + dt/4 (k1 + 3 f(u + 2/3 dt k1))  (AB3ConstantCache's
// perform_step!).  One RHS launch per step once started; the two older slopes stay resident between calls (ts_k[1], ts_k[2]).
int hg_step_ab3(hg_ctx* ctx, double dt, int64_t nsteps, int32_t restart) {
  if (!ctx || nsteps < 0) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_step_ab3: no resident state"; return HG_ERR_STATE; }
  if (ctx->opt.path == 1) { ctx->err = "hg_step_ab3 needs the fused path (path=0)"; return HG_ERR_ARG; }
  TRY(need_exchange_guard(ctx, "hg_step_ab3", true));
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const size_t n3 = 3 * (size_t)ctx->fh.Ns;
  for (int m = 0; m < 4; ++m)
    if (d.ts_k[m].n != n3) { TRY(al(ctx, d.ts_k[m], n3)); restart = 1; }
  if (d.rk_tmp.n != n3) TRY(al(ctx, d.rk_tmp, n3));
  if (restart) ctx->ab3_step = 1;
  ctx->state_gen++;
  for (int64_t s = 0; s < nsteps; ++s) {
    double* k1 = d.ts_k[0].p;
    TRY(hg::fused_rhs(ctx, d.Q.p, k1, false, 0.0));
    if (ctx->ab3_step <= 2) {
      const double* ks1[1] = {k1};
      const double c1[1] = {2.0 / 3.0 * dt};
      TRY(hg::fused_lincomb(ctx, d.rk_tmp.p, d.Q.p, 1, ks1, c1));
      TRY(hg::fused_rhs(ctx, d.rk_tmp.p, d.ts_k[3].p, false, 0.0));
      const double* ks2[2] = {k1, d.ts_k[3].p};
      const double c2[2] = {dt / 4.0, 3.0 * dt / 4.0};
      TRY(hg::fused_lincomb(ctx, d.Q.p, d.Q.p, 2, ks2, c2));
      std::swap(d.ts_k[0].p, ctx->ab3_step == 1 ? d.ts_k[2].p : d.ts_k[1].p);   // f_0 -> k3, f_1 -> k2
      ctx->ab3_step++;
    } else {
      const double* ks3[3] = {k1, d.ts_k[1].p, d.ts_k[2].p};
      const double c3[3] = {23.0 * dt / 12.0, -16.0 * dt / 12.0, 5.0 * dt / 12.0};
      TRY(hg::fused_lincomb(ctx, d.Q.p, d.Q.p, 3, ks3, c3));
      std::swap(d.ts_k[1].p, d.ts_k[2].p);   // k3 <- k2 (the old k3 buffer becomes free)
      std::swap(d.ts_k[0].p, d.ts_k[1].p);   // k2 <- k1
    }
  }
  return HG_OK;
}

// DiffEqBase.fastpow, the power OrdinaryDiffEq's PI controller used in the generation the reference ran (DifferentialEquations
// 7.15): Float32 arithmetic, log2 from a rational approximation on the significand reduced to [0.75, 1.5) (Goldberg, "Fast
// approximate logarithms", (s - 1)(a (s - 1) + b) / ((s - 1) + c)), then exp2.  ~1e-5 relative error -- enough to change
// every step size in the fifth digit, which is why reproducing the reference's saved trajectories needs it.
static float fastlog2f(float x) {
  const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
  uint32_t u;
  std::memcpy(&u, &x, 4);
  const int32_t ex = (int32_t)((u & 0x7F800000u) >> 23);
  uint32_t u2;
  float fexp;
  if (u & 0x00400000u) {   // significand >= 1.5: halve it (exponent field 126) and compensate in the exponent
    u2 = (u & 0x007FFFFFu) | 0x3F000000u;
    fexp = (float)(ex - 126);
  } else {
    u2 = (u & 0x007FFFFFu) | 0x3F800000u;
    fexp = (float)(ex - 127);
  }
  float sig;
  std::memcpy(&sig, &u2, 4);
  volatile float s = sig - 1.0f;                 // volatile: every operation rounded to Float32, no contraction
  volatile float t1 = a * s;
  volatile float t2 = t1 + b;
  volatile float t3 = s * t2;
  volatile float t4 = s + c;
  volatile float t5 = t3 / t4;
  return fexp + t5;
}
double hg_fastpow(double x, double y) {
  if (x == 0.0) return 0.0;
  volatile float e = (float)y * fastlog2f((float)x);
  return (double)exp2f(e);
}
int hg_set_controller_pow(hg_ctx* ctx, int32_t mode) {
  if (!ctx || mode < 0 || mode > 1) return HG_ERR_ARG;
  ctx->controller_fastpow = mode == 1;
  return HG_OK;
}

// Tsit5 (Tsitouras 2011) with OrdinaryDiffEq's PI step-size controller -- what the reference's forward and sensitivity
// drivers run by default: solve(prob, Tsit5(), adaptive=..., dt=dt, saveat=t_save; abstol=1e-6, reltol=1e-3)
// (swe_2D_forward_simulation.jl:38-41, swe_2D_sensitivity.jl:38-43).  Device-resident: seven fused RHS launches per step
// (six with FSAL), stage states by k_lincomb, the scaled error norm by a fixed-shape reduction; one 8-byte D2H per step
// for the accept / reject decision.  Save times: dense = 0 treats them as tstops (the step is clipped to land on them, the
// controller's own proposal is kept); dense = 1 is OrdinaryDiffEq's saveat: the steps ignore the save times and every
// accepted step evaluates Tsit5's fourth-order dense output u + h sum_i b_i(theta) k_i at the save times it has passed
// (one k_lincomb launch per save), copying the new state when a save time is the step end.
namespace {
int solve_tsit5(hg_ctx* ctx, double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol, const double* t_save,
                int64_t n_save, double* Q_save, int64_t* stats, bool dense) {
  if (!ctx || !(t1 > t0) || !(dt > 0.0) || n_save < 0 || (n_save > 0 && (!t_save || !Q_save))) return HG_ERR_ARG;
  if (adaptive && (!(abstol > 0.0) || !(reltol > 0.0))) { ctx->err = "hg_solve_tsit5: tolerances must be positive"; return HG_ERR_ARG; }
  if (!ctx->state_set) { ctx->err = "hg_solve_tsit5: no resident state"; return HG_ERR_STATE; }
  ctx->state_gen++;
  if (ctx->opt.path == 1) { ctx->err = "hg_solve_tsit5 needs the fused path (path=0)"; return HG_ERR_ARG; }
  // multi-rank contexts: the stages exchange their halos through the library transport; the adaptive controller needs the
  // error norm of the WHOLE mesh, i.e. a sum over ranks -- done on the host by the caller's all-reduce (hg_comm_set_allreduce)
  if (ctx->n_halo > 0 && !hg_comm_ready(ctx)) { ctx->err = "hg_solve_tsit5: multi-rank context without a library-owned halo exchange (hg_comm_connect)"; return HG_ERR_ARG; }
  if (ctx->n_halo > 0 && adaptive && !ctx->allreduce) {
    ctx->err = "hg_solve_tsit5: an adaptive solve on a multi-rank context needs hg_comm_set_allreduce (the error norm is a sum over ranks)";
    return HG_ERR_ARG;
  }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  static const double A[7][6] = {
      {0, 0, 0, 0, 0, 0},
      {0.161, 0, 0, 0, 0, 0},
      {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
      {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
      {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
      {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
      {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
  static const double BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                               0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
  const double beta2 = 2.0 / 25.0, beta1 = 7.0 / 50.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1e-4;
  hg::FusedDev& d = ctx->fd;
  const size_t n3 = 3 * (size_t)ctx->fh.Ns;
  for (int m = 0; m < 7; ++m)
    if (d.ts_k[m].n != n3) { TRY(al(ctx, d.ts_k[m], n3)); CK(ctx, cudaMemsetAsync(d.ts_k[m].p, 0, n3 * 8, ctx->stream)); }
  if (d.ts_new.n != n3) { TRY(al(ctx, d.ts_new, n3)); CK(ctx, cudaMemsetAsync(d.ts_new.p, 0, n3 * 8, ctx->stream)); }
  if (d.rk_tmp.n != n3) TRY(al(ctx, d.rk_tmp, n3));
  const int nblk = hg::fused_err_blocks(ctx);
  if (d.ts_part.n != (size_t)nblk) TRY(al(ctx, d.ts_part, nblk));
  if (d.ts_sum.n != 1) TRY(al(ctx, d.ts_sum, 1));
  // the stops: save times inside (t0, t1] in ascending order, then t1
  std::vector<std::pair<double, int64_t>> stops;
  for (int64_t i = 0; i < n_save; ++i) {
    if (t_save[i] == t0) TRY(download3(ctx, d.Q.p, Q_save + (size_t)i * 3 * ctx->N));
    else if (t_save[i] > t0 && t_save[i] <= t1) stops.push_back({t_save[i], i});
    else { ctx->err = "hg_solve_tsit5: save time outside [t0, t1]"; return HG_ERR_ARG; }
  }
  std::sort(stops.begin(), stops.end());
  std::vector<std::pair<double, int64_t>> pending;        // dense output: save times still ahead, ascending
  if (dense) { pending.swap(stops); }
  size_t next_save = 0;
  if (stops.empty() || stops.back().first < t1) stops.push_back({t1, -1});
  // Tsit5Interp: b_i(theta) = theta (r_i1 + theta (r_i2 + theta (r_i3 + theta r_i4)))
  static const double RI[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                  {0.0, 0.13169999999999998, -0.2234, 0.1017},
                                  {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                                  {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                                  {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                                  {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                                  {0.0, 1.5, -4.0, 2.5}};
  ctx->ab3_step = 1;   // the stage buffers double as hg_step_ab3's history
  ctx->last_steps.clear();
  int64_t n_acc = 0, n_rej = 0, n_rhs = 0;
  double t = t0, dt_ctrl = dt, qold = qoldinit;
  double* k[7];
  for (int m = 0; m < 7; ++m) k[m] = d.ts_k[m].p;
  TRY(hg::fused_rhs(ctx, d.Q.p, k[0], false, 0.0));
  ++n_rhs;
  for (size_t si = 0; si < stops.size(); ++si) {
    const double ts = stops[si].first;
    while (t < ts) {
      double h = std::min(dt_ctrl, ts - t);
      const bool clipped = h < dt_ctrl;
      if (ts - (t + h) < 1e-12 * std::max(1.0, std::fabs(ts))) h = ts - t;   // do not leave a sliver before the stop
      double coef[7];
      for (int i = 1; i < 7; ++i) {
        for (int j = 0; j < i; ++j) coef[j] = h * A[i][j];
        double* y = i < 6 ? d.rk_tmp.p : d.ts_new.p;
        TRY(hg::fused_lincomb(ctx, y, d.Q.p, i, k, coef));
        TRY(hg::fused_rhs(ctx, y, k[i], false, 0.0));
        ++n_rhs;
      }
      bool accept = true;
      double q = 1.0, q11 = 0.0, eest = 0.0;
      if (adaptive) {
        for (int m = 0; m < 7; ++m) coef[m] = h * BT[m];
        TRY(hg::fused_err_norm(ctx, d.Q.p, d.ts_new.p, 7, k, coef, abstol, reltol, d.ts_part.p, d.ts_sum.p));
        double sum = 0.0;
        CK(ctx, cudaMemcpyAsync(&sum, d.ts_sum.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        double tot[2] = {sum, 3.0 * (double)ctx->N};
        if (ctx->n_halo > 0 && ctx->allreduce(tot, 2, ctx->allreduce_user) != 0) { ctx->err = "hg_solve_tsit5: the caller's all-reduce failed"; return HG_ERR_STATE; }
        eest = std::sqrt(tot[0] / tot[1]);
        if (!(eest == eest)) { ctx->err = "hg_solve_tsit5: the error estimate is NaN"; return HG_ERR_STATE; }
        if (eest == 0.0) { q11 = 0.0; q = 1.0 / qmax; }
        else {
          q11 = ctx->controller_fastpow ? hg_fastpow(eest, beta1) : std::pow(eest, beta1);
          q = q11 / (ctx->controller_fastpow ? hg_fastpow(qold, beta2) : std::pow(qold, beta2));
          q = std::max(1.0 / qmax, std::min(1.0 / qmin, q / gamma));
        }
        accept = eest <= 1.0;
      }
      if (accept) {
        const double tnew = (h == ts - t) ? ts : t + h;
        for (; next_save < pending.size() && pending[next_save].first <= tnew; ++next_save) {
          double* out = Q_save + (size_t)pending[next_save].second * 3 * ctx->N;
          if (pending[next_save].first == tnew) { TRY(download3(ctx, d.ts_new.p, out)); continue; }
          const double th = (pending[next_save].first - t) / h;
          for (int m = 0; m < 7; ++m) coef[m] = h * (th * (RI[m][0] + th * (RI[m][1] + th * (RI[m][2] + th * RI[m][3]))));
          TRY(hg::fused_lincomb(ctx, d.rk_tmp.p, d.Q.p, 7, k, coef));   // rk_tmp is free between steps
          TRY(download3(ctx, d.rk_tmp.p, out));
        }
        std::swap(d.Q.p, d.ts_new.p);            // the candidate becomes the state (equal sizes, like hg_step_euler's swap)
        std::swap(d.ts_k[0].p, d.ts_k[6].p);     // FSAL: k7 of this step is k1 of the next
        for (int m = 0; m < 7; ++m) k[m] = d.ts_k[m].p;
        t = tnew;
        ++n_acc;
        ctx->last_steps.push_back(h);
        if (adaptive) {
          qold = std::max(eest, qoldinit);
          const double prop = h / q;
          dt_ctrl = clipped ? std::max(prop, dt_ctrl) : prop;   // a step clipped by a stop does not shrink the proposal
        }
      } else {
        ++n_rej;
        dt_ctrl = h / std::min(1.0 / qmin, q11 / gamma);
        if (!(dt_ctrl > 1e-14 * std::max(1.0, std::fabs(t)))) { ctx->err = "hg_solve_tsit5: step size underflow"; return HG_ERR_STATE; }
      }
    }
    t = ts;
    if (stops[si].second >= 0) TRY(download3(ctx, d.Q.p, Q_save + (size_t)stops[si].second * 3 * ctx->N));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) { stats[0] = n_acc; stats[1] = n_rej; stats[2] = n_rhs; }
  return check_err_flag(ctx);
}
}  // namespace

int hg_solve_tsit5(hg_ctx* ctx, double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol, const double* t_save,
                   int64_t n_save, double* Q_save, int64_t* stats) {
  return solve_tsit5(ctx, t0, t1, dt, adaptive, abstol, reltol, t_save, n_save, Q_save, stats, false);
}
int hg_solve_tsit5_dense(hg_ctx* ctx, double t0, double t1, double dt, int32_t adaptive, double abstol, double reltol,
                         const double* t_save, int64_t n_save, double* Q_save, int64_t* stats) {
  return solve_tsit5(ctx, t0, t1, dt, adaptive, abstol, reltol, t_save, n_save, Q_save, stats, true);
}

int hg_euler_adjoint(hg_ctx* ctx, const double* Q0, const double* params, int64_t np, int32_t active, double dt, int64_t nsteps,
                     const double* lambda_T, double* Q_T, double* Q0bar, double* pbar) {
  if (!ctx || !Q0 || !lambda_T || !Q0bar || nsteps < 1 || !(dt > 0.0)) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_euler_adjoint needs the fused path"; return HG_ERR_ARG; }
  TRY(no_closure(ctx, "hg_euler_adjoint"));
  // multi-rank contexts: every RHS / VJP of the sweeps exchanges state (and cotangent) halos through the library transport;
  // Q0bar covers the owned cells, pbar is this rank's partial sum (the caller adds the ranks' pbar in rank order)
  if (ctx->n_halo > 0 && !hg_comm_ready(ctx)) { ctx->err = "hg_euler_adjoint: multi-rank context without a library-owned halo exchange (hg_comm_connect)"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  if (ctx->active != HG_PARAM_NONE && !pbar) { ctx->err = "hg_euler_adjoint: pbar is NULL"; return HG_ERR_ARG; }
  hg::FusedDev& d = ctx->fd;
  const size_t n3 = 3 * (size_t)ctx->fh.Ns;
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  // checkpoint spacing ~ sqrt(nsteps): (nsteps/C + C) resident states
  const int64_t C = std::max<int64_t>(1, (int64_t)std::ceil(std::sqrt((double)nsteps)));
  const int64_t nck = (nsteps + C - 1) / C;
  hg::DBuf<double> ck, seg, lam, lam_tmp, pacc;
  CK(ctx, ck.alloc((size_t)nck * n3)); CK(ctx, seg.alloc((size_t)(C + 1) * n3));
  CK(ctx, lam.alloc(n3)); CK(ctx, lam_tmp.alloc(n3)); CK(ctx, pacc.alloc((size_t)std::max<int64_t>(npar, 1)));
  CK(ctx, cudaMemsetAsync(pacc.p, 0, pacc.bytes(), ctx->stream));
  CK(ctx, cudaMemsetAsync(lam.p, 0, lam.bytes(), ctx->stream));
  // ---- forward sweep, storing the state at the start of every segment
  TRY(hg_set_state(ctx, Q0));
  for (int64_t s = 0; s < nsteps; ++s) {
    if (s % C == 0) CK(ctx, cudaMemcpyAsync(ck.p + (s / C) * n3, d.Q.p, n3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(hg::fused_rhs(ctx, d.Q.p, d.Q2.p, true, dt));
    std::swap(d.Q.p, d.Q2.p);
  }
  if (Q_T) TRY(download3(ctx, d.Q.p, Q_T));
  // ---- terminal cotangent (reference order -> internal)
  CK(ctx, cudaMemcpyAsync(d.stage.p, lambda_T, 3 * ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream));
  TRY(hg::fused_permute(ctx, true, d.stage.p, lam.p));
  // ---- reverse sweep, one segment at a time
  for (int64_t k = nck - 1; k >= 0; --k) {
    const int64_t s0 = k * C, s1 = std::min<int64_t>(nsteps, s0 + C);
    CK(ctx, cudaMemcpyAsync(seg.p, ck.p + k * n3, n3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int64_t s = s0; s < s1; ++s) TRY(hg::fused_rhs(ctx, seg.p + (s - s0) * n3, seg.p + (s - s0 + 1) * n3, true, dt));
    for (int64_t s = s1 - 1; s >= s0; --s)
      TRY(hg::fused_adjoint_step(ctx, seg.p + (s - s0) * n3, seg.p + (s - s0 + 1) * n3, lam.p, lam_tmp.p, pacc.p, npar, dt));
  }
  TRY(download3(ctx, lam.p, Q0bar));
  if (npar > 0) CK(ctx, cudaMemcpyAsync(pbar, pacc.p, npar * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

// ---- discrete adjoint of nsteps of a fixed-step explicit Runge-Kutta method (method 0: classical RK4, 1: Tsit5) --
// "discretise, then differentiate": exactly the derivative of what hg_step_rk4 / hg_solve_tsit5(adaptive = 0) compute, the
// counterpart for the SciML solvers of what hg_euler_adjoint is for the customized Euler loop (swe_2D_inversion.jl:304-339).
//   forward   Y_i = u_n + h sum_{j<i} a_ij k_j,  k_i = f(Y_i),  u_{n+1} = u_n + h sum_i b_i k_i
//   reverse   kbar_i = h b_i lam + h sum_{j>i} a_ji Ybar_j,  Ybar_i = J_u(Y_i)^T kbar_i,  pbar += J_p(Y_i)^T kbar_i  (i = s..1),
//             lam <- lam + sum_i Ybar_i
// The stage states of a step are recomputed from its start state; start states are kept per segment (~sqrt(nsteps)
// checkpoints), like hg_euler_adjoint.
namespace {
struct RkTable {
  int s;
  double a[7][6];
  double b[7];
};
const RkTable kRK4 = {4, {{0}, {0.5}, {0, 0.5}, {0, 0, 1.0}}, {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0}};
const RkTable kTsit5 = {6,
                        {{0},
                         {0.161},
                         {-0.008480655492356989, 0.335480655492357},
                         {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
                         {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
                         {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383}},
                        {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};

// one step from un: stage states into Y[1..s-1] (Y_0 is un itself), slopes into K[0..s-1], result into unext
int rk_forward_step(hg_ctx* ctx, const RkTable& tb, double h, const double* un, double* const* Y, double* const* K, double* unext) {
  double coef[7];
  for (int i = 0; i < tb.s; ++i) {
    const double* y = un;
    if (i > 0) {
      for (int j = 0; j < i; ++j) coef[j] = h * tb.a[i][j];
      TRY(hg::fused_lincomb(ctx, Y[i], un, i, K, coef));
      y = Y[i];
    }
    TRY(hg::fused_rhs(ctx, y, K[i], false, 0.0));
  }
  for (int i = 0; i < tb.s; ++i) coef[i] = h * tb.b[i];
  return hg::fused_lincomb(ctx, unext, un, tb.s, K, coef);
}
}  // namespace

// hs[nsteps]: the step sizes (all equal for hg_rk_adjoint; the accepted steps of an adaptive solve for hg_rk_adjoint_steps)
static int rk_adjoint_impl(hg_ctx* ctx, int32_t method, const double* Q0, const double* params, int64_t np, int32_t active,
                           const double* hs, int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar) {
  if (ctx->opt.path == 1) { ctx->err = "hg_rk_adjoint needs the fused path"; return HG_ERR_ARG; }
  TRY(no_closure(ctx, "hg_rk_adjoint"));
  if (ctx->n_halo > 0 && !hg_comm_ready(ctx)) { ctx->err = "hg_rk_adjoint: multi-rank context without a library-owned halo exchange (hg_comm_connect)"; return HG_ERR_ARG; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  TRY(bind_params(ctx, params, np, active));
  if (ctx->active != HG_PARAM_NONE && !pbar) { ctx->err = "hg_rk_adjoint: pbar is NULL"; return HG_ERR_ARG; }
  const RkTable& tb = method == 0 ? kRK4 : kTsit5;
  hg::FusedDev& d = ctx->fd;
  const size_t n3 = 3 * (size_t)ctx->fh.Ns;
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  const int64_t C = std::max<int64_t>(1, (int64_t)std::ceil(std::sqrt((double)nsteps)));
  const int64_t nck = (nsteps + C - 1) / C;
  hg::DBuf<double> ck, seg, lam, zero, pacc, Ybuf[7], Kbuf[7];
  CK(ctx, ck.alloc((size_t)nck * n3)); CK(ctx, seg.alloc((size_t)(C + 1) * n3));
  CK(ctx, lam.alloc(n3)); CK(ctx, zero.alloc(n3)); CK(ctx, pacc.alloc((size_t)std::max<int64_t>(npar, 1)));
  double *Y[7] = {}, *K[7] = {};
  for (int i = 0; i < tb.s; ++i) {
    CK(ctx, Ybuf[i].alloc(n3)); CK(ctx, Kbuf[i].alloc(n3));
    CK(ctx, cudaMemsetAsync(Ybuf[i].p, 0, n3 * 8, ctx->stream)); CK(ctx, cudaMemsetAsync(Kbuf[i].p, 0, n3 * 8, ctx->stream));
    Y[i] = Ybuf[i].p; K[i] = Kbuf[i].p;
  }
  CK(ctx, cudaMemsetAsync(pacc.p, 0, pacc.bytes(), ctx->stream));
  CK(ctx, cudaMemsetAsync(lam.p, 0, lam.bytes(), ctx->stream));
  CK(ctx, cudaMemsetAsync(zero.p, 0, zero.bytes(), ctx->stream));
  CK(ctx, cudaMemsetAsync(seg.p, 0, seg.bytes(), ctx->stream));
  // ---- forward sweep, storing the state at the start of every segment
  TRY(hg_set_state(ctx, Q0));
  for (int64_t s = 0; s < nsteps; ++s) {
    if (s % C == 0) CK(ctx, cudaMemcpyAsync(ck.p + (s / C) * n3, d.Q.p, n3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(rk_forward_step(ctx, tb, hs[s], d.Q.p, Y, K, d.Q2.p));
    std::swap(d.Q.p, d.Q2.p);
  }
  if (Q_T) TRY(download3(ctx, d.Q.p, Q_T));
  // ---- terminal cotangent (reference order -> internal)
  CK(ctx, cudaMemcpyAsync(d.stage.p, lambda_T, 3 * ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream));
  TRY(hg::fused_permute(ctx, true, d.stage.p, lam.p));
  const int cfg = hg::fused_cfg_id(ctx);
  // ---- reverse sweep, one segment at a time; K[] is reused for the stage cotangents kbar_i once a step's Y_i are known
  for (int64_t k = nck - 1; k >= 0; --k) {
    const int64_t s0 = k * C, s1 = std::min<int64_t>(nsteps, s0 + C);
    CK(ctx, cudaMemcpyAsync(seg.p, ck.p + k * n3, n3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int64_t s = s0; s < s1 - 1; ++s) TRY(rk_forward_step(ctx, tb, hs[s], seg.p + (s - s0) * n3, Y, K, seg.p + (s - s0 + 1) * n3));
    for (int64_t s = s1 - 1; s >= s0; --s) {
      const double* un = seg.p + (s - s0) * n3;
      const double dt = hs[s];
      TRY(rk_forward_step(ctx, tb, dt, un, Y, K, d.Q2.p));                 // the step's stage states Y_1 .. Y_{s-1}
      for (int i = 0; i < tb.s; ++i) {                                     // kbar_i = h b_i lam
        const double* one[1] = {lam.p};
        const double cb[1] = {dt * tb.b[i]};
        TRY(hg::fused_lincomb(ctx, K[i], zero.p, 1, one, cb));
      }
      for (int i = tb.s - 1; i >= 0; --i) {
        TRY(hg::fused_vjp(ctx, cfg, i == 0 ? un : Y[i], K[i], d.Qbar.p));  // Ybar_i and the parameter adjoint of this stage
        TRY(hg::fused_acc_pbar(ctx, npar, pacc.p, 1.0));
        TRY(hg::fused_axpy(ctx, lam.p, lam.p, d.Qbar.p, 1.0, nullptr, nullptr, 0.0));
        for (int j = 0; j < i; ++j)
          if (tb.a[i][j] != 0.0) TRY(hg::fused_axpy(ctx, K[j], K[j], d.Qbar.p, dt * tb.a[i][j], nullptr, nullptr, 0.0));
      }
    }
  }
  TRY(download3(ctx, lam.p, Q0bar));
  if (npar > 0) CK(ctx, cudaMemcpyAsync(pbar, pacc.p, npar * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return check_err_flag(ctx);
}

int hg_rk_adjoint(hg_ctx* ctx, int32_t method, const double* Q0, const double* params, int64_t np, int32_t active, double dt,
                  int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar) {
  if (!ctx || !Q0 || !lambda_T || !Q0bar || nsteps < 1 || !(dt > 0.0) || method < 0 || method > 1) return HG_ERR_ARG;
  const std::vector<double> hs((size_t)nsteps, dt);
  return rk_adjoint_impl(ctx, method, Q0, params, np, active, hs.data(), nsteps, lambda_T, Q_T, Q0bar, pbar);
}

int hg_rk_adjoint_steps(hg_ctx* ctx, int32_t method, const double* Q0, const double* params, int64_t np, int32_t active,
                        const double* h, int64_t nsteps, const double* lambda_T, double* Q_T, double* Q0bar, double* pbar) {
  if (!ctx || !Q0 || !lambda_T || !Q0bar || !h || nsteps < 1 || method < 0 || method > 1) return HG_ERR_ARG;
  for (int64_t s = 0; s < nsteps; ++s)
    if (!(h[s] > 0.0)) { ctx->err = "hg_rk_adjoint_steps: step sizes must be positive"; return HG_ERR_ARG; }
  return rk_adjoint_impl(ctx, method, Q0, params, np, active, h, nsteps, lambda_T, Q_T, Q0bar, pbar);
}

int hg_last_steps(const hg_ctx* ctx, double* h, int64_t capacity, int64_t* n) {
  if (!ctx || !n) return HG_ERR_ARG;
  *n = (int64_t)ctx->last_steps.size();
  if (h)
    for (int64_t s = 0; s < std::min<int64_t>(capacity, *n); ++s) h[s] = ctx->last_steps[(size_t)s];
  return HG_OK;
}

int hg_custom_ode_solve(hg_ctx* ctx, const double* Q0, const double* params, int64_t np, int32_t active, double t0,
                        double t1, double dt, double* sol, int64_t cap, int64_t* n_saves) {
  if (!ctx || !Q0 || !sol || !n_saves || !(dt > 0.0)) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  // length(t_start:dt:t_end), custom_ODE_solvers.jl:40
  const int64_t nsteps = t1 < t0 ? 0 : (int64_t)std::floor((t1 - t0) / dt + 1e-9) + 1;
  *n_saves = nsteps;
  if (cap < nsteps) { ctx->err = "hg_custom_ode_solve: sol has room for " + std::to_string(cap) + " of " + std::to_string(nsteps) + " saves"; return HG_ERR_ARG; }
  TRY(need_exchange_guard(ctx, "hg_custom_ode_solve", nsteps > 1));
  TRY(bind_params(ctx, params, np, active));
  TRY(hg_set_state(ctx, Q0));
  for (int64_t s = 0; s < nsteps; ++s) {
    TRY(hg_step_euler(ctx, dt, 1));
    TRY(hg_get_state(ctx, sol + s * 3 * ctx->N));  // save_freq = 1: every step is saved (:80-82)
  }
  return check_err_flag(ctx);
}

int hg_time_rhs(hg_ctx* ctx, int32_t n, int32_t fused_euler, double dt, float* ms) {
  if (!ctx || !ms || n <= 0) return HG_ERR_ARG;
  if (!ctx->state_set) { ctx->err = "hg_time_rhs: no resident state"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = 0; i < n; ++i) {
    if (fused_euler) TRY(hg_step_euler(ctx, dt, 1));
    else TRY(hg_rhs_resident(ctx));
  }
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return check_err_flag(ctx);
}

// device time of n launches of the fused forward-mode kernel with K directions each (tangents = the resident cotangent
// buffer's content replicated; Manning zones as directions when they are the active parameter)
int hg_time_jvp(hg_ctx* ctx, int32_t K, int32_t n, float* ms) {
  if (!ctx || !ms || n <= 0 || K <= 0) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_time_jvp needs the fused path"; return HG_ERR_ARG; }
  if (!ctx->state_set || !ctx->lam_set) { ctx->err = "hg_time_jvp: state or lambda not set"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  const size_t Ns3 = 3 * (size_t)ctx->fh.Ns;
  if (d.j_V.n < (size_t)K * Ns3) { TRY(al(ctx, d.j_V, (size_t)K * Ns3)); TRY(al(ctx, d.j_out, (size_t)K * Ns3)); }
  for (int32_t k = 0; k < K; ++k) CK(ctx, cudaMemcpyAsync(d.j_V.p + k * Ns3, d.lam.p, Ns3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  const int64_t npar = ctx->active == HG_PARAM_NONE ? 0 : ctx->n_params;
  const double* d_pdot = nullptr;
  if (npar > 0) {
    std::vector<double> pd((size_t)(K * npar), 0.0);
    for (int64_t k = 0; k < K; ++k) pd[(size_t)(k * npar + (k % npar))] = 1.0;
    if (d.j_pdot.n < pd.size()) TRY(al(ctx, d.j_pdot, pd.size()));
    CK(ctx, cudaMemcpyAsync(d.j_pdot.p, pd.data(), pd.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    d_pdot = d.j_pdot.p;
  }
  const int cfg = hg::fused_cfg_id(ctx);
  TRY(hg::fused_jvp(ctx, cfg, d.Q.p, d.j_V.p, d_pdot, d.dQ.p, d.j_out.p, K));   // warm
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = 0; i < n; ++i) TRY(hg::fused_jvp(ctx, cfg, d.Q.p, d.j_V.p, d_pdot, d.dQ.p, d.j_out.p, K));
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return check_err_flag(ctx);
}

int hg_time_vjp(hg_ctx* ctx, int32_t n, float* ms) {
  if (!ctx || !ms || n <= 0) return HG_ERR_ARG;
  if (ctx->opt.path == 1) { ctx->err = "hg_time_vjp needs the fused path"; return HG_ERR_ARG; }
  TRY(no_closure(ctx, "hg_time_vjp"));
  if (!ctx->state_set) { ctx->err = "hg_time_vjp: no resident state"; return HG_ERR_STATE; }
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::FusedDev& d = ctx->fd;
  if (!ctx->lam_set) {  // default cotangent: the resident dQdt-shaped buffer filled with ones
    std::vector<double> ones(3 * ctx->N, 1.0);
    TRY(set_lambda(ctx, ones.data()));
  }
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = 0; i < n; ++i) TRY(hg::fused_vjp(ctx, hg::fused_cfg_id(ctx), d.Q.p, d.lam.p, d.Qbar.p));
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return check_err_flag(ctx);
}

int hg_mesh_stats(const hg_ctx* ctx, int64_t* nc, int64_t* nf, int64_t* snf, int64_t* nt, int64_t* bytes) {
  if (!ctx) return HG_ERR_ARG;
  if (nc) *nc = ctx->N;
  if (nf) *nf = ctx->F;
  if (snf) *snf = ctx->sumnf;
  if (nt) *nt = ctx->fh.n_tiles;
  if (bytes) *bytes = ctx->device_bytes;
  return HG_OK;
}

int hg_plan_stats(const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f, const hg_options* opt,
                  int64_t* stats, int64_t* perm_out) {
  if (!stats) return HG_ERR_ARG;
  std::unique_ptr<hg_ctx> ctx(new hg_ctx());
  if (opt) ctx->opt = *opt; else hg_default_options(&ctx->opt);
  std::vector<int32_t> cf_ptr, cf_nb, cf_face;
  std::vector<double> cf_nx, cf_ny, cf_len;
  int rc = hg::build_host(ctx.get(), m, b, f, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc == HG_OK) rc = hg::build_tiles(ctx.get(), m, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc != HG_OK) { set_global_err(ctx->err); return rc; }
  const hg::FusedHost& fh = ctx->fh;
  int64_t nint = 0, nhalo = 0, nfaces = 0;
  for (int32_t t = 0; t < fh.n_tiles; ++t) {
    const int32_t* d = &fh.tile_desc[(size_t)t * hg::kTileDesc];
    nint += d[9]; nhalo += d[3]; nfaces += d[5];
  }
  stats[0] = fh.n_tiles; stats[1] = fh.max_local; stats[2] = fh.max_faces; stats[3] = hg::fused_smem_bytes(ctx.get());
  stats[4] = nhalo; stats[5] = nfaces; stats[6] = nint; stats[7] = ctx->sumnf;
  if (perm_out) for (int64_t i = 0; i < ctx->N; ++i) perm_out[i] = fh.perm[i];
  return HG_OK;
}

// Host-only: the stage tables of the host-buffer pipeline for this mesh (no GPU touched).  header = {n_chunks, rows per chunk,
// n_tiles, tile_cells, alignment margin in rows}; tile_stage[n_tiles] (may be NULL) = the stage a tile runs in, i.e. the chunk
// after which all its cells and halo cells have landed; chunk_done[n_chunks] (may be NULL) = the stage after which result
// chunk c may leave.  The CPU tests check both against the mesh adjacency.
int hg_plan_pipeline(const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f, const hg_options* opt, int64_t* header,
                     int32_t* tile_stage, int32_t* chunk_done) {
  if (!header) return HG_ERR_ARG;
  std::unique_ptr<hg_ctx> ctx(new hg_ctx());
  if (opt) ctx->opt = *opt; else hg_default_options(&ctx->opt);
  std::vector<int32_t> cf_ptr, cf_nb, cf_face;
  std::vector<double> cf_nx, cf_ny, cf_len;
  int rc = hg::build_host(ctx.get(), m, b, f, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc == HG_OK) rc = hg::build_tiles(ctx.get(), m, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc != HG_OK) { set_global_err(ctx->err); return rc; }
  const hg::FusedHost& fh = ctx->fh;
  header[0] = fh.n_chunks; header[1] = fh.chunk_cells; header[2] = fh.n_tiles; header[3] = fh.T; header[4] = hg::kPipeAlign;
  if (tile_stage)
    for (int32_t s = 0; s < fh.n_chunks; ++s)
      for (int32_t k = fh.stage_ptr[s]; k < fh.stage_ptr[s + 1]; ++k) tile_stage[fh.tile_order[k]] = s;
  if (chunk_done) for (int32_t c = 0; c < fh.n_chunks; ++c) chunk_done[c] = fh.chunk_done[c];
  return HG_OK;
}

// Host-only: the tile tables hg_create would upload for this mesh, by name (no GPU touched).  The CPU tests walk them with a
// plain-Python model of the tile kernel's data path (tests/test_tile_tables_cpu.py): what the builder emits is checked against
// the mesh and the reference restatement without a device.
struct hg_plan { std::unique_ptr<hg_ctx> ctx; std::vector<int64_t> dims; };
int hg_plan_open(const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f, const hg_options* opt, hg_plan** out) {
  if (!out) return HG_ERR_ARG;
  std::unique_ptr<hg_plan> p(new hg_plan());
  p->ctx.reset(new hg_ctx());
  hg_ctx* ctx = p->ctx.get();
  if (opt) ctx->opt = *opt; else hg_default_options(&ctx->opt);
  std::vector<int32_t> cf_ptr, cf_nb, cf_face;
  std::vector<double> cf_nx, cf_ny, cf_len;
  int rc = hg::build_host(ctx, m, b, f, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc == HG_OK) rc = hg::build_tiles(ctx, m, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
  if (rc != HG_OK) { set_global_err(ctx->err); return rc; }
  const hg::FusedHost& fh = ctx->fh;
  p->dims = {ctx->N, ctx->B, fh.n_tiles, fh.T, fh.NF, fh.Ns, hg::kTileDesc, fh.n_chunks, fh.n_interior_tiles, fh.comm_band0};
  *out = p.release();
  return HG_OK;
}
void hg_plan_close(hg_plan* p) { delete p; }
// name: dims | perm iperm tile_desc halo bface_e tile_order band_order comm_order (i32) | face_lr (u32) | cf_idx (u16) | face_nx face_ny face_len (f64) |
// bc_type bc_group bc_ghost bc_cell_ref inlet_ptr bcell_ref bcell_ptr bcell_ent cf_rev (i32) | bc_nx bc_ny bc_l53 bc_l23 bc_hstill bc_zb (f64)
int hg_plan_array(const hg_plan* p, const char* name, const void** ptr, int64_t* count, int32_t* dtype) {
  if (!p || !name || !ptr || !count || !dtype) return HG_ERR_ARG;
  const hg::FusedHost& fh = p->ctx->fh;
  const hg::BcHost& bh = p->ctx->bch;
  const std::string n(name);
  auto give = [&](const auto& v, int32_t code) { *ptr = v.data(); *count = (int64_t)v.size(); *dtype = code; return HG_OK; };
  if (n == "dims") return give(p->dims, 1);
  if (n == "perm") return give(fh.perm, 3);
  if (n == "iperm") return give(fh.iperm, 3);
  if (n == "tile_desc") return give(fh.tile_desc, 3);
  if (n == "halo") return give(fh.halo, 3);
  if (n == "bface_e") return give(fh.bface_e, 3);
  if (n == "tile_order") return give(fh.tile_order, 3);
  if (n == "band_order") return give(fh.band_order, 3);
  if (n == "comm_order") return give(fh.comm_order, 3);
  if (n == "face_lr") return give(fh.face_lr, 4);
  if (n == "cf_idx") return give(fh.cf_idx, 5);
  if (n == "face_nx") return give(fh.face_nx, 0);
  if (n == "face_ny") return give(fh.face_ny, 0);
  if (n == "face_len") return give(fh.face_len, 0);
  if (n == "bc_type") return give(bh.type, 3);
  if (n == "bc_group") return give(bh.group, 3);
  if (n == "bc_ghost") return give(bh.ghost, 3);
  if (n == "bc_cell_ref") return give(bh.cell_ref, 3);
  if (n == "inlet_ptr") return give(bh.inlet_ptr, 3);
  if (n == "bcell_ref") return give(bh.bcell_ref, 3);
  if (n == "bcell_ptr") return give(bh.bcell_ptr, 3);
  if (n == "bcell_ent") return give(bh.bcell_ent, 3);
  if (n == "cf_rev") return give(bh.cf_rev, 3);
  if (n == "bc_nx") return give(bh.nx, 0);
  if (n == "bc_ny") return give(bh.ny, 0);
  if (n == "bc_l53") return give(bh.l53, 0);
  if (n == "bc_l23") return give(bh.l23, 0);
  if (n == "bc_hstill") return give(bh.hstill_g, 0);
  if (n == "bc_zb") return give(bh.zb_g, 0);
  return HG_ERR_ARG;
}

// Host-only: the rows [r0, r1) chunk c of a component moves when that component's row 0 sits at host address host_addr -- the
// geometry hg_rhs / hg_rhs_vjp use (chunk_rows above), exposed so that the CPU tests can check coverage, alignment and the
// margins the stage tables assume.
int hg_debug_chunk_rows(uint64_t host_addr, int32_t c, int32_t K, int64_t rows_per_chunk, int64_t N, int64_t* r0, int64_t* r1) {
  if (!r0 || !r1 || K < 1 || c < 0 || c >= K || rows_per_chunk < 1 || N < 1) return HG_ERR_ARG;
  chunk_rows(reinterpret_cast<const double*>(static_cast<uintptr_t>(host_addr)), c, K, rows_per_chunk, N, *r0, *r1);
  return HG_OK;
}

int hg_debug_math(hg_ctx* ctx, int32_t kind, int64_t n, const double* x, double* out) {
  if (!ctx || !x || !out || n <= 0 || kind < 0 || kind > 4) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  hg::DBuf<double> dx, dy;
  CK(ctx, dx.alloc((size_t)n)); CK(ctx, dy.alloc((size_t)n));
  CK(ctx, cudaMemcpyAsync(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  TRY(hg::fused_debug_math(ctx, kind, n, dx.p, dy.p));
  CK(ctx, cudaMemcpyAsync(out, dy.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return HG_OK;
}

int hg_flush_l2(hg_ctx* ctx) {
  if (!ctx) return HG_ERR_ARG;
  CK(ctx, cudaSetDevice(ctx->opt.device));
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    CK(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  CK(ctx, cudaMemsetAsync(ctx->flush_buf, 1, ctx->flush_bytes, ctx->stream));
  return HG_OK;
}

}  // extern "C"
