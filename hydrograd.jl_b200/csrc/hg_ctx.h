// Internal definitions shared by the host-side builder and the CUDA translation units.
// Nothing in here is part of the ABI (include/hydrograd_b200.h is).
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/hydrograd_b200.h"
#include "hg_ude.h"

namespace hg {

// HG_DEBUG_TIMING=1: wall time of the host-side preprocessing stages on stderr
struct StageTimer {
  const char* what;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit StageTimer(const char* w) : what(w) {}
  ~StageTimer() {
    if (getenv("HG_DEBUG_TIMING"))
      fprintf(stderr, "[hg] %-28s %8.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};

constexpr double EPS = 2.220446049250313e-16;  // eps(Float64) of utilities/smooth_functions.jl

enum BcType : int32_t { BC_INLETQ = 0, BC_EXITH = 1, BC_WALL = 2, BC_SYMM = 3, BC_HALO = 4 };

// ---------------------------------------------------------------- device buffer (RAII)
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t s = nullptr) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    e = cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
  }
  size_t bytes() const { return n * sizeof(T); }
};

// ---------------------------------------------------------------- constants passed by value to kernels
struct Consts {
  double g, k_n, h_small;
};

// what a consumer kernel needs to wait for the halo of this exchange (library-owned transport); n = 0: nothing to wait for
struct CommWait {
  const unsigned long long* flags = nullptr;   // local flag lines, 16 u64 apart
  unsigned long long epoch = 0;
  int32_t n = 0, from = 0, to = 0;             // neighbours; CTA index range [from, to) of the band (tiles with halo faces)
  int32_t* err = nullptr;
};

// state-dependent Manning's n (hg_set_manning_function); type 0 = off
struct MannFn {
  int32_t type = 0;
  double n_lower = 0.0, n_upper = 0.0, k = 0.0, h_mid = 0.0;
};
#ifdef __CUDACC__
// closures of parameters/process_ManningN_2D.jl:140-213, same operation order (x^9 by squaring as Julia's ^ Int)
// (products feeding a sum are rounded separately, __dmul_rn: the closures are ill-conditioned near the ends of their validity
// range and an FMA-contracted ulp would show up amplified)
__device__ __forceinline__ double manning_closure(const MannFn& m, double h, double umag, double ks) {
  const double span = m.n_upper - m.n_lower;
  if (m.type == HG_MANNING_POWER_LAW) return m.n_lower + __dmul_rn(span, pow(h + 2.220446049250313e-16, -m.k));
  if (m.type == HG_MANNING_SIGMOID) return m.n_lower + span / (1.0 + exp(m.k * (h - m.h_mid)));
  if (m.type == HG_MANNING_INVERSE) return m.n_lower + span / (1.0 + __dmul_rn(m.k, h));
  const double Re = __dmul_rn(umag, h) / 1.0e-6;
  const double h_ks = h / ks;
  const double x = Re / 850.0, x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
  const double alpha = 1.0 / (1.0 + __dmul_rn(x8, x));
  const double r2 = Re / (h_ks * 160.0);
  const double beta = 1.0 / (1.0 + __dmul_rn(r2, r2));
  const double part1 = pow(Re / 24.0, alpha);
  const double part2 = pow(1.8 * log10(Re / 2.1), 2.0 * (1.0 - alpha) * beta);
  const double part3 = pow(2.0 * log10(11.8 * h_ks), 2.0 * (1.0 - alpha) * (1.0 - beta));
  const double f = 1.0 / (part1 * part2 * part3);
  return sqrt(f / 8.0) * pow(h, 1.0 / 6.0) / sqrt(9.81);
}
// n of one cell from its raw state (clamp of semi_discretize_swe_2D.jl:101-106, u = q/h, |U| = sqrt(u^2+v^2) :143-146)
__device__ __forceinline__ double manning_of_state(const MannFn& m, double xi, double qx, double qy, double hst, double ks, double hs) {
  const double h0 = xi + hst;
  const bool dry = h0 <= hs;
  const double h = dry ? hs : h0, u = dry ? 0.0 : qx / h, v = dry ? 0.0 : qy / h;
  return manning_closure(m, h, sqrt(__dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v))), ks);
}
#endif

// Boundary entries, one per boundary face, in the reference's processing order
// (inlet-q boundaries, exit-h, wall, symm; bc_2D.jl:279-295).  Cell ids are in REFERENCE order
// for the plain path and in INTERNAL order for the fused path (two copies of `cell`).
struct BcHost {
  std::vector<int32_t> type, group, ghost, cell_ref;  // [B]
  std::vector<double> nx, ny, l53, l23, hstill_g, zb_g;  // [B]; l53 = L^(5/3), l23 = L^(2/3) (inlet only)
  std::vector<int32_t> inlet_ptr;                     // [n_inletq+1] entry ranges of each inlet boundary
  std::vector<int32_t> bcell_ref, bcell_ptr, bcell_ent; // distinct boundary-adjacent cells (reference ids) -> entries
  std::vector<int32_t> cf_rev;                        // per cell-face: index of the same face in the neighbour's list
  std::vector<int32_t> halo_off, halo_cnt;            // [B] per halo entry: offset of xi in the halo buffers, n_k
  std::vector<int64_t> halo_counts;                   // [n_halo] entries per neighbour
};

// ---------------------------------------------------------------- plain (reference-order) path
struct PlainDev {
  DBuf<int32_t> cf_ptr, cf_nb, cf_rev; // CSR of cell faces; nb >= N means ghost (N + ghost id); rev = same face in nb's list
  DBuf<double> cf_nx, cf_ny, cf_len;  // per cell-face
  DBuf<double> area, hstill, zb, S0x, S0y, mann;  // [N] reference order
  DBuf<int32_t> matid;                // [N]
  DBuf<int32_t> bc_type, bc_group, bc_ghost, bc_cell;  // [B] entry order
  DBuf<double> bc_nx, bc_ny, bc_l53, bc_l23;
  DBuf<int32_t> inlet_ptr;
  DBuf<double> ks;                    // [N] roughness height (variable Manning's n)
  DBuf<double> hstill_g, zb_g;        // [B] GHOST order (as passed in)
  DBuf<double> gh, gqx, gqy, gxi;     // [B] ghost states, ghost order
  DBuf<double> gh_d, gqx_d, gqy_d, gxi_d;  // [K][B] their tangents (forward mode, hg_jvp.cu), one copy per direction of a batch
  DBuf<double> jgh, jgqx, jgqy, jgxi;      // [K][B] ghost values of the forward-mode sweeps (per direction: no write sharing)
  DBuf<double> V, dQd, pdot;          // [3N], [3N], [np]: tangent of the state, of the RHS, of the parameters
  DBuf<double> Qin, wse;              // [n_inletq], [n_exith]
  DBuf<double> Q, dQ, params;         // [3N], [3N], [np]
  DBuf<int32_t> err;                  // device error flag
};

// ---------------------------------------------------------------- fused (tile) path
// Cells are renumbered so that tile t owns the internal cells [t*T, t*T + nc) -- every tile but the
// last has exactly T cells.  Everything a tile needs lives in CONTIGUOUS, 16-byte aligned, padded
// segments of the global arrays so that one CTA can pull its whole working set into shared memory
// with a handful of TMA bulk copies (cp.async.bulk):
//   cells  [c0, c0+nc)                  state (3 comps, stride Ns), hstill, zb, area, mann, S0x, S0y
//   faces  [fp, fp+nfp)                 lr (lL | lR<<16), nx, ny, len; interior faces first, then
//                                       boundary faces (lR = 0xFFFF), then zero-length padding to 4
//   halo   [hp, hp+nh)                  internal ids of the neighbour cells owned by other tiles
//   cf     [t*T*NF, (t+1)*T*NF)         NF slots per cell: local face ids in the reference's face order
//                                       (bit 15 = cell is on the R side); unused slots point at the
//                                       tile's zero-flux slot (index nfp)
// Local cell index space of a tile: owned cells 0..nc-1, halo cells ncp.. (ncp = nc rounded up to 2).
constexpr int kTileDesc = 12;  // ints per tile: c0 nc hp nh fp nf nfp (unused) (unused) nint bfp (pad)
// Host-buffer pipeline: the rows of a chunk start at HOST addresses that are multiples of 256 bytes (32 doubles).  The caller's
// [3N] vectors have an arbitrary N, so each component's chunk boundaries are shifted by up to kPipeAlign - 1 cells; the stage
// tables are built with that margin and do not depend on the pointers.  Misaligned rows cost ~20 % of the duplex PCIe rate
// (scripts/micro/pcie_pipeline.cu: 35.6 -> 44.2 GB/s per direction).
constexpr int64_t kPipeAlign = 32;

struct FusedHost {
  int32_t n_tiles = 0, T = 0, NF = 4, max_local = 0, max_faces = 0, max_halo = 0, max_cell_faces = 0;
  int64_t Ns = 0;                    // padded component stride of cell-indexed arrays
  std::vector<int32_t> perm;         // internal -> reference cell id
  std::vector<int32_t> iperm;        // reference -> internal
  std::vector<int32_t> tile_desc;    // [n_tiles * kTileDesc]
  std::vector<int32_t> halo;
  std::vector<uint32_t> face_lr;
  std::vector<double> face_nx, face_ny, face_len;
  std::vector<int32_t> bface_e;      // boundary entry of each tile's boundary faces (tile order)
  std::vector<uint16_t> cf_idx;      // [n_tiles * T * NF]
  // host-buffer pipeline (hg_rhs): reference-order chunks arrive one by one over PCIe; a tile can run once the
  // chunks holding its cells and halo cells have landed; a chunk can leave once its tiles are done
  int32_t n_chunks = 1;
  int64_t chunk_cells = 0;           // nominal rows per chunk, a multiple of kPipeAlign (see chunk_rows in hg_api.cu)
  std::vector<int32_t> tile_order;   // tiles sorted by the stage at which they become ready
  std::vector<int32_t> stage_ptr;    // [n_chunks+1] ranges of tile_order
  std::vector<int32_t> chunk_done;   // [n_chunks] stage after which every cell of the chunk has been computed
  // multi-GPU overlap: tiles without halo-boundary faces first (they can run while the halo is in flight), then the band
  std::vector<int32_t> band_order;   // [n_tiles]
  int32_t n_interior_tiles = 0;
  std::vector<int32_t> comm_order;   // [n_tiles] library-owned transport: interior tiles, band in the middle, interior tiles
  int32_t comm_band0 = 0;            // first band position in comm_order
};

struct FusedDev {
  DBuf<int32_t> perm, iperm, tile_desc, halo, bface_e, tile_order, band_order, comm_order;
  DBuf<double> stage_out, stage_lam;
  DBuf<uint32_t> face_lr;
  DBuf<uint16_t> cf_idx;
  DBuf<double> face_nx, face_ny, face_len;
  DBuf<double> area, hstill, zb, S0x, S0y, mann;  // [Ns] internal order
  DBuf<double> ks;                                // [Ns] roughness height (variable Manning's n)
  DBuf<int32_t> matid;
  DBuf<uint8_t> matid8;                           // the same zone ids in one byte each (few zones: the zone reduction of the VJP reads these)
  DBuf<int32_t> bc_type, bc_group, bc_cell;        // [B] entry order, internal cell ids
  DBuf<double> bc_nx, bc_ny, bc_l53, bc_l23, bc_hstill, bc_zb;
  DBuf<int32_t> inlet_ptr;
  DBuf<double> inlet_coef;                         // [n_inletq]  Q_k / total_A
  DBuf<double> Qin, wse;
  DBuf<double> Q, Q2, dQ, lam, Qbar, stage, params, pbar, nbar, s0bar;  // state-like: [3*Ns]
  DBuf<double> ent_c, ent_n, ent_z, ent_h, Qinbar, inlet_A, zone_part, zone_part2;   // VJP: per boundary entry / per inlet
  DBuf<int32_t> bcell, bcell_ref, bcell_ptr, bcell_ent;                 // boundary-adjacent cells -> their entries
  DBuf<int32_t> halo_off, halo_cnt;                                     // [B]
  DBuf<double> ens_Q, ens_Q2, ens_mann, ens_Qin, ens_coef, ens_A;       // parameter ensembles: [M][...]
  DBuf<double> rk_k, rk_acc, rk_tmp;                                   // RK4 stages
  DBuf<double> ts_k[7], ts_new, ts_part, ts_sum;                       // Tsit5 stages, candidate state, error-norm scratch
  DBuf<double> halo_send, halo_recv;                                    // [6 * n_halo_entries]
  DBuf<double> j_V, j_out, j_pdot, j_p0, j_p1, j_p2, j_coef;             // fused forward mode (hg_fjvp.cu): tangents in / out, parameter tangents
  DBuf<double> ude_theta, ude_stats, ude_part;                          // UDE network: parameters, LayerNorm statistics, partial sums
  DBuf<int32_t> err;
};

// host copies of the bindable frozen fields (so that un-binding a parameter restores them)
struct Frozen {
  std::vector<double> mann_ref, zb_ref, zbg_ghost, S0_ref, Qin, wse;
};

}  // namespace hg

// library-owned halo exchange (hg_comm.cu)
struct hg_comm {
  int32_t n = 0;                          // neighbours (= halo boundaries of the context)
  int64_t stride = 0, flags_off = 0;      // doubles per parity buffer; byte offset of the flag lines
  size_t bytes = 0;
  void* buf = nullptr;                    // one cudaMalloc: recv[0] | recv[1] | one 128-byte flag line per neighbour
  double* recv[2] = {nullptr, nullptr};
  unsigned long long* flags = nullptr;    // written by the peers: epoch of the last completed push
  hg::DBuf<double*> d_dst0, d_dst1;       // [n] where this rank's block k lands in peer k's parity buffer 0 / 1
  hg::DBuf<unsigned long long*> d_flag;   // [n] this rank's flag line at peer k
  hg::DBuf<int32_t> d_ptr;                // [n+1] entry ranges of the blocks
  std::vector<void*> opened;              // cudaIpcOpenMemHandle mappings
  unsigned long long epoch = 0;           // exchanges issued so far
  bool connected = false, auto_exchange = true, pushed = false;
};

struct hg_ctx {
  hg_options opt{};
  int64_t N = 0, F = 0, B = 0, sumnf = 0;
  int64_t n_inletq = 0, n_exith = 0, n_wall = 0, n_symm = 0, n_mat = 0, nbcell = 0;
  int64_t n_halo = 0, halo_e0 = 0, n_halo_entries = 0;
  cudaStream_t own_stream = nullptr, s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_cmp;
  int n_sm = 148;
  int64_t ens_members = 0;
  bool ens_per_member_mann = false;
  bool lam_set = false;
  hg::Consts c{};
  hg::MannFn mfn{};
  hg::ude::Model ude{};        // hg_set_ude_model
  hg::ude::ThetaMap ude_map{};
  int64_t ude_user_params = 0; // length of the caller's theta
  bool ude_set = false;
  bool controller_fastpow = false;   // hg_set_controller_pow
  std::vector<double> last_steps;    // accepted step sizes of the last hg_solve_tsit5 (hg_last_steps)
  int32_t ab3_step = 1;        // hg_step_ab3: 1, 2 = Ralston start-up steps, 3 = multistep formula
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  bool state_set = false;
  uint64_t state_gen = 0;      // bumped whenever the resident state changes (hg_state_generation): lets a pullback reuse the state of its forward call
  int32_t active = HG_PARAM_NONE;
  int64_t n_params = 0;
  hg::BcHost bch;
  hg::PlainDev pd;
  hg::FusedHost fh;
  bool fjvp_ready = false;     // hg_fjvp.cu: kernel attributes set
  int32_t fh_force_nf = 0;     // build_tiles: face slots per cell of the attempt being built (4 or 8)
  hg_comm* comm = nullptr;
  hg_allreduce_fn allreduce = nullptr;   // sum over ranks on the host (adaptive solves on multi-rank contexts), hg_comm_set_allreduce
  void* allreduce_user = nullptr;   // library-owned halo exchange (hg_comm.cu); null = the caller moves halo_send -> halo_recv
  hg::FusedDev fd;
  double* h_pinned = nullptr;  // staging [6N] pinned host memory
  size_t h_pinned_bytes = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  int64_t device_bytes = 0;
  hg::Frozen fr;
  std::vector<double> last_params;
  int32_t last_active = -1;
  std::vector<int32_t> matid_ref;
};

inline bool hg_comm_ready(const hg_ctx* ctx) { return ctx->comm != nullptr && ctx->comm->connected; }

namespace hg {
// host-side builder (hg_host.cpp)
int build_host(hg_ctx* ctx, const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f,
               std::vector<int32_t>& cf_ptr, std::vector<int32_t>& cf_nb, std::vector<double>& cf_nx,
               std::vector<double>& cf_ny, std::vector<double>& cf_len, std::vector<int32_t>& cf_face);
int build_tiles(hg_ctx* ctx, const hg_mesh_desc* m, const std::vector<int32_t>& cf_ptr,
                const std::vector<int32_t>& cf_nb, const std::vector<double>& cf_nx, const std::vector<double>& cf_ny,
                const std::vector<double>& cf_len, const std::vector<int32_t>& cf_face);

// plain path launchers (hg_plain.cu)
int plain_rhs(hg_ctx* ctx, const double* dQ_in_Q, double* d_out);
// forward mode on the plain tables (hg_jvp.cu)
int plain_jvp(hg_ctx* ctx, const double* d_Q, const double* d_V, const double* d_pdot, double* d_out, double* d_out_dot);
int plain_jvp_batch(hg_ctx* ctx, const double* d_Q, const double* d_V, int64_t sV, const double* d_pdot, int64_t sP, double* d_out,
                    double* d_out_dot, int64_t K);
int sens_lincomb(hg_ctx* ctx, int64_t len, double* y, const double* x, int n, const double* const* k, const double* coef);
int sens_err_blocks(int64_t n3);
int sens_err_norm(hg_ctx* ctx, int64_t n3, int rows, const double* u, const double* unew, int n, const double* const* k,
                  const double* coef, double abstol, double reltol, double* d_part, double* d_sum);
// fused path launchers (hg_fused.cu)
int fused_rhs(hg_ctx* ctx, const double* d_Q, double* d_out, bool euler, double dt);
int fused_rhs_tiles(hg_ctx* ctx, const double* d_Q, double* d_out, int32_t tile_base, int32_t n_tiles, bool with_band = false);
int fused_rhs_phase(hg_ctx* ctx, const double* d_Q, double* d_out, int phase);
int fused_permute_range(hg_ctx* ctx, bool to_internal, const double* src, double* dst, int64_t r0, int64_t r1);
int fused_smem_bytes(const hg_ctx* ctx);
int fused_prepare(hg_ctx* ctx);
bool fused_config_ok(const hg_ctx* ctx);
int fused_permute(hg_ctx* ctx, bool to_internal, const double* src, double* dst);
int fused_bind_manning(hg_ctx* ctx, const double* d_params);
int fused_bind_zb(hg_ctx* ctx, const double* d_params_ref);
int fused_cfg_id(const hg_ctx* ctx);
void fused_inlet_coef(hg_ctx* ctx, const double* d_Q, bool after_push = false);
// fused VJP (hg_vjp.cu)
int fused_vjp_prepare(hg_ctx* ctx, int cfg_id);
int fused_vjp(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_lam, double* d_Qbar);
int fused_vjp_tiles(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_lam, double* d_Qbar, const int32_t* tile_order,
                    int32_t tile_base, int32_t n_run, int comm_mode = 0, bool pdl = false);
int fused_vjp_finish(hg_ctx* ctx, const double* d_Q, double* d_Qbar);
int fused_nbar_to_ref(hg_ctx* ctx, double* d_dst);
int fused_halo_pack(hg_ctx* ctx, bool with_lambda);
int fused_adjoint_step(hg_ctx* ctx, const double* Qn, const double* Qn1, double* lam, double* lam_tmp, double* pbar_acc, int64_t np, double dt);
int fused_acc_pbar(hg_ctx* ctx, int64_t np, double* acc, double a);
int fused_axpy(hg_ctx* ctx, double* y, const double* x, const double* k, double a, const double* acc_in, double* acc_out, double b);
int fused_lincomb(hg_ctx* ctx, double* y, const double* x, int n, const double* const* k, const double* coef);
int fused_err_blocks(const hg_ctx* ctx);
int fused_err_norm(hg_ctx* ctx, const double* u, const double* unew, int n, const double* const* k, const double* coef, double abstol,
                   double reltol, double* d_part, double* d_sum);
// fused forward mode (hg_fjvp.cu)
int fused_jvp(hg_ctx* ctx, int cfg_id, const double* d_Q, const double* d_V, const double* d_pdot, double* d_out, double* d_out_d, int64_t K);
int fused_bed_from(hg_ctx* ctx, const double* d_zb_ref, double* d_zb, double* d_S0x, double* d_S0y);
// library-owned halo exchange (hg_comm.cu): push d_Q (and d_lam) of the cut cells into the neighbours' receive buffers
int comm_push(hg_ctx* ctx, const double* d_Q, const double* d_lam);
int fused_debug_math(hg_ctx* ctx, int32_t kind, int64_t n, const double* d_x, double* d_out);
int fused_rhs_ensemble(hg_ctx* ctx, const double* d_Q, double* d_out, bool euler, double dt);
// UDE closure (hg_ude.cu)
int ude_prepare(hg_ctx* ctx);
int ude_spec_id(const hg_ctx* ctx);
int ude_eval_n(hg_ctx* ctx, const double* d_Q);
int ude_adjoint(hg_ctx* ctx, const double* d_Q, double* d_Qbar);
}  // namespace hg
