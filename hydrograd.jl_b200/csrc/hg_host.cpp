// Host-side builder: turns the reference-shaped flat tables of the ABI (mesh_2D / BoundaryConditions2D
// fields, include/hydrograd_b200.h) into the internal layouts:
//   * a compact cell->face CSR in reference order (plain path), and
//   * a locality-ordered, tiled, face-once layout (fused path): cells renumbered by recursive
//     coordinate bisection so that each CTA tile is a compact patch, per-tile halo lists, per-tile
//     face lists with 16-bit local indices, per-cell local CSR with an orientation bit.
// This replaces the per-call Dict / Vector-of-Vector lookups of compute_inviscid_fluxes
// (semi_discretize_swe_2D.jl:286-330) with one O(N log N) preprocessing step at hg_create.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <omp.h>

#include "hg_ctx.h"

namespace hg {

#define HG_FAIL(ctx, code, ...)                       \
  do {                                                \
    char _b[512];                                     \
    snprintf(_b, sizeof(_b), __VA_ARGS__);            \
    (ctx)->err = _b;                                  \
    return (code);                                    \
  } while (0)

int build_host(hg_ctx* ctx, const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f,
               std::vector<int32_t>& cf_ptr, std::vector<int32_t>& cf_nb, std::vector<double>& cf_nx,
               std::vector<double>& cf_ny, std::vector<double>& cf_len, std::vector<int32_t>& cf_face) {
  StageTimer timer("build_host");
  if (!m || !b || !f) HG_FAIL(ctx, HG_ERR_ARG, "null descriptor");
  const int64_t N = m->n_cells, F = m->n_faces, B = m->n_ghost, ld = m->ld, base = m->index_base;
  if (N <= 0 || F <= 0 || B < 0 || ld <= 0) HG_FAIL(ctx, HG_ERR_ARG, "bad sizes N=%ld F=%ld B=%ld ld=%ld", (long)N, (long)F, (long)B, (long)ld);
  if (N + B >= (int64_t)1 << 31) HG_FAIL(ctx, HG_ERR_ARG, "mesh too large for 32-bit cell ids");
  if (base != 0 && base != 1) HG_FAIL(ctx, HG_ERR_ARG, "index_base must be 0 or 1");
  if (!m->cell_nfaces || !m->cell_faces || !m->cell_neighbors || !m->cell_normals || !m->face_is_boundary ||
      !m->face_lengths || !m->cell_areas)
    HG_FAIL(ctx, HG_ERR_ARG, "null mesh array");
  if (!f->hstill || !f->zb_cells || !f->S0_cells || !f->ManningN_cells || (B > 0 && (!f->hstill_ghost || !f->zb_ghost)))
    HG_FAIL(ctx, HG_ERR_ARG, "null field array");
  if (f->riemann_solver && std::strcmp(f->riemann_solver, "Roe") != 0) {
    // semi_discretize_swe_2D.jl:356-361
    if (!std::strcmp(f->riemann_solver, "HLL") || !std::strcmp(f->riemann_solver, "HLLC"))
      HG_FAIL(ctx, HG_ERR_SOLVER, "%s solver not implemented yet", f->riemann_solver);
    HG_FAIL(ctx, HG_ERR_SOLVER, "Wrong choice of RiemannSolver");
  }
  ctx->N = N; ctx->F = F; ctx->B = B;
  ctx->c = {f->g, f->k_n, f->h_small};
  ctx->n_inletq = b->n_inletq; ctx->n_exith = b->n_exith; ctx->n_wall = b->n_wall; ctx->n_symm = b->n_symm;
  ctx->n_mat = f->n_mat;
  const int64_t nbc = b->n_inletq + b->n_exith + b->n_wall + b->n_symm + b->n_halo;
  ctx->n_halo = b->n_halo;
  if (b->n_halo < 0 || (b->n_halo > 0 && (!b->halo_flip || !b->halo_area))) HG_FAIL(ctx, HG_ERR_ARG, "halo boundary data missing");
  if (b->n_halo > 0 && ctx->opt.path == 1)
    HG_FAIL(ctx, HG_ERR_ARG, "multi-rank contexts (n_halo > 0) need the fused path: the plain / strict path has no halo boundaries");
  if (b->n_inletq < 0 || b->n_exith < 0 || b->n_wall < 0 || b->n_symm < 0) HG_FAIL(ctx, HG_ERR_ARG, "negative boundary count");
  if (B > 0 && (!b->bc_ptr || !b->ghost_ids || !b->internal_cells || !b->outward_normals))
    HG_FAIL(ctx, HG_ERR_ARG, "null boundary array");
  if (b->n_inletq > 0 && (!b->face_lengths || !f->inletQ_TotalQ)) HG_FAIL(ctx, HG_ERR_ARG, "inlet-q data missing");
  if (b->n_exith > 0 && !f->exitH_WSE) HG_FAIL(ctx, HG_ERR_ARG, "exit-h data missing");

  // ---- cell -> face CSR in reference order
  cf_ptr.assign(N + 1, 0);
  for (int64_t i = 0; i < N; ++i) {
    int64_t nf = m->cell_nfaces[i];
    if (nf < 1 || nf > ld) HG_FAIL(ctx, HG_ERR_ARG, "cell %ld has %ld faces (ld=%ld)", (long)i, (long)nf, (long)ld);
    cf_ptr[i + 1] = cf_ptr[i] + (int32_t)nf;
  }
  const int64_t S = cf_ptr[N];
  ctx->sumnf = S;
  // (value-initialising ~2 GB of tables on a 16M-cell mesh costs a second on one thread: one thread per table)
#pragma omp parallel sections
  {
#pragma omp section
    cf_nb.resize(S);
#pragma omp section
    cf_nx.resize(S);
#pragma omp section
    cf_ny.resize(S);
#pragma omp section
    cf_len.resize(S);
#pragma omp section
    cf_face.resize(S);
  }
  int64_t bad_cell = -1, bad_face = -1;
  int bad_kind = 0;   // 1 face id, 2 ghost id, 3 neighbour id
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    for (int64_t j = 0; j < cf_ptr[i + 1] - cf_ptr[i]; ++j) {
      int64_t fv = m->cell_faces[i + N * j];
      int64_t fid = (fv < 0 ? -fv : fv) - base;
      int kind = 0;
      int64_t nb = m->cell_neighbors[i + N * j] - base;
      int64_t k = cf_ptr[i] + j;
      if (fid < 0 || fid >= F) kind = 1;
      else if (m->face_is_boundary[fid]) {
        if (nb < 0 || nb >= B) kind = 2;
        else cf_nb[k] = (int32_t)(N + nb);
      } else {
        if (nb < 0 || nb >= N || nb == i) kind = 3;
        else cf_nb[k] = (int32_t)nb;
      }
      if (kind) {
#pragma omp critical(hg_csr_fail)
        if (bad_cell < 0 || i < bad_cell) { bad_cell = i; bad_face = j; bad_kind = kind; }
        continue;
      }
      cf_nx[k] = m->cell_normals[i + N * (j + ld * 0)];
      cf_ny[k] = m->cell_normals[i + N * (j + ld * 1)];
      cf_len[k] = m->face_lengths[fid];
      cf_face[k] = (int32_t)fid;
    }
  }
  if (bad_cell >= 0)
    HG_FAIL(ctx, HG_ERR_ARG, "cell %ld face %ld: %s id out of range", (long)bad_cell, (long)bad_face,
            bad_kind == 1 ? "face" : bad_kind == 2 ? "ghost" : "neighbour");

  // ---- boundary entries in processing order
  BcHost& h = ctx->bch;
  h = BcHost();
  if (nbc > 0 && b->bc_ptr[0] != 0) HG_FAIL(ctx, HG_ERR_ARG, "bc_ptr[0] must be 0");
  if ((nbc > 0 ? b->bc_ptr[nbc] : 0) != B) HG_FAIL(ctx, HG_ERR_ARG, "boundary entries (%ld) do not cover the %ld ghost cells", (long)(nbc > 0 ? b->bc_ptr[nbc] : 0), (long)B);
  h.type.resize(B); h.group.resize(B); h.ghost.resize(B); h.cell_ref.resize(B);
  h.nx.resize(B); h.ny.resize(B); h.l53.assign(B, 0.0); h.l23.assign(B, 0.0); h.hstill_g.resize(B); h.zb_g.resize(B);
  h.inlet_ptr.assign(1, 0);
  std::vector<char> seen(B, 0);
  int64_t kb = 0;
  const int64_t counts[5] = {b->n_inletq, b->n_exith, b->n_wall, b->n_symm, b->n_halo};
  h.halo_off.assign(B, 0); h.halo_cnt.assign(B, 0);
  ctx->halo_e0 = B; ctx->n_halo_entries = 0;
  for (int t = 0; t < 5; ++t) {
    for (int64_t kk = 0; kk < counts[t]; ++kk, ++kb) {
      if (b->bc_ptr[kb + 1] < b->bc_ptr[kb]) HG_FAIL(ctx, HG_ERR_ARG, "bc_ptr not monotone");
      for (int64_t e = b->bc_ptr[kb]; e < b->bc_ptr[kb + 1]; ++e) {
        int64_t gi = b->ghost_ids[e] - base, c = b->internal_cells[e] - base;
        if (gi < 0 || gi >= B || seen[gi]) HG_FAIL(ctx, HG_ERR_ARG, "boundary entry %ld: bad or repeated ghost id", (long)e);
        if (c < 0 || c >= N) HG_FAIL(ctx, HG_ERR_ARG, "boundary entry %ld: bad internal cell", (long)e);
        seen[gi] = 1;
        h.type[e] = t; h.group[e] = (int32_t)kk; h.ghost[e] = (int32_t)gi; h.cell_ref[e] = (int32_t)c;
        h.nx[e] = b->outward_normals[e]; h.ny[e] = b->outward_normals[B + e];
        h.hstill_g[e] = f->hstill_ghost[gi]; h.zb_g[e] = f->zb_ghost[gi];
        if (t == BC_INLETQ) {
          double L = b->face_lengths[e];
          h.l53[e] = std::pow(L, 5.0 / 3.0);  // bc_2D.jl:674
          h.l23[e] = std::pow(L, 2.0 / 3.0);  // bc_2D.jl:691
        }
        if (t == BC_HALO) {
          const int64_t nk = b->bc_ptr[kb + 1] - b->bc_ptr[kb];
          if (ctx->halo_e0 == B) ctx->halo_e0 = b->bc_ptr[kb];
          h.group[e] = b->halo_flip[e] ? 1 : 0;          // the flip flag rides in `group`
          h.l23[e] = b->halo_area[e];                    // ... and the remote cell's area in `l23`
          h.halo_off[e] = (int32_t)(6 * (b->bc_ptr[kb] - ctx->halo_e0) + (e - b->bc_ptr[kb]));
          h.halo_cnt[e] = (int32_t)nk;
          if (!(b->halo_area[e] > 0.0)) HG_FAIL(ctx, HG_ERR_ARG, "halo entry %ld: bad remote cell area", (long)e);
        }
      }
      if (t == BC_HALO) h.halo_counts.push_back(b->bc_ptr[kb + 1] - b->bc_ptr[kb]);
      if (t == BC_INLETQ) h.inlet_ptr.push_back((int32_t)b->bc_ptr[kb + 1]);
    }
  }
  ctx->n_halo_entries = B - ctx->halo_e0;
  // consistency: the ghost of entry e must be the neighbour of its internal cell
  for (int64_t e = 0; e < B; ++e) {
    int32_t c = h.cell_ref[e];
    bool ok = false;
    for (int32_t k = cf_ptr[c]; k < cf_ptr[c + 1]; ++k) ok |= (cf_nb[k] == (int32_t)(N + h.ghost[e]));
    if (!ok) HG_FAIL(ctx, HG_ERR_ARG, "boundary entry %ld: ghost %d is not a neighbour of cell %d", (long)e, h.ghost[e], c);
  }
  // ---- per cell-face: where the same face sits in the neighbour's list (transposed Green-Gauss of the VJP)
  h.cf_rev.assign(S, -1);
  {
    // for every interior cell-face: the position of the same face in the neighbour's list (scan of <= ld entries), in parallel
    int64_t bad = -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
      for (int32_t k = cf_ptr[i]; k < cf_ptr[i + 1]; ++k) {
        const int32_t nb = cf_nb[k];
        if (nb >= N) continue;
        int32_t kk = -1;
        for (int32_t q = cf_ptr[nb]; q < cf_ptr[nb + 1]; ++q)
          if (cf_face[q] == cf_face[k] && cf_nb[q] == (int32_t)i) { kk = q; break; }
        if (kk < 0) {
#pragma omp critical(hg_cfrev_fail)
          if (bad < 0 || k < bad) bad = k;
        }
        h.cf_rev[k] = kk;
      }
    if (bad >= 0) HG_FAIL(ctx, HG_ERR_ARG, "interior face %d is listed by one cell only", cf_face[bad]);
  }
  // ---- distinct boundary-adjacent cells -> their boundary entries (deterministic scatter of BC adjoints)
  {
    std::vector<int32_t> order(B);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return h.cell_ref[x] < h.cell_ref[y]; });
    h.bcell_ptr.assign(1, 0);
    for (int64_t q = 0; q < B; ++q) {
      const int32_t e = order[q];
      if (q == 0 || h.cell_ref[e] != h.cell_ref[order[q - 1]]) {
        if (q) h.bcell_ptr.push_back((int32_t)h.bcell_ent.size());
        h.bcell_ref.push_back(h.cell_ref[e]);
      }
      h.bcell_ent.push_back(e);
    }
    if (B) h.bcell_ptr.push_back((int32_t)h.bcell_ent.size());
    ctx->nbcell = (int64_t)h.bcell_ref.size();
  }
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------
namespace {
struct Rcb {
  const double* cx;
  const double* cy;
  int32_t T;
  std::vector<int32_t>& idx;
  bool ok = true;   // every leaf starts at a multiple of T
  void run(int64_t lo, int64_t hi) {
    const int64_t n = hi - lo;
    if (n <= T) {
      std::sort(idx.begin() + lo, idx.begin() + hi);
      if (lo % T != 0) ok = false;
      return;
    }
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int64_t i = lo; i < hi; ++i) {
      double x = cx[idx[i]], y = cy[idx[i]];
      x0 = std::min(x0, x); x1 = std::max(x1, x); y0 = std::min(y0, y); y1 = std::max(y1, y);
    }
    const int64_t k = (n + T - 1) / T;       // leaves below this node
    const int64_t mid = lo + (k / 2) * (int64_t)T;  // all leaves but the last are exactly T cells
    const double* key = (x1 - x0 >= y1 - y0) ? cx : cy;
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [key](int32_t a, int32_t b) {
      return key[a] < key[b] || (key[a] == key[b] && a < b);
    });
    // the two halves are independent (disjoint index ranges): OpenMP tasks down to ~64K cells
    if (n > 65536) {
#pragma omp task shared(idx)
      run(lo, mid);
#pragma omp task shared(idx)
      run(mid, hi);
#pragma omp taskwait
    } else {
      run(lo, mid);
      run(mid, hi);
    }
  }
};
}  // namespace

// Kuhn's augmenting path on the 16 x 16 bank graph: L bank l looks for an R bank, preferring the pair with most faces left
static bool bank_augment(int l, uint32_t& seen, const uint32_t* adj, int* matchR, int* matchL, const int (*cnt)[16]) {
  // adj[l]: bit r set while faces (l, r) are left.  Candidates = the R banks unseen on entry, visited by (faces left desc,
  // bank asc); `seen` only grows, so picking the best still-unseen one each time visits them in exactly that order without
  // building and sorting a list.  The bank graph is sparse (two or three R banks per L bank), hence the bit masks.
  const uint32_t cand = adj[l] & ~seen;
  for (;;) {
    int r = -1;
    for (uint32_t q = cand & ~seen; q; q &= q - 1) {
      const int b = __builtin_ctz(q);
      if (r < 0 || cnt[l][b] > cnt[l][r]) r = b;
    }
    if (r < 0) return false;
    seen |= 1u << r;
    if (matchR[r] < 0 || bank_augment(matchR[r], seen, adj, matchR, matchL, cnt)) { matchR[r] = l; matchL[l] = r; return true; }
  }
}

static int build_tiles_T(hg_ctx* ctx, const hg_mesh_desc* m, const std::vector<int32_t>& cf_ptr,
                         const std::vector<int32_t>& cf_nb, const std::vector<double>& cf_nx, const std::vector<double>& cf_ny,
                         const std::vector<double>& cf_len, const std::vector<int32_t>& cf_face);

// Tile the mesh for one of the compiled kernel configurations (hg_fused.cu): the requested tile size if
// its tiles fit the configuration's shared-memory caps, else the small general configuration (T = 128,
// up to 8 faces per cell).
int build_tiles(hg_ctx* ctx, const hg_mesh_desc* m, const std::vector<int32_t>& cf_ptr,
                const std::vector<int32_t>& cf_nb, const std::vector<double>& cf_nx, const std::vector<double>& cf_ny,
                const std::vector<double>& cf_len, const std::vector<int32_t>& cf_face) {
  int32_t want = ctx->opt.tile_cells > 0 ? ctx->opt.tile_cells : 256;
  if (want != 128 && want != 192 && want != 224 && want != 240 && want != 256 && want != 384 && want != 512) HG_FAIL(ctx, HG_ERR_ARG, "tile_cells must be 128, 192, 224, 240, 256, 384 or 512");
  int32_t maxnf = 0;
  for (int64_t i = 0; i < ctx->N; ++i) maxnf = std::max(maxnf, cf_ptr[i + 1] - cf_ptr[i]);
  if (maxnf > 4) want = 128;
  // attempts: the requested tile size, then T = 128 with the slot count the mesh needs (4 for triangles / quadrilaterals),
  // then T = 128 with the roomy 8-slot configuration
  const int32_t try_T[3] = {want, 128, 128};
  const int32_t try_nf[3] = {maxnf <= 4 ? 4 : 8, maxnf <= 4 ? 4 : 8, 8};
  for (int attempt = 0; attempt < 3; ++attempt) {
    if (attempt > 0 && try_T[attempt] == try_T[attempt - 1] && try_nf[attempt] == try_nf[attempt - 1]) continue;
    ctx->opt.tile_cells = try_T[attempt];
    ctx->fh_force_nf = try_nf[attempt];
    int rc = build_tiles_T(ctx, m, cf_ptr, cf_nb, cf_nx, cf_ny, cf_len, cf_face);
    if (rc != HG_OK) return rc;
    if (fused_config_ok(ctx)) return HG_OK;
  }
  HG_FAIL(ctx, HG_ERR_ARG, "mesh does not fit the compiled tile configurations (tile needs %d local cells, %d faces): "
          "renumber the mesh for locality or pass cell_centroids", ctx->fh.max_local, ctx->fh.max_faces);
}

static int build_tiles_T(hg_ctx* ctx, const hg_mesh_desc* m, const std::vector<int32_t>& cf_ptr,
                         const std::vector<int32_t>& cf_nb, const std::vector<double>& cf_nx, const std::vector<double>& cf_ny,
                         const std::vector<double>& cf_len, const std::vector<int32_t>& cf_face) {
  const int64_t N = ctx->N, F = ctx->F, B = ctx->B;
  FusedHost& fh = ctx->fh;
  fh = FusedHost();
  const int32_t T = ctx->opt.tile_cells;
  fh.T = T;
  int32_t maxnf = 0;
  for (int64_t i = 0; i < N; ++i) maxnf = std::max(maxnf, cf_ptr[i + 1] - cf_ptr[i]);
  fh.max_cell_faces = maxnf;
  fh.NF = ctx->fh_force_nf ? ctx->fh_force_nf : (maxnf <= 4 ? 4 : 8);
  const int32_t NF = fh.NF;
  if (maxnf > 8) HG_FAIL(ctx, HG_ERR_ARG, "cells with %d faces are not supported (max 8 = gMax_Nodes_per_Element)", maxnf);
  fh.Ns = ((N + 15) / 16) * 16 + 16;  // slack: the last tile's TMA copy may read one element past N

  // ---- ordering: tile t = internal cells [t*T, min(N, (t+1)*T))
  fh.perm.resize(N);
  std::iota(fh.perm.begin(), fh.perm.end(), 0);
  if (ctx->opt.reorder && m->cell_centroids) {
    StageTimer timer("rcb renumbering");
    Rcb r{m->cell_centroids, m->cell_centroids + N, T, fh.perm};
#pragma omp parallel
#pragma omp single
    r.run(0, N);
    if (!r.ok) HG_FAIL(ctx, HG_ERR_ARG, "internal error: an RCB leaf does not start at a multiple of the tile size");
  }
  fh.n_tiles = (int32_t)((N + T - 1) / T);
  fh.iperm.resize(N);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) fh.iperm[fh.perm[i]] = (int32_t)i;

  std::vector<int32_t> ghost_entry(B);
  for (int64_t e = 0; e < B; ++e) ghost_entry[ctx->bch.ghost[e]] = (int32_t)e;

  fh.tile_desc.assign((size_t)fh.n_tiles * kTileDesc, 0);
  fh.cf_idx.assign((size_t)fh.n_tiles * T * NF, (uint16_t)0xFFFFu);   // "unset": pass C replaces what is left by the zero-flux slot

  // Every tile is built independently (OpenMP) into its own TileOut with tile-local scratch -- a cell is "in the tile" iff
  // its internal id lies in [c0, c1), halo cells and face ids are looked up in small sorted lists -- and the per-tile pieces
  // are concatenated afterwards at offsets from a prefix sum: same tables as a serial pass, in a fraction of the time.
  struct TileOut {
    std::vector<int32_t> halo, bface_e;
    std::vector<uint32_t> face_lr;
    std::vector<double> nx, ny, len;
    int32_t nh = 0, nint = 0, nf = 0, nfp = 0, max_local = 0;
  };
  std::vector<TileOut> outs(fh.n_tiles);
  int first_err = HG_OK;
  std::string first_msg;
  auto tile_fail = [&](int code, const std::string& msg) {
#pragma omp critical(hg_tile_fail)
    if (first_err == HG_OK) { first_err = code; first_msg = msg; }
  };
  StageTimer tiles_timer("tile tables");
  double TT[6] = {0, 0, 0, 0, 0, 0};   // thread-seconds per stage of the tile loop (printed with HG_DEBUG_TIMING)
  // sA / sB: the gather-map slots (local cell * NF + position in the cell's face list) of the one or two OWNED cells of the
  // face -- pass C fills them once the face's place in the tile is known
  struct TF { int32_t lL, lR, fid, sA, sB; double nx, ny, len; };
#pragma omp parallel
  {
  // per-thread scratch, reused from tile to tile (the allocator was a quarter of this loop)
  std::vector<TF> tf, sorted, out;
  std::vector<uint64_t> key;
  std::vector<int32_t> bucket;
#pragma omp for schedule(dynamic, 16)
  for (int32_t t = 0; t < fh.n_tiles; ++t) {
    double tt0 = omp_get_wtime(), tt1;
#define TTM(i) do { tt1 = omp_get_wtime(); _Pragma("omp atomic") TT[i] += tt1 - tt0; tt0 = tt1; } while (0)
    TileOut& o = outs[t];
    const int32_t c0 = t * T, c1 = (int32_t)std::min<int64_t>(N, (int64_t)c0 + T), nc = c1 - c0, ncp = (nc + 1) & ~1;
    // pass 0: the one-layer halo, in ascending internal order (so that the indirect loads of neighbouring lanes
    // fall into the same sectors wherever the neighbouring tiles' cells are contiguous)
    for (int32_t c = c0; c < c1; ++c) {
      const int32_t r = fh.perm[c];
      for (int32_t k = cf_ptr[r]; k < cf_ptr[r + 1]; ++k) {
        if (cf_nb[k] >= N) continue;
        const int32_t cn = fh.iperm[cf_nb[k]];
        if (cn < c0 || cn >= c1) o.halo.push_back(cn);
      }
    }
    std::sort(o.halo.begin(), o.halo.end());
    o.halo.erase(std::unique(o.halo.begin(), o.halo.end()), o.halo.end());
    o.nh = (int32_t)o.halo.size();
    const int32_t nloc = ncp + o.nh;
    auto loc = [&](int32_t c) {   // tile-local index of an owned or halo cell
      return (c >= c0 && c < c1) ? c - c0 : ncp + (int32_t)(std::lower_bound(o.halo.begin(), o.halo.begin() + o.nh, c) - o.halo.begin());
    };
    TTM(0);
    // pass A: interior faces, found by the first owned cell (internal order) that sees them ...
    tf.clear();
    bool bad = false;
    for (int32_t c = c0; c < c1 && !bad; ++c) {
      const int32_t r = fh.perm[c];
      for (int32_t k = cf_ptr[r]; k < cf_ptr[r + 1]; ++k) {
        if (cf_nb[k] >= N) continue;
        const int32_t fid = cf_face[k];
        const int32_t rn = cf_nb[k], cn = fh.iperm[rn];
        if (cn >= c0 && cn < c) continue;   // both cells owned: the face was taken when the smaller internal id saw it
        // canonical orientation: L = smaller REFERENCE id, normal taken from the L cell's own table
        int32_t lL, lR; double nx, ny;
        const bool cn_owned = cn >= c0 && cn < c1;
        int32_t kk = -1;   // the face in the neighbour's own list
        if (r > rn || cn_owned) {
          for (int32_t q = cf_ptr[rn]; q < cf_ptr[rn + 1]; ++q) if (cf_face[q] == fid) { kk = q; break; }
          if (kk < 0 || cf_nb[kk] != r) {
            tile_fail(HG_ERR_ARG, "face " + std::to_string(fid) + ": cells " + std::to_string(r) + " and " + std::to_string(rn) + " disagree on adjacency");
            bad = true;
            break;
          }
        }
        if (r < rn) {
          lL = loc(c); lR = loc(cn); nx = cf_nx[k]; ny = cf_ny[k];
        } else {
          lL = loc(cn); lR = loc(c); nx = cf_nx[kk]; ny = cf_ny[kk];
        }
        tf.push_back({lL, lR, fid, (c - c0) * NF + (k - cf_ptr[r]), cn_owned ? (cn - c0) * NF + (kk - cf_ptr[rn]) : -1, nx, ny, cf_len[k]});
      }
    }
    if (bad) continue;
    TTM(1);
    // ... then ordered by (lR - lL, lL): faces with the same index offset are consecutive, so a warp's L reads
    // and R reads of the cell arrays in shared memory each hit consecutive addresses (no bank conflicts), and so
    // do the per-cell flux gathers of phase 3.  Evaluation order per cell is unaffected (bitwise same result).
    {   // (sorting 64-bit keys and permuting once is twice as fast as sorting the records with a comparator)
      key.resize(tf.size());
      for (size_t k = 0; k < tf.size(); ++k)
        key[k] = ((uint64_t)(uint32_t)(tf[k].lR - tf[k].lL + 0x10000) << 40) | ((uint64_t)(uint32_t)tf[k].lL << 20) | (uint64_t)k;
      std::sort(key.begin(), key.end());
      sorted.resize(tf.size());
      for (size_t k = 0; k < tf.size(); ++k) sorted[k] = tf[key[k] & 0xFFFFFu];
      tf.swap(sorted);
    }
    // ... and regrouped into aligned blocks of 16 whose L cells fall into 16 distinct shared-memory banks
    // (8-byte banks: local index mod 16) and whose R cells do too: a 64-bit shared-memory load of a warp is served
    // per half-warp, so the 12 (RHS) / 22 (VJP) per-side gathers of the face phase become conflict-free.  Greedy
    // decomposition of the bipartite multigraph (bank of L) x (bank of R) into matchings; leftovers fill the tail.
    TTM(2);
    if (ctx->opt.reserved[4] == 0 && tf.size() >= 32) {
      // bucket (bL, bR): the faces with that pair of banks in their original order (one counting sort into a flat array --
      // 256 small vectors per tile were 16M allocations on a 16M-cell mesh); cnt = how many are still unplaced
      int head[16][16] = {}, cnt[16][16] = {}, off[16][16];
      for (int32_t k = 0; k < (int32_t)tf.size(); ++k) ++cnt[tf[k].lL & 15][tf[k].lR & 15];
      {
        int run = 0;
        for (int l = 0; l < 16; ++l) for (int r = 0; r < 16; ++r) { off[l][r] = run; run += cnt[l][r]; }
      }
      bucket.resize(tf.size());
      {
        int fill[16][16] = {};
        for (int32_t k = 0; k < (int32_t)tf.size(); ++k) { const int l = tf[k].lL & 15, r = tf[k].lR & 15; bucket[off[l][r] + fill[l][r]++] = k; }
      }
      out.clear();
      out.reserve(tf.size());
      size_t left = tf.size();
      int degL[16] = {};   // unplaced faces per L bank
      uint32_t adj[16] = {};
      for (int l = 0; l < 16; ++l) for (int r = 0; r < 16; ++r) { degL[l] += cnt[l][r]; if (cnt[l][r] > 0) adj[l] |= 1u << r; }
      while (left > 0) {
        // maximum bipartite matching L banks -> R banks over the unplaced faces (Kuhn's augmenting paths, 16 x 16)
        int matchR[16], matchL[16], order[16];
        std::fill(matchR, matchR + 16, -1);
        std::fill(matchL, matchL + 16, -1);
        std::iota(order, order + 16, 0);
        std::stable_sort(order, order + 16, [&](int x, int y) { return degL[x] > degL[y]; });   // fullest banks first
        for (int ol = 0; ol < 16; ++ol) {
          if (degL[order[ol]] == 0) continue;
          uint32_t seen = 0;
          bank_augment(order[ol], seen, adj, matchR, matchL, cnt);
        }
        int npick = 0, useL[16] = {}, useR[16] = {};
        for (int l = 0; l < 16; ++l) {
          const int r = matchL[l];
          if (r < 0) continue;
          out.push_back(tf[bucket[off[l][r] + head[l][r]++]]);
          if (--cnt[l][r] == 0) adj[l] &= ~(1u << r);
          --degL[l]; ++useL[l]; ++useR[r];
          ++npick;
        }
        // incomplete matching (only near the end of a tile): fill the block with the faces that add the fewest conflicts
        const int want = (int)std::min<size_t>(16, left);
        while (npick < want) {
          int bl = -1, br = -1, best = 1 << 30;
          for (int l = 0; l < 16; ++l)
            for (int r = 0; r < 16; ++r) {
              if (cnt[l][r] == 0) continue;
              const int cost = 64 * std::max(useL[l], useR[r]) + 4 * (useL[l] + useR[r]) - std::min(cnt[l][r], 3);
              if (cost < best) { best = cost; bl = l; br = r; }
            }
          out.push_back(tf[bucket[off[bl][br] + head[bl][br]++]]);
          if (--cnt[bl][br] == 0) adj[bl] &= ~(1u << br);
          --degL[bl]; ++useL[bl]; ++useR[br];
          ++npick;
        }
        left -= npick;
      }
      tf.swap(out);
      // (Tried: also permuting the faces INSIDE each block so that the phase-3 slot gathers -- the j-th faces of 16
      // consecutive cells -- hit distinct banks; a greedy placement left 1.75 wavefronts per half-warp vs 1.77, since one
      // collision per group already costs a wavefront and late blocks have no freedom left.  Not kept.)
    }
    if (getenv("HG_DEBUG_TILES") && t == fh.n_tiles / 2) {
      // average number of wavefronts a half-warp's 64-bit gather of the L cells / R cells takes (1 = conflict-free)
      double wl = 0, wr = 0; int nb = 0;
      for (size_t k0 = 0; k0 + 16 <= tf.size(); k0 += 16, ++nb) {
        int cl[16] = {0}, cr[16] = {0}, ml = 0, mr = 0;
        for (size_t k = k0; k < k0 + 16; ++k) { ml = std::max(ml, ++cl[tf[k].lL & 15]); mr = std::max(mr, ++cr[tf[k].lR & 15]); }
        wl += ml; wr += mr;
      }
      fprintf(stderr, "[hg] tile %d: %zu interior faces, half-warp gather wavefronts L %.2f R %.2f\n", t, tf.size(), wl / nb, wr / nb);
    }
    TTM(3);
    // pass C (fused into the emission of the faces): NF slots per cell, local face ids in the reference's face order with the
    // side bit (set when the cell is the R side); unused slots point at the zero-flux slot nfp (adding 0.0 last leaves the
    // left-to-right sum unchanged)
    uint16_t* slots = &fh.cf_idx[(size_t)t * T * NF];
    const size_t nface_max = tf.size() + (size_t)nc * NF + 4;
    o.face_lr.reserve(nface_max); o.nx.reserve(nface_max); o.ny.reserve(nface_max); o.len.reserve(nface_max);
    for (const TF& f : tf) {
      const int32_t lf = (int32_t)o.face_lr.size();
      slots[f.sA] = (uint16_t)(lf | (f.lR == f.sA / NF ? 0x8000 : 0));
      if (f.sB >= 0) slots[f.sB] = (uint16_t)(lf | (f.lR == f.sB / NF ? 0x8000 : 0));
      o.face_lr.push_back((uint32_t)f.lL | ((uint32_t)f.lR << 16));
      o.nx.push_back(f.nx); o.ny.push_back(f.ny); o.len.push_back(f.len);
    }
    o.nint = (int32_t)o.face_lr.size();
    // pass B: boundary faces (their ghost state is evaluated on the fly from the owned internal cell)
    for (int32_t c = c0; c < c1; ++c) {
      const int32_t r = fh.perm[c];
      for (int32_t k = cf_ptr[r]; k < cf_ptr[r + 1]; ++k) {
        if (cf_nb[k] < N) continue;
        slots[(c - c0) * NF + (k - cf_ptr[r])] = (uint16_t)o.face_lr.size();
        o.face_lr.push_back((uint32_t)(c - c0) | (0xFFFFu << 16));
        o.bface_e.push_back(ghost_entry[cf_nb[k] - N]);
        o.nx.push_back(cf_nx[k]); o.ny.push_back(cf_ny[k]); o.len.push_back(cf_len[k]);
      }
    }
    o.nf = (int32_t)o.face_lr.size();
    if (nloc >= 0xFFFF || o.nf >= 0x8000) {
      tile_fail(HG_ERR_ARG, "tile " + std::to_string(t) + " too large (local cells " + std::to_string(nloc) + ", faces " + std::to_string(o.nf) + ")");
      continue;
    }
    while (o.face_lr.size() % 4) {  // zero-length padding faces keep the segments 16-byte sized
      o.face_lr.push_back(0u); o.nx.push_back(1.0); o.ny.push_back(0.0); o.len.push_back(0.0);
    }
    o.nfp = (int32_t)o.face_lr.size();
    for (int32_t l = 0; l < T * NF; ++l)
      if (slots[l] == 0xFFFFu) slots[l] = (uint16_t)o.nfp;
    if (getenv("HG_DEBUG_TILES") && t == fh.n_tiles / 2) {
      // average wavefronts of a half-warp's gather of one flux row at slot j (1 = conflict-free; same face = broadcast)
      double w = 0; int ng = 0;
      for (int32_t l0 = 0; l0 + 16 <= nc; l0 += 16)
        for (int32_t j = 0; j < NF; ++j) {
          int cnt[16] = {0}, mx = 0;
          int32_t seen[16]; int ns = 0;
          for (int32_t l = l0; l < l0 + 16; ++l) {
            const int32_t f = slots[l * NF + j] & 0x7FFF;
            bool dup = false;
            for (int q = 0; q < ns; ++q) dup |= seen[q] == f;
            if (dup) continue;
            seen[ns++] = f;
            mx = std::max(mx, ++cnt[f & 15]);
          }
          w += mx; ++ng;
        }
      fprintf(stderr, "[hg] tile %d: phase-3 slot gather wavefronts per half-warp %.2f\n", t, w / std::max(ng, 1));
    }
    while (o.halo.size() % 4) o.halo.push_back(0);
    o.max_local = ncp + o.nh;
    TTM(4);
  }
  }   // omp parallel
  if (getenv("HG_DEBUG_TIMING")) fprintf(stderr, "[hg] tile stages (thread-seconds): halo %.2f passA %.2f sort %.2f banks %.2f rest %.2f\n", TT[0], TT[1], TT[2], TT[3], TT[4]);
  if (first_err != HG_OK) { ctx->err = first_msg; return first_err; }
  // ---- concatenate at prefix-sum offsets
  {
    StageTimer cat_timer("  concatenate");
    std::vector<size_t> fb(fh.n_tiles + 1, 0), hb(fh.n_tiles + 1, 0), bb(fh.n_tiles + 1, 0);
    for (int32_t t = 0; t < fh.n_tiles; ++t) {
      fb[t + 1] = fb[t] + outs[t].face_lr.size(); hb[t + 1] = hb[t] + outs[t].halo.size(); bb[t + 1] = bb[t] + outs[t].bface_e.size();
    }
#pragma omp parallel sections
    {
#pragma omp section
      { fh.face_lr.resize(fb.back()); fh.halo.resize(hb.back()); fh.bface_e.resize(bb.back()); }
#pragma omp section
      fh.face_nx.resize(fb.back());
#pragma omp section
      fh.face_ny.resize(fb.back());
#pragma omp section
      fh.face_len.resize(fb.back());
    }
    if (fb.back() >= ((size_t)1 << 31) || hb.back() >= ((size_t)1 << 31)) HG_FAIL(ctx, HG_ERR_ARG, "tile tables exceed 32-bit offsets");
#pragma omp parallel for schedule(static)
    for (int32_t t = 0; t < fh.n_tiles; ++t) {
      TileOut& o = outs[t];
      std::copy(o.face_lr.begin(), o.face_lr.end(), fh.face_lr.begin() + fb[t]);
      std::copy(o.nx.begin(), o.nx.end(), fh.face_nx.begin() + fb[t]);
      std::copy(o.ny.begin(), o.ny.end(), fh.face_ny.begin() + fb[t]);
      std::copy(o.len.begin(), o.len.end(), fh.face_len.begin() + fb[t]);
      std::copy(o.halo.begin(), o.halo.end(), fh.halo.begin() + hb[t]);
      std::copy(o.bface_e.begin(), o.bface_e.end(), fh.bface_e.begin() + bb[t]);
      const int32_t c0 = t * T, nc = (int32_t)std::min<int64_t>(N, (int64_t)c0 + T) - c0;
      int32_t* d = &fh.tile_desc[(size_t)t * kTileDesc];
      d[0] = c0; d[1] = nc; d[2] = (int32_t)hb[t]; d[3] = o.nh; d[4] = (int32_t)fb[t]; d[5] = o.nf; d[6] = o.nfp;
      d[7] = 0; d[8] = 0; d[9] = o.nint; d[10] = (int32_t)bb[t]; d[11] = 0;
      TileOut().halo.swap(o.halo);   // (the locals go out of scope with `outs`; nothing else to free early)
    }
    for (int32_t t = 0; t < fh.n_tiles; ++t) {
      fh.max_local = std::max(fh.max_local, outs[t].max_local);
      fh.max_faces = std::max(fh.max_faces, outs[t].nfp);
      fh.max_halo = std::max(fh.max_halo, outs[t].nh);
    }
  }
  fh.max_local = (fh.max_local + 1) & ~1;
  if (getenv("HG_DEBUG_FINGERPRINT")) {   // fingerprint of the tile tables (to check that a change of the builder leaves them as they were)
    auto fnv = [](const void* p, size_t n) { uint64_t h = 1469598103934665603ull; const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } return h; };
    fprintf(stderr, "[hg] tile tables fingerprint: lr %016llx cf %016llx halo %016llx nx %016llx len %016llx desc %016llx bface %016llx\n",
            (unsigned long long)fnv(fh.face_lr.data(), fh.face_lr.size() * 4), (unsigned long long)fnv(fh.cf_idx.data(), fh.cf_idx.size() * 2),
            (unsigned long long)fnv(fh.halo.data(), fh.halo.size() * 4), (unsigned long long)fnv(fh.face_nx.data(), fh.face_nx.size() * 8),
            (unsigned long long)fnv(fh.face_len.data(), fh.face_len.size() * 8), (unsigned long long)fnv(fh.tile_desc.data(), fh.tile_desc.size() * 4),
            (unsigned long long)fnv(fh.bface_e.data(), fh.bface_e.size() * 4));
  }
  // ---- multi-GPU overlap: a tile belongs to the band if one of its boundary faces is a halo face
  {
    StageTimer band_timer("  band order");
    std::vector<int32_t> band;
    fh.band_order.clear();
    for (int32_t t = 0; t < fh.n_tiles; ++t) {
      const int32_t* d = &fh.tile_desc[(size_t)t * kTileDesc];
      bool is_band = false;
      for (int32_t q = 0; q < d[5] - d[9]; ++q)
        if (ctx->bch.type[fh.bface_e[d[10] + q]] == BC_HALO) { is_band = true; break; }
      (is_band ? band : fh.band_order).push_back(t);
    }
    fh.n_interior_tiles = (int32_t)fh.band_order.size();
    // library-owned transport (one launch): the band sits in the MIDDLE of the order -- the neighbours' pushes land a few
    // microseconds after the launch starts, so the band tiles never wait there, and their slower halo faces are not the tail
    fh.comm_band0 = fh.n_interior_tiles / 2;
    fh.comm_order.assign(fh.band_order.begin(), fh.band_order.begin() + fh.comm_band0);
    fh.comm_order.insert(fh.comm_order.end(), band.begin(), band.end());
    fh.comm_order.insert(fh.comm_order.end(), fh.band_order.begin() + fh.comm_band0, fh.band_order.end());
    fh.band_order.insert(fh.band_order.end(), band.begin(), band.end());
  }
  // ---- stages of the host-buffer pipeline
  {
    StageTimer stage_timer("  pipeline stages");
    int32_t K = N >= (1 << 20) ? (int32_t)std::min<int64_t>(32, std::max<int64_t>(8, N >> 19)) : 1;   // ~0.5M cells (12 MB) per chunk
    if (ctx->opt.reserved[1] > 0) K = (int32_t)std::min<int64_t>(ctx->opt.reserved[1], std::max<int64_t>(1, N / 1024));   // tuning override
    fh.n_chunks = K;
    // chunk c nominally holds the reference rows [c csz, (c+1) csz); a component whose host address needs a shift sh < kPipeAlign
    // moves rows [c csz - sh, (c+1) csz - sh).  So after chunk c every component has landed rows < (c+1) csz - kPipeAlign (all of
    // them after the last chunk), and result chunk c can take rows from c csz - kPipeAlign on.
    const int64_t csz = K > 1 ? ((N + K - 1) / K + kPipeAlign - 1) / kPipeAlign * kPipeAlign : N;
    fh.chunk_cells = csz;
    auto stage_of_row = [&](int64_t r) { return (int32_t)std::min<int64_t>(K - 1, (r + kPipeAlign) / csz); };
    std::vector<int32_t> tstage(fh.n_tiles, 0);
#pragma omp parallel for schedule(static)
    for (int32_t t = 0; t < fh.n_tiles; ++t) {
      const int32_t* d = &fh.tile_desc[(size_t)t * kTileDesc];
      int32_t st = 0;
      for (int32_t c = d[0]; c < d[0] + d[1]; ++c) st = std::max(st, stage_of_row(fh.perm[c]));
      for (int32_t q = d[2]; q < d[2] + d[3]; ++q) st = std::max(st, stage_of_row(fh.perm[fh.halo[q]]));
      // tiles with inlet-q faces wait for the boundary-wide conveyance sum, i.e. for the last chunk; so do the tiles with halo
      // faces (multi-rank contexts: the neighbours push their cut cells once their own last chunk has landed)
      for (int32_t q = 0; q < d[5] - d[9]; ++q)
        if (ctx->bch.type[fh.bface_e[d[10] + q]] == BC_INLETQ || ctx->bch.type[fh.bface_e[d[10] + q]] == BC_HALO) st = K - 1;
      tstage[t] = st;
    }
    fh.tile_order.resize(fh.n_tiles);
    std::iota(fh.tile_order.begin(), fh.tile_order.end(), 0);
    std::stable_sort(fh.tile_order.begin(), fh.tile_order.end(), [&](int32_t x, int32_t y) { return tstage[x] < tstage[y]; });
    fh.stage_ptr.assign(K + 1, 0);
    for (int32_t t = 0; t < fh.n_tiles; ++t) fh.stage_ptr[tstage[t] + 1]++;
    for (int32_t k = 0; k < K; ++k) fh.stage_ptr[k + 1] += fh.stage_ptr[k];
    fh.chunk_done.assign(K, 0);
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < K; ++c) {
      int32_t done = 0;
      const int64_t hi = c == K - 1 ? N : std::min<int64_t>(N, (c + 1) * csz);
      for (int64_t r = std::max<int64_t>(0, c * csz - kPipeAlign); r < hi; ++r) done = std::max(done, tstage[fh.iperm[r] / T]);
      fh.chunk_done[c] = done;
    }
  }
  if (fh.halo.empty()) fh.halo.push_back(0);
  if (fh.bface_e.empty()) fh.bface_e.push_back(0);
  return HG_OK;
}

}  // namespace hg
