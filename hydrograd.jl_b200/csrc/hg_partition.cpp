// Domain decomposition for multi-GPU runs, host only (north_star: recursive coordinate bisection, one-layer
// ghost-cell halo).  The reference is a single serial process, so nothing here restates reference code; the rules are the
// ones DESIGN.md section 6 states:
//   hg_partition_rcb       recursive coordinate bisection of the cell centroids into P parts; groups of cells that must
//                          stay together (the cells of an inlet-q boundary: its conveyance sum runs over all of its faces,
//                          bc_2D.jl:665-691) are moved as a whole to the rank that owns most of them
//   hg_partition_extract   the rank-local mesh in the flat layout of include/hydrograd_b200.h: owned cells, the physical
//                          boundaries restricted to them, and one "halo boundary" per neighbouring rank whose entries are
//                          the cut faces in an order both ranks agree on (sorted by the global ids of the two cells) with
//                          the flip flag of the canonical orientation (L = smaller global id) -- the tables hg_create takes,
//                          and what a peer needs for hg_comm_connect (neighbour ranks, entries per neighbour)
// A non-Python host (Julia via ccall) partitions, creates one context per GPU and connects them with these two calls
// plus hg_comm_*; hydrograd.jl_b200/parallel.py calls the same code.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "../../include/hydrograd_b200.h"
#include "hg_case.h"

namespace {

void rcb(const double* cx, const double* cy, std::vector<int64_t>& idx, int64_t lo, int64_t hi, int32_t p0, int32_t p, int32_t* part) {
  if (p == 1) {
    for (int64_t i = lo; i < hi; ++i) part[idx[i]] = p0;
    return;
  }
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
  for (int64_t i = lo; i < hi; ++i) {
    x0 = std::min(x0, cx[idx[i]]); x1 = std::max(x1, cx[idx[i]]);
    y0 = std::min(y0, cy[idx[i]]); y1 = std::max(y1, cy[idx[i]]);
  }
  const double* key = (x1 - x0) >= (y1 - y0) ? cx : cy;
  const int32_t pl = p / 2;
  const int64_t n = hi - lo;
  const int64_t k = (int64_t)std::nearbyint((double)n * pl / p);   // round-half-even, like Python's round()
  // a stable sort by the coordinate, ties by position: the split is the one numpy's stable argsort gives
  std::stable_sort(idx.begin() + lo, idx.begin() + hi, [key](int64_t a, int64_t b) { return key[a] < key[b]; });
  rcb(cx, cy, idx, lo, lo + k, p0, pl, part);
  rcb(cx, cy, idx, lo + k, hi, p0 + pl, p - pl, part);
}

int fail(char* err, int64_t errlen, const std::string& m) {
  if (err && errlen > 0) { std::strncpy(err, m.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
  return HG_ERR_ARG;
}

}  // namespace

extern "C" {

int hg_partition_rcb(int64_t N, const double* cx, const double* cy, int32_t P, int64_t n_groups, const int64_t* group_ptr,
                     const int64_t* group_cells, int32_t* part) {
  if (N <= 0 || !cx || !cy || P < 1 || !part || n_groups < 0 || (n_groups > 0 && (!group_ptr || !group_cells))) return HG_ERR_ARG;
  std::vector<int64_t> idx(N);
  std::iota(idx.begin(), idx.end(), 0);
  rcb(cx, cy, idx, 0, N, 0, P, part);
  for (int64_t g = 0; g < n_groups; ++g) {
    std::vector<int64_t> cnt(P, 0);
    for (int64_t q = group_ptr[g]; q < group_ptr[g + 1]; ++q) {
      if (group_cells[q] < 0 || group_cells[q] >= N) return HG_ERR_ARG;
      cnt[part[group_cells[q]]]++;
    }
    const int32_t best = (int32_t)(std::max_element(cnt.begin(), cnt.end()) - cnt.begin());   // ties: the lowest rank
    for (int64_t q = group_ptr[g]; q < group_ptr[g + 1]; ++q) part[group_cells[q]] = best;
  }
  return HG_OK;
}

int hg_partition_extract(hg_case** out, const hg_mesh_desc* m, const hg_bc_desc* b, const hg_fields_desc* f, const int32_t* part,
                         int32_t rank, const int64_t* gid, char* err, int64_t errlen) {
  if (!out) return HG_ERR_ARG;
  *out = nullptr;
  if (!m || !b || !f || !part) return fail(err, errlen, "hg_partition_extract: null argument");
  const int64_t N = m->n_cells, F = m->n_faces, B = m->n_ghost, ld = m->ld, base = m->index_base;
  if (b->n_halo != 0) return fail(err, errlen, "hg_partition_extract: the input mesh must be the global one (no halo boundaries)");
  auto G = [&](int64_t c) { return gid ? gid[c] : c; };
  std::vector<int64_t> own;
  std::vector<int64_t> g2l(N, -1);
  for (int64_t c = 0; c < N; ++c)
    if (part[c] == rank) { g2l[c] = (int64_t)own.size(); own.push_back(c); }
  const int64_t n = (int64_t)own.size();
  if (n == 0) return fail(err, errlen, "hg_partition_extract: rank " + std::to_string(rank) + " owns no cell");
  auto face_of = [&](int64_t c, int64_t j) { const int64_t v = m->cell_faces[c + N * j]; return (v < 0 ? -v : v) - base; };
  auto nb_of = [&](int64_t c, int64_t j) { return m->cell_neighbors[c + N * j] - base; };

  hg_case* cs = new hg_case();
  auto& I = cs->i64; auto& D = cs->f64; auto& U = cs->u8;

  // ---- physical boundaries restricted to the owned cells (processing order kept; empty ones dropped)
  const int64_t counts_in[4] = {b->n_inletq, b->n_exith, b->n_wall, b->n_symm};
  int64_t new_counts[4] = {0, 0, 0, 0};
  std::vector<int64_t> ptr(1, 0), e_gh, ic_all, keepQ, keepW;
  std::vector<double> nx_all, ny_all, len_all;
  int64_t kb = 0;
  for (int t = 0; t < 4; ++t)
    for (int64_t kk = 0; kk < counts_in[t]; ++kk, ++kb) {
      int64_t inside = 0;
      const int64_t e0 = b->bc_ptr[kb], e1 = b->bc_ptr[kb + 1];
      for (int64_t e = e0; e < e1; ++e) inside += g2l[b->internal_cells[e] - base] >= 0;
      if (inside == 0) continue;
      if (t == 0 && inside != e1 - e0) {
        delete cs;
        return fail(err, errlen, "an inlet-q boundary is split across ranks: its conveyance sum runs over all of its faces -- pass the "
                                 "inlet's cells as a group to hg_partition_rcb");
      }
      new_counts[t]++;
      for (int64_t e = e0; e < e1; ++e) {
        const int64_t lc = g2l[b->internal_cells[e] - base];
        if (lc < 0) continue;
        e_gh.push_back(b->ghost_ids[e] - base); ic_all.push_back(lc);
        nx_all.push_back(b->outward_normals[e]); ny_all.push_back(b->outward_normals[B + e]);
        len_all.push_back(b->face_lengths ? b->face_lengths[e] : 0.0);
      }
      ptr.push_back((int64_t)e_gh.size());
      if (t == 0) keepQ.push_back(kk);
      if (t == 1) keepW.push_back(kk);
    }
  const int64_t n_phys = (int64_t)e_gh.size();

  // ---- cut faces, grouped by neighbouring rank, each group sorted by (min gid, max gid) of the two cells
  struct Cut { int64_t lc, j, r, lo, hi; int32_t q; };
  std::vector<Cut> cuts;
  for (int64_t lc = 0; lc < n; ++lc) {
    const int64_t c = own[lc];
    for (int64_t j = 0; j < m->cell_nfaces[c]; ++j) {
      const int64_t fid = face_of(c, j);
      if (fid < 0 || fid >= F) { delete cs; return fail(err, errlen, "hg_partition_extract: face id out of range"); }
      if (m->face_is_boundary[fid]) continue;
      const int64_t r = nb_of(c, j);
      if (r < 0 || r >= N) { delete cs; return fail(err, errlen, "hg_partition_extract: neighbour id out of range"); }
      if (part[r] != rank) cuts.push_back({lc, j, r, std::min(G(c), G(r)), std::max(G(c), G(r)), part[r]});
    }
  }
  std::stable_sort(cuts.begin(), cuts.end(), [](const Cut& x, const Cut& y) {
    if (x.q != y.q) return x.q < y.q;
    if (x.lo != y.lo) return x.lo < y.lo;
    return x.hi < y.hi;
  });
  std::vector<int64_t> neighbors, hcounts;
  for (const Cut& c : cuts) {
    if (neighbors.empty() || neighbors.back() != c.q) { neighbors.push_back(c.q); hcounts.push_back(0); }
    hcounts.back()++;
  }
  for (size_t k = 0; k < neighbors.size(); ++k) ptr.push_back(ptr.back() + hcounts[k]);
  const int64_t Bl = n_phys + (int64_t)cuts.size();

  // ---- local tables
  std::vector<int64_t> old2new(std::max<int64_t>(B, 1), -1);
  for (int64_t q = 0; q < n_phys; ++q) old2new[e_gh[q]] = q;
  auto& l_nf = I["cell_nfaces"]; auto& l_faces = I["cell_faces"]; auto& l_neigh = I["cell_neighbors"];
  auto& l_norm = D["cell_normals"];
  l_nf.resize(n); l_faces.assign(n * ld, 0); l_neigh.assign(n * ld, 0); l_norm.assign(n * ld * 2, 0.0);
  std::vector<int64_t> used;
  used.reserve(n * 3);
  for (int64_t lc = 0; lc < n; ++lc) {
    const int64_t c = own[lc];
    l_nf[lc] = m->cell_nfaces[c];
    for (int64_t j = 0; j < ld; ++j) {
      l_norm[lc + n * (j + ld * 0)] = m->cell_normals[c + N * (j + ld * 0)];
      l_norm[lc + n * (j + ld * 1)] = m->cell_normals[c + N * (j + ld * 1)];
    }
    for (int64_t j = 0; j < m->cell_nfaces[c]; ++j) {
      const int64_t fid = face_of(c, j), r = nb_of(c, j);
      used.push_back(fid);
      if (m->face_is_boundary[fid]) {
        if (r < 0 || r >= B || old2new[r] < 0) { delete cs; return fail(err, errlen, "hg_partition_extract: a boundary face of an owned cell has no boundary entry"); }
        l_neigh[lc + n * j] = old2new[r];
      } else if (part[r] == rank) {
        l_neigh[lc + n * j] = g2l[r];
      }
    }
  }
  for (size_t q = 0; q < cuts.size(); ++q) l_neigh[cuts[q].lc + n * cuts[q].j] = n_phys + (int64_t)q;   // the cut slot itself
  std::sort(used.begin(), used.end());
  used.erase(std::unique(used.begin(), used.end()), used.end());
  const int64_t Fl = (int64_t)used.size();
  auto f2l = [&](int64_t fid) { return (int64_t)(std::lower_bound(used.begin(), used.end(), fid) - used.begin()); };
  auto& l_isb = U["face_is_boundary"]; auto& l_flen = D["face_lengths"];
  l_isb.assign(Fl, 0); l_flen.resize(Fl);
  for (int64_t q = 0; q < Fl; ++q) { l_flen[q] = m->face_lengths[used[q]]; l_isb[q] = m->face_is_boundary[used[q]] ? 1 : 0; }
  for (int64_t lc = 0; lc < n; ++lc)
    for (int64_t j = 0; j < l_nf[lc]; ++j) l_faces[lc + n * j] = f2l(face_of(own[lc], j));
  for (const Cut& c : cuts) l_isb[l_faces[c.lc + n * c.j]] = 1;

  // ---- boundary entry arrays: physical entries, then the halo entries
  auto& bc_ic = I["bc_internal_cells"]; auto& bc_n = D["bc_normals"]; auto& bc_len = D["bc_lengths"];
  auto& flip = U["halo_flip"]; auto& harea = D["halo_area"];
  bc_ic = ic_all; bc_len = len_all;
  flip.assign(n_phys, 0); harea.assign(n_phys, 1.0);
  auto& halo_remote = I["halo_remote"]; auto& halo_cells = I["halo_cells"];
  for (const Cut& c : cuts) {
    const int64_t cc = own[c.lc];
    bc_ic.push_back(c.lc);
    nx_all.push_back(m->cell_normals[cc + N * (c.j + ld * 0)]); ny_all.push_back(m->cell_normals[cc + N * (c.j + ld * 1)]);
    bc_len.push_back(m->face_lengths[face_of(cc, c.j)]);
    flip.push_back(G(c.r) < G(cc) ? 1 : 0); harea.push_back(m->cell_areas[c.r]);
    halo_remote.push_back(c.r); halo_cells.push_back(c.lc);
  }
  bc_n = nx_all; bc_n.insert(bc_n.end(), ny_all.begin(), ny_all.end());
  I["bc_ptr"] = ptr;
  auto& bc_gh = I["bc_ghost_ids"];
  bc_gh.resize(Bl);
  std::iota(bc_gh.begin(), bc_gh.end(), 0);

  // ---- fields
  auto take = [&](const double* src, std::vector<double>& dst) { dst.resize(n); for (int64_t q = 0; q < n; ++q) dst[q] = src[own[q]]; };
  take(m->cell_areas, D["cell_areas"]);
  if (m->cell_centroids) {
    auto& cc = D["cell_centroids"];
    cc.resize(2 * n);
    for (int64_t q = 0; q < n; ++q) { cc[q] = m->cell_centroids[own[q]]; cc[n + q] = m->cell_centroids[N + own[q]]; }
  }
  take(f->hstill, D["hstill"]); take(f->zb_cells, D["zb_cells"]); take(f->ManningN_cells, D["ManningN_cells"]);
  auto& s0 = D["S0_cells"];
  s0.resize(2 * n);
  for (int64_t q = 0; q < n; ++q) { s0[q] = f->S0_cells[own[q]]; s0[n + q] = f->S0_cells[N + own[q]]; }
  auto& hg_ = D["hstill_ghost"]; auto& zg = D["zb_ghost"];
  for (int64_t q = 0; q < n_phys; ++q) { hg_.push_back(f->hstill_ghost[e_gh[q]]); zg.push_back(f->zb_ghost[e_gh[q]]); }
  for (const Cut& c : cuts) { hg_.push_back(f->hstill[c.r]); zg.push_back(f->zb_cells[c.r]); }
  if (f->matID_cells) { auto& mid = I["matID_cells"]; mid.resize(n); for (int64_t q = 0; q < n; ++q) mid[q] = f->matID_cells[own[q]]; }
  auto& qin = D["inletQ_TotalQ"]; auto& wse = D["exitH_WSE"];
  for (int64_t k : keepQ) qin.push_back(f->inletQ_TotalQ[k]);
  for (int64_t k : keepW) wse.push_back(f->exitH_WSE[k]);
  I["own"] = own; I["neighbors"] = neighbors; I["counts"] = hcounts;
  // dims = {N, F, B, ld, index_base, n_inletq, n_exith, n_wall, n_symm, n_mat, n_halo, n_halo_entries}
  const int64_t dims[12] = {n, Fl, Bl, ld, 0, new_counts[0], new_counts[1], new_counts[2], new_counts[3], f->n_mat,
                            (int64_t)neighbors.size(), (int64_t)cuts.size()};
  std::memcpy(cs->dims, dims, sizeof(dims));
  *out = cs;
  return HG_OK;
}

}  // extern "C"
