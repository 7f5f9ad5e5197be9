// Linear-time SRH-2D case reader + mesh / boundary / bed builder, host only.
//
// Produces the flat tables of include/hydrograd_b200.h directly from `.srhhydro` / `.srhgeom` / `.srhmat`,
// with the same conventions as the reference's reader and mesh builder:
//   utilities/SRH_2D/SRH_2D_SRHHydro.jl 25-93, SRH_2D_SRHGeom.jl 150-219 + 251-382, SRH_2D_SRHMat.jl 30-108,
//   utilities/process_SRH_2D_input.jl 136-153 (matID), meshes/mesh_2D.jl 75-453 + 456-652,
//   fvm/boundary_conditions/bc_2D.jl 152-304 + 307-570, parameters/process_bed_2D.jl 9-66,
//   fvm/discretization/fvm_schemes_2D.jl 3-30, 89-117, 133-167, parameters/process_ManningN_2D.jl 3-47.
// The reference builder is O(N*B) / O(N*zone size) (membership scans over whole matrices and lists, mesh_2D.jl:146-152,
// 391-396, process_SRH_2D_input.jl:138-153) and cannot feed million-cell meshes; this one hashes edges once.
// Ids in the produced tables are 1-based like Julia's (index_base = 1), N x 8 tables are column-major.
// Known, deliberate deviation: ghost cells are numbered by ascending boundary-face id
// instead of Julia's Dict iteration order (SRH_2D_SRHGeom.jl:290); results do not depend on it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hydrograd_b200.h"

#include "hg_case.h"

namespace {
constexpr int LD = 8;  // gMax_Nodes_per_Element

std::vector<std::string> split(const std::string& line) {
  std::vector<std::string> out;
  std::istringstream is(line);
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}
std::string unquote(std::string s) {
  s.erase(std::remove(s.begin(), s.end(), '"'), s.end());
  return s;
}
std::string dir_of(const std::string& p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}
std::string lower(std::string s) {
  for (auto& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
// The reference's committed control files were written on a case-insensitive file system: e.g.
// examples/SWE_2D/forward_simulation/Savannah_River_ManningN_ks_h_Umag/run_control.json:5 names `Savana_SI.srhhydro`
// while the file is `savana_SI.srhhydro`.  If the exact name is absent, take the unique case-folded match of its directory.
std::string resolve_path(const std::string& path) {
  {
    std::ifstream f(path);
    if (f) return path;
  }
  const std::string dir = dir_of(path);
  const size_t k = path.find_last_of('/');
  const std::string want = lower(k == std::string::npos ? path : path.substr(k + 1));
  std::string found;
  int hits = 0;
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d))
      if (lower(e->d_name) == want) { found = e->d_name; ++hits; }
    closedir(d);
  }
  return hits == 1 ? dir + "/" + found : path;
}
inline uint64_t key(int64_t a, int64_t b) { return ((uint64_t)std::min(a, b) << 32) | (uint64_t)std::max(a, b); }

int fail(hg_case* c, const std::string& m) {
  c->err = m;
  return HG_ERR_ARG;
}

int load(hg_case* cs, const std::string& hydro_path_in) {
  const std::string hydro_path = resolve_path(hydro_path_in);
  // ---- srhhydro
  std::map<int64_t, double> mann;
  std::map<int64_t, std::string> bc;
  std::map<int64_t, double> iq, ews;
  std::string grid, matf;
  {
    std::ifstream f(hydro_path);
    if (!f) return fail(cs, "The SRHHYDRO file " + hydro_path + " does not exist");
    std::string line;
    while (std::getline(f, line)) {
      auto p = split(line);
      if (p.size() <= 1) continue;
      if (p[0] == "ManningsN" && p.size() >= 3) mann[std::stoll(p[1])] = std::stod(p[2]);
      else if (p[0] == "BC" && p.size() >= 3) { if (p[2] != "MONITORING") bc[std::stoll(p[1])] = p[2]; }
      else if (p[0] == "IQParams" && p.size() >= 3) iq[std::stoll(p[1])] = std::stod(p[2]);
      else if (p[0] == "EWSParamsC" && p.size() >= 3) ews[std::stoll(p[1])] = std::stod(p[2]);
      else if (p[0] == "Grid") grid = unquote(p[1]);
      else if (p[0] == "HydroMat") matf = unquote(p[1]);
    }
  }
  if (grid.empty() || matf.empty()) return fail(cs, "srhhydro file lacks Grid / HydroMat");
  const std::string dir = dir_of(hydro_path);

  // ---- srhgeom
  std::vector<std::vector<int64_t>> elems;
  std::vector<double> xyz;
  std::map<int64_t, std::vector<int64_t>> nstr;
  {
    std::ifstream f(resolve_path(dir + "/" + grid));
    if (!f) return fail(cs, "cannot open " + dir + "/" + grid);
    std::string line;
    int64_t cur = -1;
    while (std::getline(f, line)) {
      auto p = split(line);
      if (p.empty()) continue;
      if (p[0] == "Elem") {
        int64_t id = std::stoll(p[1]);
        if ((int64_t)elems.size() < id) elems.resize(id);
        for (size_t k = 2; k < p.size(); ++k) elems[id - 1].push_back(std::stoll(p[k]));
      } else if (p[0] == "Node") {
        int64_t id = std::stoll(p[1]);
        if ((int64_t)xyz.size() < 3 * id) xyz.resize(3 * id, 0.0);
        for (size_t k = 2; k < p.size() && k < 5; ++k) xyz[3 * (id - 1) + (k - 2)] = std::stod(p[k]);
      } else if (p[0] == "NodeString") {
        cur = std::stoll(p[1]);
        auto& v = nstr[cur];
        for (size_t k = 2; k < p.size(); ++k) v.push_back(std::stoll(p[k]));
      } else if (lower(p[0]) == "name" || lower(p[0]) == "gridunit") {
      } else if (lower(p[0]).find("srhgeom") == std::string::npos) {
        if (cur > 0) for (auto& w : p) nstr[cur].push_back(std::stoll(w));   // NodeString continuation line
      }
    }
  }
  const int64_t N = (int64_t)elems.size(), nNodes = (int64_t)xyz.size() / 3;
  if (N == 0 || nNodes == 0) return fail(cs, "empty mesh");
  for (auto& e : elems)
    if (e.size() < 3 || e.size() > LD) return fail(cs, "element with fewer than 3 or more than 8 nodes");

  // ---- edges: id at first appearance while sweeping cells, then local edges (SRH_2D_SRHGeom.jl:251-285)
  std::unordered_map<uint64_t, int64_t> edge_id;
  edge_id.reserve((size_t)N * 2);
  std::vector<int64_t> e_n1, e_n2, e_c1, e_c2;   // sorted node pair, the one or two cells
  for (int64_t c = 0; c < N; ++c) {
    const auto& nl = elems[c];
    for (size_t i = 0; i < nl.size(); ++i) {
      const int64_t a = nl[i], b = nl[(i + 1) % nl.size()];
      auto it = edge_id.find(key(a, b));
      if (it == edge_id.end()) {
        edge_id.emplace(key(a, b), (int64_t)e_n1.size() + 1);
        e_n1.push_back(std::min(a, b)); e_n2.push_back(std::max(a, b)); e_c1.push_back(c + 1); e_c2.push_back(0);
      } else {
        if (e_c2[it->second - 1] != 0) return fail(cs, "an edge is shared by more than two cells");
        e_c2[it->second - 1] = c + 1;
      }
    }
  }
  const int64_t F = (int64_t)e_n1.size();
  std::vector<int64_t> bfaces;                    // ascending face id = ghost order
  for (int64_t f = 0; f < F; ++f) if (e_c2[f] == 0) bfaces.push_back(f + 1);
  const int64_t B = (int64_t)bfaces.size();
  std::vector<int64_t> ghost_of(F + 1, 0);
  for (int64_t g = 0; g < B; ++g) ghost_of[bfaces[g]] = g + 1;

  // ---- boundaries from node strings named in the BC dict (WEIR / PRESSURE skipped), default wall for the rest
  std::map<int64_t, std::vector<int64_t>> bedges;
  std::vector<char> used(F + 1, 0);
  for (auto& kv : nstr) {
    auto it = bc.find(kv.first);
    if (it == bc.end() || it->second.find("WEIR") != std::string::npos || it->second.find("PRESSURE") != std::string::npos) continue;
    auto& lst = bedges[kv.first];
    for (size_t i = 0; i + 1 < kv.second.size(); ++i) {
      auto e = edge_id.find(key(kv.second[i], kv.second[i + 1]));
      if (e == edge_id.end()) return fail(cs, "Boundary edge in NodeString " + std::to_string(kv.first) + " not found in edge list. Mesh is wrong.");
      lst.push_back(e->second);
      used[e->second] = 1;
    }
  }
  {
    std::vector<int64_t> rest;
    for (int64_t f : bfaces) if (!used[f]) rest.push_back(f);
    if (!rest.empty()) {
      const int64_t id = (int64_t)bedges.size() + 1;   // defaultWallBoundaryID (SRH_2D_SRHGeom.jl:335)
      bedges[id] = rest;
      if (!bc.count(id)) bc[id] = "wall";              // mesh_2D.jl:172-181
    }
  }
  const int64_t nB = (int64_t)bc.size();

  // ---- cell tables + geometry (mesh_2D.jl:456-652)
  auto& nfa = cs->i64["cell_nfaces"]; nfa.assign(N, 0);
  auto& cfa = cs->i64["cell_faces"]; cfa.assign(N * LD, 0);
  auto& cnb = cs->i64["cell_neighbors"]; cnb.assign(N * LD, 0);
  auto& cnodes = cs->i64["cell_nodes"]; cnodes.assign(N * LD, 0);
  auto& cnorm = cs->f64["cell_normals"]; cnorm.assign(N * LD * 2, 0.0);
  auto& area = cs->f64["cell_areas"]; area.assign(N, 0.0);
  auto& cent = cs->f64["cell_centroids"]; cent.assign(2 * N, 0.0);
  auto& zb = cs->f64["zb_cells"]; zb.assign(N, 0.0);
  std::vector<int8_t> dirn(F + 1, 1);   // +1 when the owning cell walks the edge in sorted-node order
  for (int64_t c = 0; c < N; ++c) {
    const auto& nl = elems[c];
    const int64_t n = (int64_t)nl.size();
    nfa[c] = n;
    double a2 = 0.0, sx = 0.0, sy = 0.0, zs = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t a = nl[i], b = nl[(i + 1) % n];
      const double x1 = xyz[3 * (a - 1)], y1 = xyz[3 * (a - 1) + 1], x2 = xyz[3 * (b - 1)], y2 = xyz[3 * (b - 1) + 1];
      const double cr = x1 * y2 - x2 * y1;
      a2 += cr; sx += (x1 + x2) * cr; sy += (y1 + y2) * cr;
      zs += xyz[3 * (a - 1) + 2];
      const int64_t fid = edge_id[key(a, b)];
      cnodes[c + N * i] = a;
      cfa[c + N * i] = fid;
      const int64_t other = e_c1[fid - 1] == c + 1 ? e_c2[fid - 1] : e_c1[fid - 1];
      cnb[c + N * i] = other ? other : ghost_of[fid];
      if (!other) dirn[fid] = (a < b) ? 1 : -1;
      const double fx = x2 - x1, fy = y2 - y1, nx = fy, ny = -fx, ln = std::sqrt(nx * nx + ny * ny);
      cnorm[c + N * (i + LD * 0)] = nx / ln;
      cnorm[c + N * (i + LD * 1)] = ny / ln;
    }
    if (a2 < 0) return fail(cs, "Cell " + std::to_string(c + 1) + " is not counter-clockwise");
    const double A = std::fabs(a2) / 2;
    area[c] = A;
    cent[c] = sx / (6 * A);
    cent[N + c] = sy / (6 * A);
    zb[c] = zs / (double)n;                       // nodes_to_cells_scalar (fvm_schemes_2D.jl:108-117)
  }
  auto& flen = cs->f64["face_lengths"]; flen.assign(F, 0.0);
  std::vector<double> fnx(F), fny(F);
  auto& isb = cs->u8["face_is_boundary"]; isb.assign(F, 0);
  for (int64_t f = 0; f < F; ++f) {
    const int64_t a = e_n1[f], b = e_n2[f];
    const double nx = xyz[3 * (b - 1) + 1] - xyz[3 * (a - 1) + 1], ny = -(xyz[3 * (b - 1)] - xyz[3 * (a - 1)]);
    const double ln = std::sqrt(nx * nx + ny * ny);
    flen[f] = ln; fnx[f] = nx / ln; fny[f] = ny / ln;
    isb[f] = e_c2[f] == 0;
  }

  // ---- boundary entries in processing order inlet-q, exit-h, wall, symm (bc_2D.jl:163-241, 279-295)
  std::vector<int64_t> order[4];
  for (int64_t ib = 1; ib <= nB; ++ib) {
    auto it = bc.find(ib);
    if (it == bc.end()) return fail(cs, "Key iBoundary does not exist in srhhydro_BC: " + std::to_string(ib));
    const std::string t = lower(it->second);
    if (t == "inlet-q") { if (!iq.count(ib)) return fail(cs, "IQParams missing for boundary " + std::to_string(ib)); order[0].push_back(ib); }
    else if (t == "exit-h") { if (!ews.count(ib)) return fail(cs, "EWSParamsC missing for boundary " + std::to_string(ib)); order[1].push_back(ib); }
    else if (t == "wall") order[2].push_back(ib);
    else if (t == "symm") order[3].push_back(ib);
  }
  auto& bptr = cs->i64["bc_ptr"]; bptr.assign(1, 0);
  auto& bgh = cs->i64["bc_ghost_ids"];
  auto& bic = cs->i64["bc_internal_cells"];
  auto& blen = cs->f64["bc_lengths"];
  std::vector<double> bnx, bny;
  auto& Qin = cs->f64["inletQ_TotalQ"];
  auto& wse = cs->f64["exitH_WSE"];
  for (int t = 0; t < 4; ++t)
    for (int64_t ib : order[t]) {
      for (int64_t fid : bedges[ib]) {
        bgh.push_back(ghost_of[fid]); bic.push_back(e_c1[fid - 1]);
        bnx.push_back(dirn[fid] * fnx[fid - 1]); bny.push_back(dirn[fid] * fny[fid - 1]);
        blen.push_back(flen[fid - 1]);
      }
      bptr.push_back((int64_t)bgh.size());
      if (t == 0) Qin.push_back(iq[ib]);
      if (t == 1) wse.push_back(ews[ib]);
    }
  if ((int64_t)bgh.size() != B) return fail(cs, "boundary lists do not cover every boundary face");
  auto& bn = cs->f64["bc_normals"]; bn = bnx; bn.insert(bn.end(), bny.begin(), bny.end());

  // ---- bed data: update_bed_data (process_bed_2D.jl:46-66)
  auto& zbg = cs->f64["zb_ghost"]; zbg.assign(B, 0.0);
  for (int64_t g = 0; g < B; ++g) zbg[g] = zb[e_c1[bfaces[g] - 1] - 1];
  auto& S0 = cs->f64["S0_cells"]; S0.assign(2 * N, 0.0);
  for (int64_t c = 0; c < N; ++c) {
    double gx = 0.0, gy = 0.0;
    for (int64_t i = 0; i < nfa[c]; ++i) {
      const int64_t fid = cfa[c + N * i];
      const double zf = e_c2[fid - 1] ? (zb[e_c1[fid - 1] - 1] + zb[e_c2[fid - 1] - 1]) / 2.0 : zb[e_c1[fid - 1] - 1];
      gx = gx + cnorm[c + N * (i + LD * 0)] * zf * flen[fid - 1];
      gy = gy + cnorm[c + N * (i + LD * 1)] * zf * flen[fid - 1];
    }
    S0[c] = -1.0 * (gx / area[c]);
    S0[N + c] = -1.0 * (gy / area[c]);
  }

  // ---- materials: first zone containing the cell, 0 = default (process_SRH_2D_input.jl:136-153)
  auto& matid = cs->i64["matID_cells"]; matid.assign(N, 0);
  {
    std::ifstream f(resolve_path(dir + "/" + matf));
    if (!f) return fail(cs, "SRHMAT file " + dir + "/" + matf + " does not exist");
    std::vector<char> set(N, 0);
    std::string line;
    int64_t cur = 0;
    auto assign = [&](const std::string& w) {
      const int64_t c = std::stoll(w);
      if (c >= 1 && c <= N && !set[c - 1] && cur != 0) { matid[c - 1] = cur; set[c - 1] = 1; }
    };
    while (std::getline(f, line)) {
      auto p = split(line);
      if (p.empty() || p[0] == "SRHMAT" || p[0] == "NMaterials" || p[0] == "MatName") continue;
      if (p[0] == "Material") { cur = std::stoll(p[1]); for (size_t k = 2; k < p.size(); ++k) assign(p[k]); }
      else for (auto& w : p) assign(w);
    }
  }
  auto& nz = cs->f64["ManningN_zone"];
  for (int64_t z = 0; z < (int64_t)mann.size(); ++z) {
    if (!mann.count(z)) return fail(cs, "ManningsN ids are not 0..n-1");
    nz.push_back(mann[z]);
  }
  auto& ncell = cs->f64["ManningN_cells"]; ncell.assign(N, 0.0);
  for (int64_t c = 0; c < N; ++c) {
    if (matid[c] < 0 || matid[c] >= (int64_t)nz.size()) return fail(cs, "Material of cell " + std::to_string(c + 1) + " does not have Manning's n");
    ncell[c] = nz[matid[c]];
  }
  cs->f64["node_coords"] = xyz;
  int64_t* d = cs->dims;
  d[0] = N; d[1] = F; d[2] = B; d[3] = LD; d[4] = 1; d[5] = (int64_t)order[0].size(); d[6] = (int64_t)order[1].size();
  d[7] = (int64_t)order[2].size(); d[8] = (int64_t)order[3].size(); d[9] = (int64_t)nz.size(); d[10] = nNodes;
  return HG_OK;
}
}  // namespace

extern "C" {

int hg_case_load_srh2d(hg_case** out, const char* srhhydro_path, char* err, int64_t errlen) {
  if (!out || !srhhydro_path) return HG_ERR_ARG;
  hg_case* c = new hg_case();
  int rc;
  try {
    rc = load(c, srhhydro_path);
  } catch (const std::exception& e) {
    c->err = std::string("parse error: ") + e.what();
    rc = HG_ERR_ARG;
  }
  if (rc != HG_OK) {
    if (err && errlen > 0) { std::strncpy(err, c->err.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
    delete c;
    *out = nullptr;
    return rc;
  }
  *out = c;
  return HG_OK;
}

void hg_case_free(hg_case* c) { delete c; }

int hg_case_dims(const hg_case* c, int64_t* dims /* [16] */) {
  if (!c || !dims) return HG_ERR_ARG;
  std::memcpy(dims, c->dims, sizeof(c->dims));
  return HG_OK;
}

/* dtype: 0 = float64, 1 = int64, 2 = uint8 */
int hg_case_array(const hg_case* c, const char* name, const void** ptr, int64_t* count, int32_t* dtype) {
  if (!c || !name || !ptr || !count || !dtype) return HG_ERR_ARG;
  auto a = c->f64.find(name);
  if (a != c->f64.end()) { *ptr = a->second.data(); *count = (int64_t)a->second.size(); *dtype = 0; return HG_OK; }
  auto b = c->i64.find(name);
  if (b != c->i64.end()) { *ptr = b->second.data(); *count = (int64_t)b->second.size(); *dtype = 1; return HG_OK; }
  auto u = c->u8.find(name);
  if (u != c->u8.end()) { *ptr = u->second.data(); *count = (int64_t)u->second.size(); *dtype = 2; return HG_OK; }
  return HG_ERR_ARG;
}
}
