// Output side of the path ("writers"): the files the reference's forward driver writes from the final
// state, host only.
//
//   numbers          Julia prints a Float64 with Base.Ryu.writeshortest (shortest digits that round-trip; positional
//                    notation when the decimal point falls in -4 < pt <= 6, otherwise d.ddde[-]x; always a ".0" on
//                    whole numbers).  JSON3.pretty (JSON3 1.14, Project.toml:82) re-reads what JSON3.write produced, and
//                    its reader hands whole-valued numbers back as Int64 -- unless the array they sit in STARTS with a
//                    non-whole number, in which case it stops looking for integers in that array.  So the reference's
//                    files carry "0" for the flat parts of zb_cell_truth (first element 0.0) but "0.0" for the zero
//                    discharges inside forward_simulation_results (first element 0.13).  The rule was identified from
//                    the reference's committed files and holds for all 26039 whole values in them (tests/test_results_cpu.py).
//   JSON             JSON3.pretty: 4 spaces per level, one element per line, keys in the order given (the reference's is
//                    Julia's Dict iteration order), "}" + newline at the end
//                    (applications/forward_simulation/process_forward_simulation_results_2D.jl:54-75,
//                    applications/sensitivity/swe_2D_sensitivity.jl:70,90).  NaN / Inf are refused, like JSON3 does.
//   VTK              export_to_vtk_2D, utilities/swe_2D_tools.jl:145-214: legacy ASCII unstructured grid, polygons.
//   derived fields   process_forward_simulation_results_2D.jl:27-48 (xi, wse, h, u = q/(h+h_small), friction of
//                    semi_discretize_swe_2D.jl:544-547 in its left-to-right order), the Manning closures of
//                    parameters/process_ManningN_2D.jl:140-213 with their diagnostic outputs (h/ks, f, Re), the dry / wet
//                    flags of fvm/discretization/process_dry_wet.jl:2-35 (only the VTK file reads them), the water volume of
//                    swe_2D_tools.jl:4-7.
// Formatting runs in parallel chunks (std::thread) into per-chunk buffers that are written in order: the reference's
// println-per-value loop is the bottleneck of its output on large meshes, 10^8 numbers should take seconds.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "../../include/hydrograd_b200.h"

namespace {

// digits and exponent of the shortest round-trip representation: x = 0.d1 d2 ... dn * 10^pt
inline int shortest_digits(double ax, char* digits, int* pt) {
  char tmp[40];
  auto r = std::to_chars(tmp, tmp + sizeof(tmp), ax, std::chars_format::scientific);   // d[.ddd]e[+-]xx
  int n = 0;
  const char* p = tmp;
  for (; p < r.ptr && *p != 'e'; ++p)
    if (*p != '.') digits[n++] = *p;
  ++p;
  int ex = 0;
  const bool neg = *p == '-';
  ++p;
  for (; p < r.ptr; ++p) ex = 10 * ex + (*p - '0');
  *pt = (neg ? -ex : ex) + 1;
  return n;
}

// Base.Ryu.writeshortest(x) with the defaults of print / show / JSON3.write
inline int format_julia(double x, char* out) {
  char* o = out;
  if (x != x) { std::memcpy(o, "NaN", 3); return 3; }
  if (std::signbit(x)) *o++ = '-';
  const double ax = std::fabs(x);
  if (std::isinf(ax)) { std::memcpy(o, "Inf", 3); return (int)(o - out) + 3; }
  if (ax == 0.0) { std::memcpy(o, "0.0", 3); return (int)(o - out) + 3; }
  char d[24];
  int pt;
  const int n = shortest_digits(ax, d, &pt);
  if (-4 < pt && pt <= 6) {
    if (pt <= 0) {
      *o++ = '0'; *o++ = '.';
      for (int i = 0; i < -pt; ++i) *o++ = '0';
      std::memcpy(o, d, n); o += n;
    } else if (pt >= n) {
      std::memcpy(o, d, n); o += n;
      for (int i = 0; i < pt - n; ++i) *o++ = '0';
      *o++ = '.'; *o++ = '0';
    } else {
      std::memcpy(o, d, pt); o += pt;
      *o++ = '.';
      std::memcpy(o, d + pt, n - pt); o += n - pt;
    }
  } else {
    *o++ = d[0]; *o++ = '.';
    if (n == 1) *o++ = '0';
    else { std::memcpy(o, d + 1, n - 1); o += n - 1; }
    *o++ = 'e';
    auto r = std::to_chars(o, o + 8, pt - 1);
    o = r.ptr;
  }
  return (int)(o - out);
}

// what JSON3.pretty leaves of a Float64 that JSON3.write wrote: whole values within Int64 come back as integers
inline int format_json3(double x, char* out) {
  if (x == std::nearbyint(x) && std::fabs(x) < 9.2e18) {
    auto r = std::to_chars(out, out + 24, (long long)x);
    return (int)(r.ptr - out);
  }
  return format_julia(x, out);
}

inline int format_any(double x, int style, char* out) { return style == HG_FMT_JSON3 ? format_json3(x, out) : format_julia(x, out); }

unsigned n_threads_for(int64_t n) {
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(hw, n / 65536));
}

// formats item i of [0, n) with `one(i, buf)` (appends to the string) in parallel chunks, writes the chunks in order
template <class F>
bool write_items(FILE* f, int64_t n, size_t reserve_per_item, F one) {
  const unsigned T = n_threads_for(n);
  const int64_t per_round = (int64_t)T * 262144;       // bounds the memory of the chunk buffers
  std::vector<std::string> buf(T);
  std::atomic<bool> failed{false};
  for (int64_t base = 0; base < n; base += per_round) {
    const int64_t m = std::min(per_round, n - base), chunk = (m + T - 1) / T;
    auto work = [&](unsigned t) {
      try {                                       // an exception must not leave a worker thread
        std::string& s = buf[t];
        s.clear();
        const int64_t a = base + (int64_t)t * chunk, b = std::min(base + m, a + chunk);
        if (a < b) s.reserve((size_t)(b - a) * reserve_per_item);
        for (int64_t i = a; i < b; ++i) one(i, s);
      } catch (...) {
        failed.store(true);
      }
    };
    if (T == 1) work(0);
    else {
      std::vector<std::thread> th;
      th.reserve(T);
      for (unsigned t = 0; t < T; ++t) {
        try {
          th.emplace_back(work, t);
        } catch (const std::system_error&) {     // no more threads: format this chunk here
          work(t);
        }
      }
      for (auto& x : th) x.join();
    }
    if (failed.load()) return false;
    for (unsigned t = 0; t < T; ++t)
      if (!buf[t].empty() && std::fwrite(buf[t].data(), 1, buf[t].size(), f) != buf[t].size()) return false;
  }
  return true;
}

void set_err(char* err, int64_t errlen, const std::string& msg) {
  if (err && errlen > 0) { std::strncpy(err, msg.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
}

std::string json_escape(const char* s) {
  std::string o;
  for (; *s; ++s) {
    const unsigned char c = (unsigned char)*s;
    if (c == '"') o += "\\\"";
    else if (c == '\\') o += "\\\\";
    else if (c == '\n') o += "\\n";
    else if (c == '\t') o += "\\t";
    else if (c == '\r') o += "\\r";
    else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
    else o += (char)c;
  }
  return o;
}
}  // namespace

// ---------------------------------------------------------------- JSON3.pretty writer
struct hg_json {
  FILE* f = nullptr;
  std::string err;
  int depth = 1;                    // inside the top-level object
  bool top_first = true;            // no key written yet
  bool have_key = false;            // a key is waiting for its value
  std::vector<bool> first;          // per open array: nothing written yet
  std::vector<int8_t> mode;         // per open array: -1 no number yet, 0 whole values as integers, 1 everything as Float64
  int style = HG_FMT_JSON3;
  // how the next number x of the innermost open array (or a top-level scalar) is written
  int number_style(double x) {
    if (style != HG_FMT_JSON3) return HG_FMT_JULIA;
    if (mode.empty()) return HG_FMT_JSON3;
    if (mode.back() < 0) mode.back() = (x == std::nearbyint(x) && std::fabs(x) < 9.2e18) ? 0 : 1;
    return mode.back() ? HG_FMT_JULIA : HG_FMT_JSON3;
  }
  bool failed = false;

  void indent(std::string& s, int d) const { s.append((size_t)d * 4, ' '); }
  // separator + indentation in front of a value (array element or the value of the pending key)
  bool lead(std::string& s) {
    if (have_key) { have_key = false; return true; }           // `"key": ` is already there
    if (first.empty()) { err = "hg_json: a value needs hg_json_key first"; return false; }
    s += first.back() ? "\n" : ",\n";
    first.back() = false;
    indent(s, depth);
    return true;
  }
  bool put(const std::string& s) {
    if (std::fwrite(s.data(), 1, s.size(), f) != s.size()) { err = "hg_json: write failed"; failed = true; return false; }
    return true;
  }
};

extern "C" {

int hg_format_f64(double x, int32_t style, char* out, int64_t cap) {
  if (!out || cap < 32 || (style != HG_FMT_JULIA && style != HG_FMT_JSON3)) return -1;
  const int n = format_any(x, style, out);
  out[n] = 0;
  return n;
}

int hg_json_open(hg_json** out, const char* path, int32_t style, char* err, int64_t errlen) {
  if (!out || !path || (style != HG_FMT_JULIA && style != HG_FMT_JSON3)) return HG_ERR_ARG;
  *out = nullptr;
  FILE* f = std::fopen(path, "wb");
  if (!f) { set_err(err, errlen, std::string("hg_json_open: cannot open ") + path); return HG_ERR_ARG; }
  hg_json* w = new hg_json();
  w->f = f;
  w->style = style;
  std::fputs("{", f);
  *out = w;
  return HG_OK;
}

const char* hg_json_error(const hg_json* w) { return w ? w->err.c_str() : "hg_json: NULL writer"; }

int hg_json_key(hg_json* w, const char* key) {
  if (!w || !key) return HG_ERR_ARG;
  if (!w->first.empty() || w->have_key) { w->err = "hg_json_key: the previous value is not finished"; return HG_ERR_STATE; }
  std::string s = w->top_first ? "\n" : ",\n";
  w->top_first = false;
  w->indent(s, 1);
  s += "\"" + json_escape(key) + "\": ";
  w->have_key = true;
  return w->put(s) ? HG_OK : HG_ERR_ARG;
}

int hg_json_begin_array(hg_json* w) {
  if (!w) return HG_ERR_ARG;
  std::string s;
  if (!w->lead(s)) return HG_ERR_STATE;
  s += "[";
  w->first.push_back(true);
  w->mode.push_back(-1);
  ++w->depth;
  return w->put(s) ? HG_OK : HG_ERR_ARG;
}

int hg_json_end_array(hg_json* w) {
  if (!w) return HG_ERR_ARG;
  if (w->first.empty()) { w->err = "hg_json_end_array: no open array"; return HG_ERR_STATE; }
  std::string s;
  --w->depth;
  if (!w->first.back()) { s += "\n"; w->indent(s, w->depth); }    // an empty array stays "[]"
  s += "]";
  w->first.pop_back();
  w->mode.pop_back();
  return w->put(s) ? HG_OK : HG_ERR_ARG;
}

int hg_json_numbers(hg_json* w, const double* x, int64_t n) {
  try {
    if (!w || n < 0 || (n > 0 && !x)) return HG_ERR_ARG;
    if (w->first.empty()) { w->err = "hg_json_numbers: no open array"; return HG_ERR_STATE; }
    for (int64_t i = 0; i < n; ++i)
      if (!std::isfinite(x[i])) {          // JSON3: "NaN not allowed to be written in JSON spec"
        char b[24];
        b[format_julia(x[i], b)] = 0;
        w->err = std::string(b) + " not allowed to be written in JSON spec (element " + std::to_string(i + 1) + ")";
        return HG_ERR_ARG;
      }
    if (n == 0) return HG_OK;
    const bool was_first = w->first.back();
    w->first.back() = false;
    const int st = w->number_style(x[0]);
    const std::string pad((size_t)w->depth * 4, ' ');
    const bool ok = write_items(w->f, n, pad.size() + 26, [&](int64_t i, std::string& s) {
      s += (i == 0 && was_first) ? "\n" : ",\n";
      s += pad;
      char b[32];
      s.append(b, (size_t)format_any(x[i], st, b));
    });
    if (!ok) { w->err = "hg_json: write failed"; w->failed = true; return HG_ERR_ARG; }
    return HG_OK;
  } catch (const std::exception& e) {
    if (w) w->err = std::string("hg_json_numbers: ") + e.what();
    return HG_ERR_ARG;
  }
}

int hg_json_number(hg_json* w, double x) {
  if (!w) return HG_ERR_ARG;
  if (!std::isfinite(x)) { w->err = "non-finite number not allowed to be written in JSON spec"; return HG_ERR_ARG; }
  std::string s;
  if (!w->lead(s)) return HG_ERR_STATE;
  char b[32];
  s.append(b, (size_t)format_any(x, w->number_style(x), b));
  return w->put(s) ? HG_OK : HG_ERR_ARG;
}

int hg_json_string(hg_json* w, const char* v) {
  if (!w || !v) return HG_ERR_ARG;
  std::string s;
  if (!w->lead(s)) return HG_ERR_STATE;
  s += "\"" + json_escape(v) + "\"";
  return w->put(s) ? HG_OK : HG_ERR_ARG;
}

/* closes the object (with the newline the reference's println(io) adds when trailing_newline != 0) and the file */
int hg_json_close(hg_json* w, int32_t trailing_newline) {
  if (!w) return HG_ERR_ARG;
  int rc = HG_OK;
  if (!w->first.empty() || w->have_key) rc = HG_ERR_STATE;
  if (w->f) {
    std::string s = w->top_first ? "}" : "\n}";
    if (trailing_newline) s += "\n";
    if (!w->put(s)) rc = HG_ERR_ARG;
    if (std::fclose(w->f) != 0) rc = HG_ERR_ARG;
  }
  if (w->failed) rc = HG_ERR_ARG;
  delete w;
  return rc;
}

// ---------------------------------------------------------------- export_to_vtk_2D (swe_2D_tools.jl:145-214)
int hg_write_vtk_2d(const char* path, int64_t n_nodes, const double* node_xyz, int64_t n_cells, int64_t ld, int32_t index_base,
                    const int64_t* cell_nodes, const int64_t* cell_nnodes, const char* field_name, const char* field_type,
                    double field_value, const hg_named_array* scalars, int64_t n_scalars, const hg_named_array* vectors,
                    int64_t n_vectors, char* err, int64_t errlen) {
  try {
    if (!path || n_nodes < 0 || n_cells < 0 || ld < 1 || (n_nodes > 0 && !node_xyz) || (n_cells > 0 && (!cell_nodes || !cell_nnodes)) ||
        n_scalars < 0 || n_vectors < 0 || (n_scalars > 0 && !scalars) || (n_vectors > 0 && !vectors) || !field_name || !field_type) {
      set_err(err, errlen, "hg_write_vtk_2d: bad argument");
      return HG_ERR_ARG;
    }
    for (int64_t c = 0; c < n_cells; ++c) {
      if (cell_nnodes[c] < 0 || cell_nnodes[c] > ld) { set_err(err, errlen, "hg_write_vtk_2d: node count of cell " + std::to_string(c + 1) + " out of range"); return HG_ERR_ARG; }
      for (int64_t j = 0; j < cell_nnodes[c]; ++j) {
        const int64_t id = cell_nodes[c + n_cells * j] - index_base;
        if (id < 0 || id >= n_nodes) { set_err(err, errlen, "hg_write_vtk_2d: node id of cell " + std::to_string(c + 1) + " out of range"); return HG_ERR_ARG; }
      }
    }
    for (int64_t k = 0; k < n_scalars + n_vectors; ++k) {
      const hg_named_array& a = k < n_scalars ? scalars[k] : vectors[k - n_scalars];
      if (!a.name || (n_cells > 0 && !a.data)) { set_err(err, errlen, "hg_write_vtk_2d: field without name or data"); return HG_ERR_ARG; }
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) { set_err(err, errlen, std::string("hg_write_vtk_2d: cannot open ") + path); return HG_ERR_ARG; }
    bool ok = true;
    std::string hd = "# vtk DataFile Version 2.0\n2D Unstructured Mesh\nASCII\nDATASET UNSTRUCTURED_GRID\n";
    if (field_name[0]) {
      hd += "FIELD FieldData 1\n";
      hd += std::string(field_name) + " 1 1 " + field_type + "\n";
      char b[32];
      // the reference passes the save index (an Int); a Float64 value would print the Julia way
      const int n = (std::strcmp(field_type, "integer") == 0 && field_value == std::nearbyint(field_value)) ? format_json3(field_value, b) : format_julia(field_value, b);
      hd.append(b, (size_t)n);
      hd += "\n";
    }
    hd += "POINTS " + std::to_string(n_nodes) + " double\n";
    ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    // nodeCoordinates is n_nodes x 3; rows as the reader keeps them (x y z per node)
    ok = ok && write_items(f, n_nodes, 80, [&](int64_t i, std::string& s) {
      char b[32];
      for (int k = 0; k < 3; ++k) {
        s.append(b, (size_t)format_julia(node_xyz[3 * i + k], b));
        s += k < 2 ? ' ' : '\n';
      }
    });
    int64_t total = 0;
    for (int64_t c = 0; c < n_cells; ++c) total += cell_nnodes[c] + 1;
    hd = "CELLS " + std::to_string(n_cells) + " " + std::to_string(total) + "\n";
    ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    ok = ok && write_items(f, n_cells, 64, [&](int64_t c, std::string& s) {
      s += std::to_string(cell_nnodes[c]);
      s += ' ';                                            // "$(length(cell)) $(join(cell .- 1, ' '))": the blank stays for an empty cell
      for (int64_t j = 0; j < cell_nnodes[c]; ++j) {
        if (j) s += ' ';
        s += std::to_string(cell_nodes[c + n_cells * j] - index_base);   // VTK ids are 0-based
      }
      s += '\n';
    });
    hd = "CELL_TYPES " + std::to_string(n_cells) + "\n";
    ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    ok = ok && write_items(f, n_cells, 2, [&](int64_t, std::string& s) { s += "7\n"; });     // VTK_POLYGON
    hd = "CELL_DATA " + std::to_string(n_cells) + "\n";
    ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    for (int64_t k = 0; k < n_scalars && ok; ++k) {
      hd = std::string("SCALARS ") + scalars[k].name + " double 1\nLOOKUP_TABLE default\n";
      ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
      const double* x = scalars[k].data;
      ok = ok && write_items(f, n_cells, 26, [&](int64_t i, std::string& s) {
        char b[32];
        s.append(b, (size_t)format_julia(x[i], b));
        s += '\n';
      });
    }
    for (int64_t k = 0; k < n_vectors && ok; ++k) {
      hd = std::string("VECTORS ") + vectors[k].name + " double\n";
      ok = ok && std::fwrite(hd.data(), 1, hd.size(), f) == hd.size();
      const double* x = vectors[k].data;                     // n_cells x 2, column-major (hcat(u, v))
      ok = ok && write_items(f, n_cells, 56, [&](int64_t i, std::string& s) {
        char b[32];
        s.append(b, (size_t)format_julia(x[i], b));
        s += ' ';
        s.append(b, (size_t)format_julia(x[i + n_cells], b));
        s += " 0.0\n";
      });
    }
    if (std::fclose(f) != 0) ok = false;
    if (!ok) { set_err(err, errlen, std::string("hg_write_vtk_2d: write failed: ") + path); return HG_ERR_ARG; }
    return HG_OK;
  } catch (const std::exception& e) {
    set_err(err, errlen, std::string("hg_write_vtk_2d: ") + e.what());
    return HG_ERR_ARG;
  }
}

// ---------------------------------------------------------------- derived fields of the forward driver
/* process_forward_simulation_results_2D.jl:27-36 + compute_friction_terms (semi_discretize_swe_2D.jl:544-547).  Any output
 * may be NULL.  The reference evaluates the friction with the UNCLAMPED h of the saved state. */
int hg_forward_truth_fields(int64_t N, const double* Q, const double* hstill, const double* wstill, const double* ManningN_cells,
                            double g, double k_n, double h_small, double* xi, double* wse, double* h, double* u, double* v,
                            double* friction_x, double* friction_y) {
  if (N < 0 || (N > 0 && (!Q || !hstill))) return HG_ERR_ARG;
  if ((wse && !wstill) || ((friction_x || friction_y) && !ManningN_cells)) return HG_ERR_ARG;
  const double eps = 2.220446049250313e-16;
  for (int64_t i = 0; i < N; ++i) {
    const double xi_i = Q[i], qx = Q[N + i], qy = Q[2 * N + i];
    const double h_i = xi_i + hstill[i];
    if (xi) xi[i] = xi_i;
    if (wse) wse[i] = xi_i + wstill[i];
    if (h) h[i] = h_i;
    if (u) u[i] = qx / (h_i + h_small);
    if (v) v[i] = qy / (h_i + h_small);
    if (friction_x || friction_y) {
      const double n = ManningN_cells[i];
      const double c = g * (n * n) / (k_n * k_n) / std::pow(h_i + h_small, 7.0 / 3.0) * std::sqrt(qx * qx + qy * qy + eps);
      if (friction_x) friction_x[i] = c * qx;
      if (friction_y) friction_y[i] = c * qy;
    }
  }
  return HG_OK;
}

/* update_ManningN_forward_simulation (process_ManningN_2D.jl:102-213): n and the diagnostics the closures return.
 * params = {n_lower, n_upper, k, h_mid} as for hg_set_manning_function; h_ks / f / Re are zero for the n(h) types. */
int hg_manning_function_cells(int32_t type, const double* params, int64_t N, const double* h, const double* umag, const double* ks,
                              double* n, double* h_ks, double* f, double* Re) {
  if (N < 0 || (N > 0 && (!h || !n))) return HG_ERR_ARG;
  if (type == HG_MANNING_H_UMAG_KS) {
    if (N > 0 && (!umag || !ks)) return HG_ERR_ARG;
  } else if (type == HG_MANNING_POWER_LAW || type == HG_MANNING_SIGMOID || type == HG_MANNING_INVERSE) {
    if (!params || !(params[2] > 0.0)) return HG_ERR_ARG;                // @assert(k > 0)
    if (type == HG_MANNING_SIGMOID && !(params[3] > 0.0)) return HG_ERR_ARG;   // @assert(h_mid > 0)
  } else {
    return HG_ERR_ARG;
  }
  for (int64_t i = 0; i < N; ++i) {
    double hk = 0.0, fi = 0.0, re = 0.0, ni;
    if (type == HG_MANNING_H_UMAG_KS) {
      re = umag[i] * h[i] / 1.0e-6;
      hk = h[i] / ks[i];
      const double x = re / 850.0, x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;     // ^9 by squaring like Julia's ^ Int
      const double alpha = 1.0 / (1.0 + x8 * x);
      const double r2 = re / (hk * 160.0);
      const double beta = 1.0 / (1.0 + r2 * r2);
      const double part1 = std::pow(re / 24.0, alpha);
      const double part2 = std::pow(1.8 * std::log10(re / 2.1), 2.0 * (1.0 - alpha) * beta);
      const double part3 = std::pow(2.0 * std::log10(11.8 * hk), 2.0 * (1.0 - alpha) * (1.0 - beta));
      fi = 1.0 / (part1 * part2 * part3);
      ni = std::sqrt(fi / 8.0) * std::pow(h[i], 1.0 / 6.0) / std::sqrt(9.81);
    } else {
      const double lo = params[0], span = params[1] - params[0], k = params[2];
      if (type == HG_MANNING_POWER_LAW) ni = lo + span * std::pow(h[i] + 2.220446049250313e-16, -k);
      else if (type == HG_MANNING_SIGMOID) ni = lo + span / (1.0 + std::exp(k * (h[i] - params[3])));
      else ni = lo + span / (1.0 + k * h[i]);
    }
    n[i] = ni;
    if (h_ks) h_ks[i] = hk;
    if (f) f[i] = fi;
    if (Re) Re[i] = re;
  }
  return HG_OK;
}

/* process_dry_wet_flags (process_dry_wet.jl:2-35).  cell_neighbors: N x ld column-major, ghost ids on boundary faces
 * (never dereferenced: the boundary test short-circuits, as in the reference). */
int hg_dry_wet_flags(int64_t N, int64_t ld, int32_t index_base, const int64_t* cell_nfaces, const int64_t* cell_faces,
                     const int64_t* cell_neighbors, const uint8_t* face_is_boundary, int64_t n_faces, const double* h,
                     const double* zb_cells, double h_small, uint8_t* b_dry_wet, uint8_t* adjacent_to_dry_land,
                     uint8_t* adjacent_to_high_dry_land) {
  if (N < 0 || ld < 1 || (N > 0 && (!cell_nfaces || !cell_faces || !cell_neighbors || !face_is_boundary || !h || !zb_cells))) return HG_ERR_ARG;
  if (!b_dry_wet || !adjacent_to_dry_land || !adjacent_to_high_dry_land) return HG_ERR_ARG;
  for (int64_t i = 0; i < N; ++i) b_dry_wet[i] = h[i] > h_small;
  for (int64_t i = 0; i < N; ++i) {
    bool adj = false, high = false;
    if (cell_nfaces[i] < 0 || cell_nfaces[i] > ld) return HG_ERR_ARG;
    for (int64_t j = 0; j < cell_nfaces[i]; ++j) {
      const int64_t fc = std::llabs(cell_faces[i + N * j]) - index_base;
      if (fc < 0 || fc >= n_faces) return HG_ERR_ARG;
      if (face_is_boundary[fc]) { adj = high = true; continue; }
      const int64_t nb = cell_neighbors[i + N * j] - index_base;
      if (nb < 0 || nb >= N) return HG_ERR_ARG;
      if (!b_dry_wet[nb]) {
        adj = true;
        if (h[i] + zb_cells[i] < zb_cells[nb]) high = true;
      }
    }
    adjacent_to_dry_land[i] = adj;
    adjacent_to_high_dry_land[i] = high;
  }
  return HG_OK;
}

/* swe_2D_calc_total_water_volume (swe_2D_tools.jl:4-7): sum(h .* cell_areas).  Julia's sum of an array is pairwise
 * (Base.mapreduce_impl, blocks of 1024, split at (first + last) >> 1), restated here; inside a block Julia's loop is @simd
 * and may be reassociated by the compiler, so the CSV agrees to rounding (1e-15 relative), not necessarily to the last digit */
static double pairwise_sum(const double* a, const double* b, int64_t lo, int64_t hi) {
  if (hi - lo <= 1024) {
    double s = a[lo] * b[lo];
    for (int64_t i = lo + 1; i < hi; ++i) s += a[i] * b[i];
    return s;
  }
  const int64_t mid = (lo + hi + 1) >> 1;
  return pairwise_sum(a, b, lo, mid) + pairwise_sum(a, b, mid, hi);
}
double hg_total_water_volume(int64_t N, const double* h, const double* cell_areas) {
  if (N <= 0 || !h || !cell_areas) return 0.0;
  return pairwise_sum(h, cell_areas, 0, N);
}
}
