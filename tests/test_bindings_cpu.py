"""Static consistency of the three descriptions of the C ABI: the header (include/hydrograd_b200.h), the ctypes table the tests
drive (hydrograd.jl_b200/_lib.py) and the Julia shim a Hydrograd.jl maintainer would load (julia/HydrogradB200.jl, which cannot
be executed here: no Julia in the image).  Prototypes and struct layouts are parsed from the sources and compared field by
field, so that an argument added on one side or a struct extended in the header cannot go unnoticed on the others."""
import ctypes as C
import os
import re

import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hg = _pkg.load()
L = hg._lib

HEADER = open(os.path.join(ROOT, "include", "hydrograd_b200.h")).read()
JULIA = open(os.path.join(ROOT, "hydrograd.jl_b200", "julia", "HydrogradB200.jl")).read()
MACROS = {k: int(v) for k, v in re.findall(r"#define\s+(HG_UDE_MAX_\w+)\s+(\d+)", HEADER)}
STRUCTS = {"hg_mesh_desc": L.MeshDesc, "hg_bc_desc": L.BcDesc, "hg_fields_desc": L.FieldsDesc, "hg_options": L.Options,
           "hg_ude_desc": L.UdeDesc, "hg_named_array": L.NamedArray}
P = C.POINTER
i32p, f32p = P(C.c_int32), P(C.c_float)


def _strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def _c_type(t):
    """ctypes type of a C parameter / field type written as in the header."""
    t = " ".join(t.replace("*", " * ").split())
    const = t.startswith("const ")
    t = t[6:] if const else t
    table = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double, "float": C.c_float, "uint8_t": C.c_uint8, "uint64_t": C.c_uint64,
             "double *": L.c_f64p, "int64_t *": L.c_i64p, "int32_t *": i32p, "float *": f32p, "uint8_t *": L.c_u8p, "char *": C.c_char_p,
             "void *": C.c_void_p, "hg_ctx *": C.c_void_p, "hg_case *": C.c_void_p, "hg_ctx * *": P(C.c_void_p), "hg_case * *": P(C.c_void_p),
             "hg_json *": C.c_void_p, "hg_json * *": P(C.c_void_p),
                 "hg_plan *": C.c_void_p, "hg_plan * *": P(C.c_void_p),
             "void * *": P(C.c_void_p), "double * *": P(L.c_f64p), "hg_allreduce_fn": L.ALLREDUCE_FN}
    if t in table:
        return table[t]
    m = re.fullmatch(r"(hg_\w+) \*", t)
    if m and m.group(1) in STRUCTS:
        return P(STRUCTS[m.group(1)])
    raise KeyError(t)


def _header_prototypes():
    src = _strip_comments(HEADER)
    out = {}
    for ret, name, args in re.findall(r"HG_API\s+([\w\s\*]+?)\s*\b(hg_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret = " ".join(ret.split())
        res = None if ret == "void" else C.c_char_p if ret == "const char*" else _c_type(ret)
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.fullmatch(r"(.*?[\s\*])(\w+)", a)             # type, then the parameter name
                params.append(_c_type(m.group(1)))
        out[name] = (res, params)
    return out


def test_ctypes_prototypes_match_the_header():
    protos = _header_prototypes()
    assert set(protos) == set(L.SYMBOLS) and len(protos) >= 50
    for name, (res, params) in protos.items():
        cres, cparams = L.SYMBOLS[name]
        assert cres == res, name
        assert list(cparams) == params, name


def _header_struct_fields(name):
    src = _strip_comments(HEADER)
    body = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        first, *rest = [d.strip() for d in decl.split(",")]
        m = re.fullmatch(r"(.*?[\s\*])(\w+)((?:\[[^\]]+\])?)", first)
        base = m.group(1)
        for nm, arr in [(m.group(2), m.group(3))] + [re.fullmatch(r"(\w+)((?:\[[^\]]+\])?)", r).groups() for r in rest]:
            t = _c_type(base)
            if arr:
                expr = arr[1:-1]
                for k, v in MACROS.items():
                    expr = expr.replace(k, str(v))
                t = t * int(eval(expr, {}, {}))
            fields.append((nm, t))
    return fields


@pytest.mark.parametrize("name", sorted(STRUCTS))
def test_ctypes_structs_match_the_header(name):
    want = _header_struct_fields(name)
    got = [(n, t) for n, t in STRUCTS[name]._fields_]
    assert [n for n, _ in got] == [n for n, _ in want]
    for (n, t), (_, w) in zip(got, want):
        assert t == w or (C.sizeof(t) == C.sizeof(w) and getattr(t, "_type_", None) == getattr(w, "_type_", None)
                          and getattr(t, "_length_", None) == getattr(w, "_length_", None)), (name, n)


# ---------------------------------------------------------------------------------------------- the Julia shim
def _jl_type(t):
    t = t.strip()
    table = {"Int64": C.c_int64, "Int32": C.c_int32, "Float64": C.c_double, "Cint": C.c_int, "Cstring": C.c_char_p, "Cvoid": None,
             "Ptr{Int64}": L.c_i64p, "Ptr{Float64}": L.c_f64p, "Ptr{UInt8}": L.c_u8p, "Ptr{Cvoid}": C.c_void_p,
             "Ref{Int64}": L.c_i64p, "Ref{Int32}": P(C.c_int32), "Ptr{Int32}": P(C.c_int32), "Ref{Ptr{Cvoid}}": P(C.c_void_p),
             "Ref{MeshDesc}": P(L.MeshDesc), "Ref{BcDesc}": P(L.BcDesc), "Ref{FieldsDesc}": P(L.FieldsDesc), "Ref{Options}": P(L.Options),
             "Ref{UdeDesc}": P(L.UdeDesc)}
    if t in table:
        return table[t]
    m = re.fullmatch(r"NTuple\{(\d+),\s*(\w+)\}", t)
    if m:
        return _jl_type(m.group(2)) * int(m.group(1))
    raise KeyError(t)


JL_STRUCTS = {"MeshDesc": L.MeshDesc, "BcDesc": L.BcDesc, "FieldsDesc": L.FieldsDesc, "Options": L.Options, "UdeDesc": L.UdeDesc}


@pytest.mark.parametrize("name", sorted(JL_STRUCTS))
def test_julia_structs_mirror_the_c_structs(name):
    body = re.search(r"^struct " + name + r"\n(.*?)^end", JULIA, flags=re.S | re.M).group(1)
    body = re.sub(r"#.*", "", body)
    fields = []
    for part in re.split(r"[;\n]", body):
        part = part.strip()
        if part:
            n, t = part.split("::")
            fields.append((n.strip(), _jl_type(t)))
    want = JL_STRUCTS[name]._fields_
    assert [n for n, _ in fields] == [n for n, _ in want]
    for (n, t), (_, w) in zip(fields, want):
        assert C.sizeof(t) == C.sizeof(w) and getattr(t, "_length_", None) == getattr(w, "_length_", None), (name, n)


def _split_top(s):
    """split at commas that are not inside brackets"""
    out, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "({[" and 1 or ch in ")}]" and -1 or 0
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def test_julia_ccalls_match_the_prototypes():
    calls = re.findall(r"ccall\(\(:(hg_\w+), LIB\),\s*(\w+),\s*\(([^)]*(?:\{[^}]*\}[^)]*)*)\)\s*,", JULIA, flags=re.S)
    assert len(calls) >= 15
    seen = set()
    for name, ret, argt in calls:
        assert name in L.SYMBOLS, name
        res, params = L.SYMBOLS[name]
        got = [_jl_type(a) for a in _split_top(" ".join(argt.split()))] if argt.strip() else []
        assert _jl_type(ret) == res, name
        assert len(got) == len(params), name
        for g, w in zip(got, params):
            same = g == w or (g is not None and w is not None and C.sizeof(g) == C.sizeof(w)
                              and (issubclass(g, C._Pointer) == issubclass(w, C._Pointer) or C.c_void_p in (g, w) or C.c_char_p in (g, w)))
            assert same, (name, g, w)
        seen.add(name)
    # the shim binds the whole drop-in path
    for must in ("hg_create", "hg_destroy", "hg_last_error", "hg_rhs", "hg_rhs_vjp", "hg_custom_ode_solve", "hg_solve_tsit5_dense",
                 "hg_set_manning_function", "hg_set_ude_model", "hg_set_state", "hg_step_ab3",
                 "hg_partition_rcb", "hg_partition_extract", "hg_comm_export", "hg_comm_connect", "hg_step_euler", "hg_get_state"):
        assert must in seen, must
