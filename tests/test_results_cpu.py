"""Results writers (csrc/hg_results.cpp, hydrograd.jl_b200/results.py) against the reference's OWN committed output files.

Byte-level pins: the reference's forward_simulation_solution_truth.json / sensitivity_results.json files are rewritten from
the values the fixtures hold (tests/golden/*/truth.npz, sensitivity.npz) and must reproduce the original bytes (sha256 in
tests/golden/json_digests.json, made by tests/golden/make_fixtures.py) -- 1.2e5 numbers in Julia's shortest round-trip
layout, JSON3.pretty's indentation and its integer quirk.  When /root/reference is present (the build container) every JSON
file the reference wrote with JSON3.pretty is round-tripped as well.  Derived fields are checked against the same truth data."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

import _pkg
from tests import cases

REF = "/root/reference/examples/SWE_2D"


@pytest.fixture(scope="module")
def res():
    _pkg.load()
    from hydrograd_jl_b200 import results
    return results


def _sha(path):
    raw = open(path, "rb").read()
    return hashlib.sha256(raw).hexdigest(), len(raw)


def test_julia_float_layout(res):
    """Base.Ryu.writeshortest: positional for -4 < pt <= 6, d.ddde[-]x otherwise, '.0' on whole numbers."""
    want = {0.0: "0.0", 1.0: "1.0", 0.2: "0.2", 27.0: "27.0", 123456.0: "123456.0", 999999.0: "999999.0",
            1234567.0: "1.234567e6", 1.0e6: "1.0e6", 1.3577239715379463e6: "1.3577239715379463e6", 0.0001: "0.0001",
            0.00012: "0.00012", 1.0e-5: "1.0e-5", 9.999e-5: "9.999e-5", -3.012681078935247e-16: "-3.012681078935247e-16",
            1.0e21: "1.0e21", 1.0e22: "1.0e22", 5e-324: "5.0e-324", 1.7976931348623157e308: "1.7976931348623157e308",
            0.1 + 0.2: "0.30000000000000004", 100.5: "100.5", 2.5: "2.5", -187.4: "-187.4", 9.81: "9.81",
            float("inf"): "Inf", float("-inf"): "-Inf"}
    for x, s in want.items():
        assert res.format_float(x) == s, (x, res.format_float(x), s)
    assert res.format_float(float("nan")) == "NaN" and res.format_float(-0.0) == "-0.0"
    # JSON3.pretty: whole values come back as integers
    for x, s in {0.0: "0", 27.0: "27", -3.0: "-3", 1.0e6: "1000000", 0.5: "0.5", 1.0e-5: "1.0e-5", 1.0e20: "1.0e20"}.items():
        assert res.format_float(x, "JSON3") == s
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-12, 12, 2000), rng.integers(-10**7, 10**7, 500) / 8.0])
    for x in xs:                      # shortest digits: the text parses back to the same double, and repr has no fewer digits
        s = res.format_float(float(x))
        assert float(s) == x
        assert len(s.replace("-", "").replace(".", "").split("e")[0].strip("0")) <= len(repr(float(x)).replace("-", "").replace(".", "").split("e")[0].strip("0"))


def _canon(txt):
    """(digits without leading / trailing zeros, decimal exponent of the first digit) of a positive decimal literal"""
    from decimal import Decimal
    sign, digits, exp = Decimal(txt).as_tuple()
    d = "".join(map(str, digits)).lstrip("0")
    e = exp + len(d) - 1 if d else 0
    return d.rstrip("0"), e


def test_shortest_digits_agree_with_python_repr_on_random_bit_patterns(res):
    """Both Julia (Ryu) and Python (repr) print the shortest decimal string that round-trips, the closest one when several
    exist: same digits and exponent, only the layout differs.  2e4 random bit patterns over the whole exponent range incl.
    subnormals, and neighbours of powers of ten and two."""
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 0x7FF0000000000000, 20000, dtype=np.int64)
    xs = list(bits.view(np.float64)) + [10.0 ** k for k in range(-300, 300, 7)] + [np.nextafter(10.0 ** k, 0) for k in range(-20, 22)] + \
        [2.0 ** k for k in range(-1074, 1023, 11)] + [np.nextafter(2.0 ** k, np.inf) for k in range(-60, 60)]
    for x in xs:
        x = float(x)
        if x == 0.0:
            continue
        s = res.format_float(x)
        assert float(s) == x
        assert _canon(s.replace("e", "E")) == _canon(repr(x)), (x, s, repr(x))
        pt = _canon(s.replace("e", "E"))[1] + 1
        assert ("e" in s) == (not (-4 < pt <= 6)), (x, s)


def _digests():
    return json.load(open(os.path.join(cases.GOLD, "json_digests.json")))


@pytest.mark.parametrize("name", ["oneD_bump", "oneD_uniform", "oneD_bump_nh", "savannah", "savannah_ks", "simple"])
def test_truth_json_is_reproduced_byte_for_byte(res, name, tmp_path):
    d = _digests()[name + "/truth"]
    z = cases.truth(name)
    obj = {}
    for k in d["keys"]:
        a = z[k]
        if k == "S0_faces_truth" and a.shape[1] == 8:          # padded ragged fixture (not the case for the channel meshes)
            pytest.skip("ragged S0_faces_truth fixture")
        obj[k] = a
    out = tmp_path / "truth.json"
    res.write_json_pretty(out, obj)
    assert _sha(out) == (d["sha256"], d["bytes"])


@pytest.mark.parametrize("name", ["savannah_sens", "oneD_bump_sens", "oneD_uniform_sens"])
def test_sensitivity_json_is_reproduced_byte_for_byte(res, name, tmp_path):
    d = _digests()[name + "/sensitivity"]
    z = np.load(os.path.join(cases.GOLD, name, "sensitivity.npz"))
    S = z["sensitivity_results"]
    # written through the driver-level mirror: [3N, n_params] column-major = the flat list of the file
    n_par = z["params_vector"].size
    res.save_sensitivity_results(tmp_path, sensitivity=S.reshape(n_par, -1).T, parameter_name=d["parameter_name"],
                                 params_vector=z["params_vector"])
    assert list(json.load(open(tmp_path / "sensitivity_results.json")).keys()) == d["keys"]
    assert _sha(tmp_path / "sensitivity_results.json") == (d["sha256"], d["bytes"])


@pytest.mark.parametrize("name", ["savannah_sens", "oneD_bump_sens", "oneD_uniform_sens"])
def test_per_parameter_sensitivity_files_are_reproduced_byte_for_byte(res, name, tmp_path):
    """postprocess_sensitivity_results_swe_2D: sensitivity_results_ManningN_<i>.json of the reference, from the matrix."""
    dg = _digests()
    z = np.load(os.path.join(cases.GOLD, name, "sensitivity.npz"))
    p = z["params_vector"]
    S = z["sensitivity_results"].reshape(p.size, -1).T                     # [3N, n_params]
    res.postprocess_sensitivity_results_swe_2D({"n_cells": S.shape[0] // 3}, S, p, "ManningN", tmp_path, write_vtk=False)
    n = 0
    for i in range(1, p.size + 1):
        d = dg.get(f"{name}/sensitivity_ManningN_{i}")
        if d is None:
            continue
        out = tmp_path / f"sensitivity_results_ManningN_{i}.json"
        assert list(json.load(open(out)).keys()) == d["keys"] == list(res.SENSITIVITY_PARAM_KEYS)
        assert _sha(out) == (d["sha256"], d["bytes"]), i
        n += 1
    assert n >= 2


def test_json3_integer_quirk_follows_the_first_element(res, tmp_path):
    """JSON3.pretty prints whole values as integers unless the array starts with a non-whole number: zb_cells (0.0 first) ->
    "0", the zero discharges inside forward_simulation_results (0.13 first) -> "0.0" -- both in the reference's
    sensitivity_analysis/ManningN/oneD_channel_with_bump/forward_simulation_results.json."""
    res.write_json_pretty(tmp_path / "a.json", {"forward_simulation_results": np.array([0.13, 0.13, 0.0, 0.0, 2.0, 0.25]),
                                                "zb_cells": np.array([0.0, 0.0, 0.125, 0.0, 3.0]), "n": 2.0, "empty": np.zeros(0),
                                                "nested": np.array([[0.0, 0.5], [0.5, 0.0]]), "name": "ManningN"})
    want = ('{\n    "forward_simulation_results": [\n        0.13,\n        0.13,\n        0.0,\n        0.0,\n        2.0,\n        0.25\n    ],\n'
            '    "zb_cells": [\n        0,\n        0,\n        0.125,\n        0,\n        3\n    ],\n    "n": 2,\n    "empty": [],\n'
            '    "nested": [\n        [\n            0,\n            0.5\n        ],\n        [\n            0.5,\n            0.0\n        ]\n    ],\n'
            '    "name": "ManningN"\n}\n')
    assert open(tmp_path / "a.json").read() == want
    res.write_json_pretty(tmp_path / "b.json", {"wstill": np.array([27.0, 27.0])}, style="julia", trailing_newline=False)
    assert open(tmp_path / "b.json").read() == '{\n    "wstill": [\n        27.0,\n        27.0\n    ]\n}'


def test_json_refuses_non_finite_numbers_like_json3(res, tmp_path):
    hg = _pkg.load()
    with pytest.raises(hg.HydrogradError) as e:
        res.write_json_pretty(tmp_path / "n.json", {"h": np.array([1.0, np.nan])})
    assert "NaN not allowed to be written in JSON spec" in str(e.value)
    with pytest.raises(hg.HydrogradError):
        res.write_json_pretty(tmp_path / "no_such_dir" / "n.json", {"h": np.array([1.0])})


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference examples not present (GPU box)")
def test_every_json3_file_of_the_reference_round_trips(res, tmp_path):
    """All JSON files the reference wrote with JSON3.pretty (35 files, 4.5e5 numbers; the three hand-edited inputs and the
    control files excluded): parsed, rewritten, identical bytes.  forward_simulation_initial_condition.json comes from another
    writer (every Float64 printed the Julia way, no integer quirk): style 'julia'."""
    skip = {"run_control.json", "Inversion_forward_simulation_initial_conditions.json", "inversion_parameter_initial_values.json"}
    n = 0
    for p in sorted(glob.glob(REF + "/**/*.json", recursive=True)):
        if os.path.basename(p) in skip or os.path.getsize(p) > 40e6:
            continue
        raw = open(p, "rb").read()
        obj = json.loads(raw)
        style = "julia" if os.path.basename(p) == "forward_simulation_initial_condition.json" else "JSON3"
        res.write_json_pretty(tmp_path / "rt.json", obj, style=style, trailing_newline=raw.endswith(b"}\n"))
        assert open(tmp_path / "rt.json", "rb").read() == raw, p
        n += 1
    assert n >= 30


@pytest.mark.parametrize("name", ["savannah", "oneD_bump", "oneD_uniform", "oneD_bump_nh"])
def test_truth_fields_from_the_final_state(res, name):
    """xi / wse / h / u / v / friction of process_forward_simulation_results_2D.jl:27-48 from a state rebuilt out of the truth
    file (q = u (h + h_small), exact to rounding)."""
    t = cases.truth(name)
    h, hs = t["h_truth"], 1.0e-3
    Q = np.concatenate([t["xi_truth"], t["u_truth"] * (h + hs), t["v_truth"] * (h + hs)])
    n_static = t["ManningN_cells_truth"]
    if name == "oneD_bump_nh":        # the friction of the truth file uses the STATIC n (reference detail), see results.py
        c = cases.load("oneD_bump")
        n_static = c.ManningN_cells
    f = res.forward_truth_fields(Q, t["hstill_truth"], t["wstill_truth"], n_static)
    assert np.array_equal(f["xi"], t["xi_truth"])
    assert np.abs(f["h"] - t["h_truth"]).max() <= 1e-15 * np.abs(t["h_truth"]).max()
    assert np.abs(f["wse"] - t["wse_truth"]).max() <= 1e-15 * np.abs(t["wse_truth"]).max()
    assert np.abs(f["u"] - t["u_truth"]).max() <= 4e-16 * max(np.abs(t["u_truth"]).max(), 1e-300)
    assert np.abs(f["friction_x"] - t["friction_x_truth"]).max() <= 2e-15 * np.abs(t["friction_x_truth"]).max()
    assert np.abs(f["friction_y"] - t["friction_y_truth"]).max() <= 2e-15 * max(np.abs(t["friction_x_truth"]).max(), 1e-300)


def test_manning_function_diagnostics_match_the_truth_file(res):
    """update_ManningN_forward_simulation for h_Umag_ks: n, h/ks, f and Re of Savannah_River_ManningN_ks_h_Umag's truth file."""
    from tests.test_oracle_golden import _savannah_ks_cells
    t = cases.truth("savannah_ks")
    umag = np.sqrt(t["u_truth"] ** 2 + t["v_truth"] ** 2)
    n, hk, f, Re = res.update_ManningN_forward_simulation(t["h_truth"], umag, _savannah_ks_cells(), "h_Umag_ks")
    for got, key, tol in ((n, "ManningN_cells_truth", 2e-13), (hk, "h_ks_cells_truth", 1e-15), (f, "friction_factor_cells_truth", 4e-13),
                          (Re, "Re_cells_truth", 1e-15)):
        assert np.abs(got - t[key]).max() <= tol * np.abs(t[key]).max(), key
    with pytest.raises(ValueError):
        res.update_ManningN_forward_simulation(t["h_truth"], umag, None, "cubic")
    hg = _pkg.load()
    with pytest.raises(hg.HydrogradError):
        res.update_ManningN_forward_simulation(t["h_truth"], umag, None, "sigmoid", dict(n_lower=0.02, n_upper=0.05, k=-1.0, h_mid=0.3))


def test_manning_n_of_h_closures_match_the_truth_file(res):
    t = cases.truth("oneD_bump_nh")
    rc = json.load(open(os.path.join(cases.GOLD, "oneD_bump_nh", "run_control.json")))
    fs = rc["forward_simulation_options"]
    prm = fs["forward_simulation_ManningN_function_parameters"]
    n, hk, f, Re = res.update_ManningN_forward_simulation(t["h_truth"], None, None, fs["forward_simulation_ManningN_function_type"], prm)
    assert np.abs(n - t["ManningN_cells_truth"]).max() <= 4e-16 * np.abs(n).max()
    assert not hk.any() and not f.any() and not Re.any()


def _py_flags(flat, h, zb, hs):
    N, ld, base = flat["n_cells"], flat["ld"], flat["index_base"]
    cf = np.asarray(flat["cell_faces"]).reshape(ld, N)
    nb = np.asarray(flat["cell_neighbors"]).reshape(ld, N)
    wet = h > hs
    adj, high = np.zeros(N, bool), np.zeros(N, bool)
    for i in range(N):
        for j in range(flat["cell_nfaces"][i]):
            if flat["face_is_boundary"][abs(cf[j, i]) - base]:
                adj[i] = high[i] = True
            elif not wet[nb[j, i] - base]:
                adj[i] = True
                high[i] |= (h[i] + zb[i]) < zb[nb[j, i] - base]
    return wet, adj, high


def test_vtk_file_and_dry_wet_flags(res, tmp_path):
    """export_to_vtk_2D / swe_2D_save_results_SciML on the Savannah case read by the product reader: structure of the legacy
    VTK file (counts, 0-based polygons, 18 scalars + 2 vectors in the reference's order), numbers in Julia's layout, flags
    against a plain restatement of process_dry_wet.jl, water volume, total_water_volume.csv."""
    hg = _pkg.load()
    from hydrograd_jl_b200 import srh2d
    flat = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, "savannah"), "savana_SI.srhhydro")
    z = np.load(os.path.join(cases.GOLD, "savannah", "ic.npz"))
    Q0 = srh2d.setup_initial_condition(flat, z["wse"] if "wse" in z.files else z[z.files[0]], z["wstill"] if "wstill" in z.files else 27.0)
    t = cases.truth("savannah")
    N = flat["n_cells"]
    h = t["h_truth"].copy()
    h[::7] = 5e-4                                                             # some dry cells so that the flags are not trivial
    Q = np.concatenate([h - flat["hstill"], t["u_truth"] * (h + 1e-3), t["v_truth"] * (h + 1e-3)])
    flags = res.process_dry_wet_flags(flat, h, flat["zb_cells"])
    for got, want in zip(flags, _py_flags(flat, h, flat["zb_cells"], 1e-3)):
        assert np.array_equal(got.astype(bool), want)
    assert 0 < flags[2].sum() < flags[1].sum() <= N
    vol = res.swe_2D_save_results_SciML(flat, [Q0, Q], tmp_path, 27.0, t["friction_x_truth"], t["friction_y_truth"])
    assert abs(vol[1] - (h * flat["cell_areas"]).sum()) <= 1e-13 * vol[1]
    csv = open(tmp_path / "total_water_volume.csv").read().split("\n")
    assert csv[0] == "total_water_volume" and [float(x) for x in csv[1:3]] == vol and csv[3] == ""
    lines = open(tmp_path / "forward_simulation_results_0002.vtk").read().split("\n")
    n_nodes = flat["node_coords"].size // 3
    assert lines[:8] == ["# vtk DataFile Version 2.0", "2D Unstructured Mesh", "ASCII", "DATASET UNSTRUCTURED_GRID", "FIELD FieldData 1",
                         "forward_simulation_saved_index 1 1 integer", "2", f"POINTS {n_nodes} double"]
    xyz = flat["node_coords"].reshape(-1, 3)
    assert lines[8] == " ".join(res.format_float(v) for v in xyz[0])
    o = 8 + n_nodes
    total = int((flat["cell_nfaces"] + 1).sum())
    assert lines[o] == f"CELLS {N} {total}"
    cn = flat["cell_nodes"].reshape(flat["ld"], N)
    for c in (0, 17, N - 1):
        k = flat["cell_nfaces"][c]
        assert lines[o + 1 + c] == f"{k} " + " ".join(str(v - 1) for v in cn[:k, c])
    o += 1 + N
    assert lines[o] == f"CELL_TYPES {N}" and set(lines[o + 1:o + 1 + N]) == {"7"}
    o += 1 + N
    assert lines[o] == f"CELL_DATA {N}"
    o += 1
    names = ["xi", "wstill", "hstill", "h", "hu", "hv", "ManningN", "ks", "h_ks", "friction_factor", "Re", "zb_cell", "WSE",
             "friction_x", "friction_y", "b_dry_wet", "b_Adjacent_to_dry_land", "b_Adjacent_to_high_dry_land"]
    for nm in names:
        assert lines[o] == f"SCALARS {nm} double 1" and lines[o + 1] == "LOOKUP_TABLE default"
        vals = lines[o + 2:o + 2 + N]
        if nm == "h":
            assert vals == [res.format_float(v) for v in Q[:N] + flat["hstill"]]      # h = xi + hstill, as the reference recomputes it
        if nm == "wstill":
            assert set(vals) == {"27.0"}
        if nm == "b_dry_wet":
            assert vals == ["1.0" if w else "0.0" for w in flags[0]]
        o += 2 + N
    assert lines[o] == "VECTORS U double"
    h0 = Q[0] + flat["hstill"][0]
    u0, v0 = Q[N] / (h0 + 1e-3), Q[2 * N] / (h0 + 1e-3)
    assert lines[o + 1] == f"{res.format_float(u0)} {res.format_float(v0)} 0.0"
    o += 1 + N
    assert lines[o] == "VECTORS slope double"
    S0 = flat["S0_cells"]
    assert lines[o + 1] == f"{res.format_float(S0[0])} {res.format_float(S0[N])} 0.0"
    assert lines[o + 1 + N:] == [""]
    # invalid FIELD arguments: message, nothing written (swe_2D_tools.jl:148-156); bad tables: error, not a crash
    assert res.export_to_vtk_2D(tmp_path / "x.vtk", xyz, cn.T, flat["cell_nfaces"], 3, "integer", 1, [], [], [], []) is False
    assert not os.path.exists(tmp_path / "x.vtk")
    bad = cn.T.copy()
    bad[5, 0] = n_nodes + 1
    with pytest.raises(hg.HydrogradError):
        res.export_to_vtk_2D(tmp_path / "y.vtk", xyz, bad, flat["cell_nfaces"], "", "", 0, [], [], [], [])


def test_postprocess_writes_the_truth_file_of_the_reference(res, tmp_path):
    """postprocess_forward_simulation_results_swe_2D on the Savannah case: from the final state to the truth file -- same keys
    in the same order as the reference's file, arrays equal to its values to rounding (the state is rebuilt from u, v)."""
    from hydrograd_jl_b200 import srh2d
    flat = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, "savannah"), "savana_SI.srhhydro")
    t = cases.truth("savannah")
    flat["hstill"] = t["hstill_truth"]
    h = t["h_truth"]
    Q = np.concatenate([t["xi_truth"], t["u_truth"] * (h + 1e-3), t["v_truth"] * (h + 1e-3)])
    out = res.postprocess_forward_simulation_results_swe_2D(flat, Q, tmp_path, t["wstill_truth"], t["ManningN_zone_values_truth"],
                                                            t["inlet_discharges_truth"])
    d = json.load(open(tmp_path / "forward_simulation_solution_truth.json"))
    assert list(d.keys()) == _digests()["savannah/truth"]["keys"] == list(res.TRUTH_KEYS)
    for k in res.TRUTH_KEYS:
        a, b = np.asarray(d[k], dtype=np.float64), t[k]
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 2e-15 * max(np.abs(b).max(), 1e-300), k
        assert np.array_equal(a, out[k])
    for k in ("zb_cell_truth", "S0_cells_truth", "hstill_truth", "wstill_truth", "xi_truth", "ManningN_cells_truth"):
        assert np.array_equal(np.asarray(d[k]), t[k]), k               # geometry from the product reader: to the bit


def _case_dir(tmp_path, name, res):
    """A case directory like the reference's examples/SWE_2D/forward_simulation/<case>: the committed inputs + run_control.json
    (+ the initial-condition file, rewritten from the fixture with the 'julia' number style of the original)."""
    import shutil
    src = os.path.join(cases.GOLD, name)
    dst = tmp_path / name
    shutil.copytree(src, dst)
    ic = os.path.join(src, "ic.npz")
    if os.path.exists(ic):
        z = np.load(ic)
        res.write_json_pretty(dst / "forward_simulation_initial_condition.json", {k: z[k] for k in z.files}, style="julia")
    return str(dst)


def test_forward_driver_reads_the_case_and_refuses_to_run_without_a_gpu(res, tmp_path):
    """run_forward_case up to the device: run_control.json, SRH-2D files and the initial condition are read by the product
    code; without a CUDA device the driver fails loudly (no CPU fallback).  The full run is tests/test_gpu_zzz_forward_driver.py."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hg = _pkg.load()
    from hydrograd_jl_b200 import forward
    for name in ("savannah", "oneD_bump"):
        with pytest.raises(hg.HydrogradError) as e:
            forward.run_forward_case(_case_dir(tmp_path, name, res))
        assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    d = _case_dir(tmp_path / "x", "oneD_bump", res)
    rc = json.load(open(os.path.join(d, "run_control.json")))
    rc["control_variables"]["bPerform_Forward_Simulation"] = False
    json.dump(rc, open(os.path.join(d, "run_control.json"), "w"))
    with pytest.raises(ValueError):
        forward.run_forward_case(d)


def test_sensitivity_driver_reads_the_case_and_refuses_to_run_without_a_gpu(res, tmp_path):
    import shutil
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hg = _pkg.load()
    from hydrograd_jl_b200 import sensitivity
    d = _case_dir(tmp_path, "savannah", res)
    shutil.copyfile(os.path.join(cases.GOLD, "savannah_sens", "run_control.json"), os.path.join(d, "run_control.json"))
    with pytest.raises(hg.HydrogradError) as e:
        sensitivity.run_sensitivity_case(d)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    rc = json.load(open(os.path.join(d, "run_control.json")))
    flat = {"n_cells": 5}
    opts = rc["sensitivity_analysis_options"]
    assert np.array_equal(sensitivity._params_vector(opts, "ManningN", flat, d), [0.02, 0.04, 0.05, 0.03, 0.045, 0.05])
    assert np.array_equal(sensitivity._params_vector(opts, "zb", flat, d), np.zeros(5))


def test_written_json_parses_back_to_the_same_values(res, tmp_path):
    """Structure fuzz: nested / ragged / empty arrays, scalars, strings with quotes, backslashes, control characters and UTF-8 --
    whatever is written must be valid JSON that parses back to the same values (numbers exactly)."""
    rng = np.random.default_rng(3)

    def rand_array(depth):
        if depth == 0 or rng.random() < 0.4:
            n = int(rng.integers(0, 6))
            return [float(x) for x in np.round(rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8), int(rng.integers(0, 12)))]
        return [rand_array(depth - 1) for _ in range(int(rng.integers(0, 4)))]

    for trial in range(20):
        obj = {f"key {trial}.{k}": rand_array(3) for k in range(4)}
        obj["name"] = 'zone "A"\\B\n\ttab \x01 é 水'
        obj["scalar"] = float(rng.standard_normal())
        obj["whole"] = 7.0
        for style in ("JSON3", "julia"):
            res.write_json_pretty(tmp_path / "f.json", obj, style=style)
            back = json.load(open(tmp_path / "f.json", encoding="utf-8"))
            assert list(back.keys()) == list(obj.keys())

            def same(a, b):
                if isinstance(a, list):
                    return isinstance(b, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
                return a == b
            for k in obj:
                assert same(obj[k], back[k]), (k, obj[k], back[k])


def test_custom_solver_results_writer(res, tmp_path):
    """swe_2D_save_results_custom: five scalars + U per saved column, the first state block labelled "h" as in the reference."""
    from hydrograd_jl_b200 import srh2d
    flat = srh2d.process_SRH_2D_input(os.path.join(cases.GOLD, "simple"), "simple.srhhydro")
    Q0 = srh2d.setup_initial_condition(flat, 1.0, 0.5, 0.1, 0.0)
    N = flat["n_cells"]
    sol = np.stack([Q0, Q0 * 1.01], axis=1)
    vol = res.swe_2D_save_results_custom(flat, sol, tmp_path)
    assert vol[0] == pytest.approx((Q0[:N] * flat["cell_areas"]).sum(), rel=1e-14)
    lines = open(tmp_path / "forward_simulation_results_0002.vtk").read().split("\n")
    assert lines[6] == "2"
    heads = [l for l in lines if l.startswith(("SCALARS", "VECTORS"))]
    assert heads == ["SCALARS h double 1", "SCALARS hu double 1", "SCALARS hv double 1", "SCALARS zb_cell double 1", "SCALARS WSE double 1",
                     "VECTORS U double"]
    i = lines.index("SCALARS h double 1")
    assert lines[i + 2] == res.format_float(sol[0, 1])
    j = lines.index("VECTORS U double")
    assert lines[j + 1] == f"{res.format_float(sol[N, 1] / sol[0, 1])} {res.format_float(sol[2 * N, 1] / sol[0, 1])} 0.0"
    with pytest.raises(ValueError):
        res.swe_2D_save_results_custom(flat, sol[:5], tmp_path)
