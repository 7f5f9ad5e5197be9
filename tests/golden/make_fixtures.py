"""Builds tests/golden/ from the reference's own committed example data (DATA only, no code).

Run once in the build container (needs /root/reference):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference, so everything the tests need is copied / condensed here:

  <case>/*.srhgeom|.srhhydro|.srhmat     mesh inputs, copied verbatim
  <case>/truth.npz                       arrays of forward_simulation_solution_truth.json (float64, exact)
  <case>/ic.npz                          forward_simulation_initial_condition.json (Savannah)
  oneD_*_sens/trajectory.npz             columns 0, 50, 100 and eight early columns of the 3N x 101 trajectory that
                                         swe_2D_sensitivity.jl:60-70 saves (the VALUES of a ForwardDiff.Dual solve)
  savannah_sens/sensitivity.npz          sensitivity_results.json of sensitivity_analysis/ManningN/Savana_River (6 zones x 3N)
  oneD_*_sens/sensitivity.npz            sensitivity_results.json: d Q(T) / d ManningN zones (ForwardDiff.jacobian of the solve)

JSON numbers are written by the reference with 17 significant digits, so the float64 values round-trip exactly.
"""
import json
import os
import shutil

import numpy as np

REF = "/root/reference/examples/SWE_2D"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "savannah": ("forward_simulation/Savannah_River", "savana_SI"),
    "oneD_bump": ("forward_simulation/oneD_channel_with_bump", "oneD_channel_with_bump_refined"),
    "oneD_uniform": ("forward_simulation/oneD_channel_uniform_flow", "oneD_channel_uniform_flow_refined"),
    "simple": ("inversion/bathymetry_inversion/simple", "simple"),
    "oneD_bump_sens": ("sensitivity_analysis/ManningN/oneD_channel_with_bump", "oneD_channel_with_bump_refined"),
    "oneD_uniform_sens": ("sensitivity_analysis/ManningN/oneD_channel_uniform_flow", "oneD_channel_uniform_flow_refined"),
}


# forward simulations with a state-dependent Manning's n (semi_discretize_swe_2D.jl:140-149): only the truth arrays and the
# closure parameters are needed (savannah_ks runs on the savannah mesh files, byte-identical in the reference)
VARIABLE_N = {
    "savannah_ks": "forward_simulation/Savannah_River_ManningN_ks_h_Umag",
    "oneD_bump_nh": "forward_simulation/oneD_channel_with_bump_ManningN_h",
}


def variable_n():
    for name, rel in VARIABLE_N.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(HERE, name)
        os.makedirs(dst, exist_ok=True)
        shutil.copyfile(os.path.join(src, "run_control.json"), os.path.join(dst, "run_control.json"))
        os.chmod(os.path.join(dst, "run_control.json"), 0o644)
        d = json.load(open(os.path.join(src, "forward_simulation_solution_truth.json")))
        np.savez_compressed(os.path.join(dst, "truth.npz"), **{k: np.array(v, dtype=np.float64) for k, v in d.items()})


def savannah_sensitivity():
    """sensitivity_results.json of the Savannah sensitivity run (same mesh / IC files as the forward case, byte-identical)."""
    src = os.path.join(REF, "sensitivity_analysis/ManningN/Savana_River")
    dst = os.path.join(HERE, "savannah_sens")
    os.makedirs(dst, exist_ok=True)
    for f in ("savana_SI.srhgeom", "savana_SI.srhhydro", "savana_SI.srhmat", "forward_simulation_initial_condition.json"):
        a = open(os.path.join(src, f), "rb").read()
        b = open(os.path.join(REF, "forward_simulation/Savannah_River", f), "rb").read()
        assert a == b, f
    shutil.copyfile(os.path.join(src, "run_control.json"), os.path.join(dst, "run_control.json"))
    os.chmod(os.path.join(dst, "run_control.json"), 0o644)
    d = json.load(open(os.path.join(src, "sensitivity_results.json")))
    np.savez_compressed(os.path.join(dst, "sensitivity.npz"),
                        **{k: np.array(v, dtype=np.float64) for k, v in d.items() if not isinstance(v, str)})


def json_digests():
    """sha256 / length of the reference's own JSON outputs whose VALUES the fixtures hold completely: the byte-level pins of
    the results writer (tests/test_results_cpu.py rewrites them from truth.npz / sensitivity.npz and compares digests)."""
    import hashlib
    files = {name: os.path.join(REF, rel, "forward_simulation_solution_truth.json") for name, (rel, _) in CASES.items()}
    files.update({name: os.path.join(REF, rel, "forward_simulation_solution_truth.json") for name, rel in VARIABLE_N.items()})
    out = {}
    for name, path in sorted(files.items()):
        if os.path.exists(path) and os.path.exists(os.path.join(HERE, name, "truth.npz")):
            raw = open(path, "rb").read()
            out[name + "/truth"] = dict(sha256=hashlib.sha256(raw).hexdigest(), bytes=len(raw), keys=list(json.loads(raw).keys()))
    for name, rel in (("savannah_sens", "sensitivity_analysis/ManningN/Savana_River"),
                      ("oneD_bump_sens", CASES["oneD_bump_sens"][0]), ("oneD_uniform_sens", CASES["oneD_uniform_sens"][0])):
        raw = open(os.path.join(REF, rel, "sensitivity_results.json"), "rb").read()
        d = json.loads(raw)
        out[name + "/sensitivity"] = dict(sha256=hashlib.sha256(raw).hexdigest(), bytes=len(raw), keys=list(d.keys()),
                                          parameter_name=d["parameter_name"])
        # per-parameter files of postprocess_sensitivity_results_swe_2D (derived from the same matrix)
        k = 1
        while os.path.exists(os.path.join(REF, rel, f"sensitivity_results_ManningN_{k}.json")):
            raw = open(os.path.join(REF, rel, f"sensitivity_results_ManningN_{k}.json"), "rb").read()
            out[f"{name}/sensitivity_ManningN_{k}"] = dict(sha256=hashlib.sha256(raw).hexdigest(), bytes=len(raw),
                                                         keys=list(json.loads(raw).keys()))
            k += 1
    with open(os.path.join(HERE, "json_digests.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    variable_n()
    savannah_sensitivity()
    json_digests()
    for name, (rel, stem) in CASES.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(HERE, name)
        os.makedirs(dst, exist_ok=True)
        for ext in (".srhgeom", ".srhhydro", ".srhmat"):
            shutil.copyfile(os.path.join(src, stem + ext), os.path.join(dst, stem + ext))
            os.chmod(os.path.join(dst, stem + ext), 0o644)
        shutil.copyfile(os.path.join(src, "run_control.json"), os.path.join(dst, "run_control.json"))
        os.chmod(os.path.join(dst, "run_control.json"), 0o644)
        t = os.path.join(src, "forward_simulation_solution_truth.json")
        if os.path.exists(t):
            d = json.load(open(t))
            arrs = {}
            for k, v in d.items():
                try:
                    arrs[k] = np.array(v, dtype=np.float64)
                except (ValueError, TypeError):
                    # S0_faces_truth is ragged [N][nF][2]; pad to [N, 8, 2]
                    a = np.zeros((len(v), 8, 2))
                    for i, row in enumerate(v):
                        for j, p in enumerate(row):
                            a[i, j] = p
                    arrs[k] = a
            np.savez_compressed(os.path.join(dst, "truth.npz"), **arrs)
        ic = os.path.join(src, "forward_simulation_initial_condition.json")
        if os.path.exists(ic):
            d = json.load(open(ic))
            np.savez_compressed(os.path.join(dst, "ic.npz"), **{k: np.array(v, dtype=np.float64) for k, v in d.items()})
        fr = os.path.join(src, "forward_simulation_results.json")
        if os.path.exists(fr):
            d = json.load(open(fr))
            print(name, "forward_simulation_results keys:", {k: np.shape(v) for k, v in d.items()})
            out = {}
            for k, v in d.items():
                a = np.array(v, dtype=np.float64)
                if k == "forward_simulation_results":        # 3N x 101 column-major, flattened
                    a = a.reshape(101, -1)                   # -> [101, 3N]
                    out[k] = a[[0, 50, 100]]
                    early = [1, 2, 3, 5, 8, 12, 20, 30]      # the transient (saves are 2 s apart)
                    out["early_index"] = np.array(early)
                    out["forward_simulation_results_early"] = a[early]
                else:
                    out[k] = a
            np.savez_compressed(os.path.join(dst, "trajectory.npz"), **out)
        sr = os.path.join(src, "sensitivity_results.json")
        if os.path.exists(sr):
            d = json.load(open(sr))
            print(name, "sensitivity_results keys:", {k: np.shape(v) for k, v in d.items()})
            np.savez_compressed(os.path.join(dst, "sensitivity.npz"),
                                **{k: np.array(v, dtype=np.float64) for k, v in d.items()
                                   if not isinstance(v, str)})


if __name__ == "__main__":
    main()
