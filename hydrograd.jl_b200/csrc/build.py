"""Builds libhydrograd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libhydrograd_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
COMMON += os.environ.get("HG_NVCC_EXTRA", "").split()   # experiment builds (e.g. -DHG_PHASE_CLOCKS); empty for the product
SOURCES = [
    ("hg_host.cpp", ["-Xcompiler", "-fopenmp"]),     # the tile builder runs on all host cores
    ("hg_srh.cpp", []),
    ("hg_partition.cpp", []),
    ("hg_results.cpp", []),
    ("hg_api.cu", ["-Xcompiler", "-fopenmp"]),       # hg_create permutes the per-cell fields on all host cores
    ("hg_plain.cu", ["-fmad=false"]),      # reference evaluation order, no FMA contraction
    ("hg_jvp.cu", ["-fmad=false"]),        # forward mode on the plain tables, same arithmetic rules
    ("hg_fused.cu", []),
    ("hg_vjp.cu", []),
    ("hg_fjvp.cu", []),
    ("hg_ude.cu", []),
    ("hg_comm.cu", []),
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(HERE, "hg_ctx.h"), os.path.join(HERE, "hg_device.cuh"), os.path.join(HERE, "hg_ude.h"), os.path.join(HERE, "hg_jvp_impl.h"), os.path.join(HERE, "hg_case.h"), os.path.join(PKG, "..", "include", "hydrograd_b200.h"), __file__]
    objs, cmds = [], []
    for src, extra in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmds.append(["nvcc"] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", s, "-o", o])
    if cmds:   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)

        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(run, cmds))
    if force or _newer(OUT, objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-cudart", "static", "-o", OUT] + objs + ["-lgomp"]
        print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
