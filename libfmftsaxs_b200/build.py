"""Builds libfmftsaxs_b200/libfmftsaxs.so (host C11 + CUDA sm_100a) in-tree with gcc and nvcc.

    python -m libfmftsaxs_b200.build

Host sources are compiled with -ffp-contract=off, sxs_exact.cu with -fmad=false: both restate
reference arithmetic that must not be fused ("Bit-faithful pieces").
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libfmftsaxs.so")

HOST_SRC = ["sfbessel.c", "saxs_utils.c", "tables.c", "form_factor_table.c", "pdb2spf.c", "profile.c",
            "min_saxs.c", "index.c", "index_rows.c", "ft_fast.c", "fftsaxs.c", "partition.c", "mol2_mini.c", "flat_api.c"]
CUDA_SRC = [("sxs_score.cu", []), ("sxs_expand.cu", []), ("sxs_exact.cu", ["-fmad=false"])]
TOOLS = ["correlate", "single_saxs", "score_ft_naive"]

INCS = ["-I" + os.path.join(REPO, "include"), "-I" + os.path.join(REPO, "include", "fmftsaxs"),
        "-I" + os.path.join(CSRC, "host"), "-I" + os.path.join(CSRC, "cuda")]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def _headers():
    hs = []
    for root in (os.path.join(REPO, "include"), CSRC):
        for d, _, files in os.walk(root):
            hs += [os.path.join(d, f) for f in files if f.endswith((".h", ".cuh"))]
    return hs


def build(verbose=False, force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    for s in HOST_SRC:
        src = os.path.join(CSRC, "host", s)
        obj = os.path.join(OBJ, s + ".o")
        if force or _newer(src, obj, hdrs):
            _run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wno-stringop-truncation", "-pthread"]
                 + INCS + ["-c", src, "-o", obj])
        objs.append(obj)
    for s, extra in CUDA_SRC:
        src = os.path.join(CSRC, "cuda", s)
        obj = os.path.join(OBJ, s + ".o")
        if force or _newer(src, obj, hdrs):
            out = _run([nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + NVCC_ARCH
                       + extra + INCS + ["-c", src, "-o", obj])
            with open(os.path.join(OBJ, s + ".ptxas.txt"), "w") as f:
                f.write(out)
            if verbose:
                print(out)
        objs.append(obj)
    if force or any(_newer(o, LIB) for o in objs):
        _run([nvcc, "-shared", "-o", LIB] + NVCC_ARCH + objs + ["-lquadmath", "-lpthread", "-lm"])
    for t in TOOLS:
        src = os.path.join(CSRC, "tools", t + ".c")
        exe = os.path.join(PKG, "bin", t)
        if os.path.exists(src) and (force or _newer(src, exe, hdrs + [LIB])):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            _run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-D_SXS_VERBOSE_"] + INCS + [src, "-o", exe,
                 "-L" + PKG, "-lfmftsaxs", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
