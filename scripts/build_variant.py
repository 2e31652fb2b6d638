"""Tuning helper: link a variant of libfmftsaxs.so whose CUDA units are compiled with extra -D flags.

    python scripts/build_variant.py NAME [-DFOO=1 ...]      ->  variants/NAME/libfmftsaxs.so

Select it at run time with SXS_LIB_PATH=variants/NAME/libfmftsaxs.so (libfmftsaxs_b200/capi.py).  variants/ is
git-ignored but travels to the GPU box.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfmftsaxs_b200 import build as B


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    B.build()
    out = os.path.join(B.REPO, "variants", name)
    os.makedirs(out, exist_ok=True)
    nvcc = "nvcc"
    objs = [os.path.join(B.OBJ, s + ".o") for s in B.HOST_SRC]
    for s, extra in B.CUDA_SRC:
        obj = os.path.join(out, s + ".o")
        txt = B._run([nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + B.NVCC_ARCH
                     + extra + flags + B.INCS + ["-c", os.path.join(B.CSRC, "cuda", s), "-o", obj])
        open(obj + ".ptxas.txt", "w").write(txt)
        objs.append(obj)
    lib = os.path.join(out, "libfmftsaxs.so")
    B._run([nvcc, "-shared", "-o", lib] + B.NVCC_ARCH + objs + ["-lquadmath", "-lpthread", "-lm"])
    print(lib)


if __name__ == "__main__":
    main()
