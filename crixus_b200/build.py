"""Builds crixus_b200/libcrixus_b200.so (CUDA kernels + C-ABI + host glue) in-tree with nvcc for sm_100a, and the
`crixus_b200/bin/crixus` command-line front end.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcrixus_b200.so")
CLI = os.path.join(HERE, "bin", "crixus")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-fno-fast-math", "-Xptxas", "-v" if os.environ.get("CX_PTXAS_V") else "-O3"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    host = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    lib_host = [h for h in host if os.path.basename(h) != "main.cpp"]
    deps = cu + host + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.h")) + \
        [os.path.join(HERE, "..", "include", "crixus_b200.h")]
    if force or _newer(LIB, deps):
        cmd = ["nvcc"] + NVCC_FLAGS + os.environ.get("CX_NVCC_EXTRA", "").split() + ["-shared", "-o", LIB] + cu + lib_host + ["-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    main = os.path.join(CSRC, "host", "main.cpp")
    if os.path.exists(main) and (force or _newer(CLI, deps)):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cli_src = [main]
        cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(HERE, "..", "include")] + cli_src + ["-o", CLI,
               "-L" + HERE, "-lcrixus_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/..",
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
