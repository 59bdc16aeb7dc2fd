"""Builds crixus_b200/libcrixus_b200.so (CUDA kernels + C-ABI + host glue) in-tree with nvcc for sm_100a, and the
`crixus_b200/bin/crixus` command-line front end.  nvcc cross-compiles without a GPU.

Every translation unit is compiled to an object of its own (in parallel, only when it or a header changed) and the
objects are linked into the shared library; NCCL is NOT a link-time dependency (csrc/dist.cu resolves it with dlopen
when a multi-GPU communicator is created)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcrixus_b200.so")
CLI = os.path.join(HERE, "bin", "crixus")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-fno-fast-math", "-Xptxas", "-v" if os.environ.get("CX_PTXAS_V") else "-O3"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(src, obj, verbose):
    cmd = ["nvcc"] + NVCC_FLAGS + os.environ.get("CX_NVCC_EXTRA", "").split() + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose=False):
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    host = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    lib_host = [h for h in host if os.path.basename(h) != "main.cpp"]
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "host", "*.h")) + \
        [os.path.join(HERE, "..", "include", "crixus_b200.h"), os.path.abspath(__file__)]
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs = [], []
    for s in cu + lib_host:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append((s, o))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for f in [ex.submit(_compile, s, o, verbose) for s, o in jobs]:
                f.result()
    if jobs or force or _newer(LIB, objs):
        cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    main = os.path.join(CSRC, "host", "main.cpp")
    if os.path.exists(main) and (force or _newer(CLI, [main, LIB])):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cmd = ["g++", "-O2", "-std=c++17", "-I" + os.path.join(HERE, "..", "include"), main, "-o", CLI,
               "-L" + HERE, "-lcrixus_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/..",
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
