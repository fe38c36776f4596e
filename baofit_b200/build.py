"""In-tree build of libbaofit_b200.so with nvcc for sm_100a (called by __graft_entry__.build()).

The shared library holds the hand-written CUDA kernels, the C ABI of include/baofit_b200.h and
the host-side C++ mirror of the reference interfaces (baofit_b200/host/). It is built in-tree
so that it travels with the repository snapshot to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libbaofit_b200.so")
BUILD = os.path.join(HERE, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fopenmp"]


def _nvcc():
    for c in ("nvcc", "/usr/local/cuda/bin/nvcc"):
        p = shutil.which(c) or (c if os.path.exists(c) else None)
        if p:
            return p
    raise RuntimeError("nvcc not found")


def sources():
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cc")))
    src += sorted(glob.glob(os.path.join(HOST, "*.cc")))
    return src


def headers():
    return (glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
            glob.glob(os.path.join(HOST, "*.h")) + [os.path.join(ROOT, "include", "baofit_b200.h")])


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = max(os.path.getmtime(h) for h in headers())
    # the host compiler in $CC/$CXX of this image lacks OpenMP specs; use the system g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs, rebuilt = [], False
    for s in sources():
        o = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            cmd = [nvcc] + ccbin + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + \
                  ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-I", HOST, "-x", "cu", "-dc" if False else "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc] + ccbin + ["-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread", "-lgomp"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(LIB)
