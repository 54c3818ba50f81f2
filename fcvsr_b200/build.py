"""Build libfcvsr_b200.so (the C-ABI kernel library) in-tree with nvcc for sm_100a.

    python -m fcvsr_b200.build            # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so lands in fcvsr_b200/_lib/ (git-ignored, travels with the
working tree).  Sources are recompiled only when newer than their object file.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libfcvsr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "--cudart", "shared"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, s[:-3] + ".o")
        objs.append(obj)
        if _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if _stale(LIB, objs):
        cmd = [NVCC, "-shared", "--cudart", "shared", "-o", LIB] + objs + ["-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
