"""Build libiris_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib", "libiris_b200.so")
SOURCES = ["iris_lib.cu", "bvh_build.cpp"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-fopenmp", "-shared", "--expt-relaxed-constexpr",
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False, extra=(), out=None):
    OUT = out or globals()["OUT"]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "iris_b200.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + list(extra) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lgomp"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    # python iris_b200/build.py [--force] [-v] [--out path.so] [-DNAME=VALUE ...]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    build(force="--force" in sys.argv or out is not None, verbose=True, extra=(["-Xptxas", "-v"] if "-v" in sys.argv else []) + defs, out=out)
