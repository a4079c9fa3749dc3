"""In-tree build of the CUDA library (``pytsc_b200/libtsc_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU.  ``-fmad=false``: the kinematics are compared
bit for bit with the fp64 CPU oracle, so multiply-adds must not be fused.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "tsc_b200.cu")
LIB = os.path.join(HERE, "libtsc_b200.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    newest = max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(INCLUDE, "tsc_b200.h")))
    return os.path.getmtime(LIB) < newest


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("TSC_B200_NVCC_EXTRA", "").split()      # e.g. -DTSC_DIV_POS_BARRIER for an A/B build
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
