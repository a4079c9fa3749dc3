"""In-tree build of the CUDA library (``pytsc_b200/libtsc_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU.  ``-fmad=false``: the kinematics are compared
bit for bit with the fp64 CPU oracle, so multiply-adds must not be fused.

Staleness is decided by CONTENT, not by mtime: the hash of the source, the header
and the flags the library was built from is kept beside it
(``libtsc_b200.so.hash``); a prebuilt library that travelled with the tree (it is
git-ignored but not gpurun-ignored) is reused only when that hash matches the
sources next to it.  ``source_hash()`` is what bench.py / smoke() print as
``build_hash``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "tsc_b200.cu")
LIB = os.path.join(HERE, "libtsc_b200.so")
STAMP = LIB + ".hash"
INCLUDE = os.path.join(ROOT, "include")
HEADER = os.path.join(INCLUDE, "tsc_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def extra_flags():
    return os.environ.get("TSC_B200_NVCC_EXTRA", "").split()      # e.g. -DTSC_DIV_POS_NO_BARRIER for an A/B build


def source_hash() -> str:
    h = hashlib.sha256()
    for path in (SRC, HEADER):
        with open(path, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    h.update(" ".join(NVCC_FLAGS + extra_flags()).encode())
    return h.hexdigest()


def built_hash():
    try:
        with open(STAMP) as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build():
    return not os.path.exists(LIB) or built_hash() != source_hash()


def build_variant(out, flags):
    """A side build of the same source with extra -D flags (phase timing, A/B experiments) for TSC_B200_LIB."""
    cmd = [nvcc_path()] + NVCC_FLAGS + list(flags) + ["-I", INCLUDE, "-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    want = source_hash()
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", tmp, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(want + "\n")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print("build_hash", source_hash()[:16])
