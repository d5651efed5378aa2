"""Build libpgemb_b200.so (the C-ABI shared library: CUDA kernels for sm_100a + host code) in-tree.

    python -m pg_embedding_b200.build          # or: python pg_embedding_b200/build.py

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpgemb_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["common.cuh", "dist_exact.cuh", "search_kernel.cuh", "aux_kernels.cuh", "bind_kernel.cuh", "search_config.h",
           os.path.join("..", "..", "include", "pgemb_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [
        nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-fmad=false",  # exact kernels use explicit _rn intrinsics; never contract anything else either
        "-Xptxas", "-v" if verbose else "-warn-spills",
        "-I", os.path.join(HERE, "..", "include"),
        "-o", OUT,
    ] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + "\n" + res.stderr)
    if verbose:
        print(res.stdout)
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
