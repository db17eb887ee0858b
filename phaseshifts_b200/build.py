"""In-tree build of ``libphsh_b200.so`` (hand-written CUDA for sm_100a + the C ABI).

``nvcc`` cross-compiles without a GPU.  ``-fmad=false`` keeps the strict integrators free of
FMA contraction (the FAST variants ask for FMAs explicitly); ``-lineinfo`` lets ncu map SASS
to source.  The .so stays in the package directory so it travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
OUT = os.path.join(_PKG, "libphsh_b200.so")
SOURCES = ["phsh_b200.cu", "phsh_files.cpp", "phsh_conphas_files.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return os.path.getmtime(os.path.join(_PKG, "..", "include", "phsh_b200.h")) > t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
