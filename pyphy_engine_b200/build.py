"""Build the C-ABI shared library (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m pyphy_engine_b200.build [--force]

Produces ``pyphy_engine_b200/_lib/libpyphy_b200.so``.  nvcc cross-compiles without
a GPU; the .so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libpyphy_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
# (source, extra flags).  The fp64 "compat" kernels must not contract a*b+c into FMAs.
UNITS = [
    ("ppb_f32.cu", ["-prec-div=false", "-prec-sqrt=false"]),
    ("ppb_f64.cu", ["-fmad=false"]),
    ("ppb_cairo.cu", ["-fmad=false"]),
    ("ppb_pack.cu", []),
    ("ppb_api.cu", []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_source() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for fn in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, fn)))
    return t


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
