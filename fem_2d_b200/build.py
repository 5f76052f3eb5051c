"""Builds libfem2d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The exact-arithmetic translation units are compiled with FMA contraction disabled (-fmad=false for device code,
-ffp-contract=off for host code); everything carries -lineinfo so ncu source pages map back to the kernels.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FEM2D_VARIANT=<tag> (tuning only): a second copy of the library built with FEM2D_NVCC_FLAGS into _variants/<tag>/, loaded by
# FEM2D_LIB=<path>; the product library is always fem_2d_b200/libfem2d_b200.so
_VARIANT = os.environ.get("FEM2D_VARIANT", "")
OUT_DIR = os.path.join(HERE, "_variants", _VARIANT, "_build") if _VARIANT else os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "_variants", _VARIANT, "libfem2d_b200.so") if _VARIANT else os.path.join(HERE, "libfem2d_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall"] + os.environ.get("FEM2D_NVCC_FLAGS", "").split()

# (source, extra flags)
UNITS = [
    ("capi.cu", []),
    ("device_plan.cu", []),
    ("kernels_exact.cu", ["-fmad=false"]),
    ("kernels_scatter.cu", ["-fmad=false"]),
    ("kernels_fast.cu", []),
    ("fields.cu", ["-fmad=false"]),
    ("petsc.cu", []),
    ("plan_host.cpp", []),
    ("host/host_capi.cpp", []),
]


def _stale(obj: str, deps) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_headers():
    hs = []
    for root, _, files in os.walk(CSRC):
        hs += [os.path.join(root, f) for f in files if f.endswith((".h", ".hpp", ".cuh"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return hs


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = _all_headers()
    objs = []
    for src, extra in UNITS:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OUT_DIR, src.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _stale(obj, [sp, __file__] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
