"""In-tree build of the CUDA library (sm_100a only).

``libmcl_b200.so`` = csrc/mcl_b200.cu (kernels + C ABI) + csrc/map_prep.cpp, compiled with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo``.  nvcc cross-compiles without a
GPU, so this runs in the build container; the resulting .so travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libmcl_b200.so")
SOURCES = [os.path.join(CSRC, "mcl_b200.cu"), os.path.join(CSRC, "map_prep.cpp")]
HEADERS = [os.path.join(CSRC, h) for h in
           ("kernels.cuh", "map_kernels.cuh", "dir_kernels.cuh", "exact_kernels.cuh", "shard.cuh", "march.cuh", "dirmap.cuh", "exact_sum.cuh",
            "device_utils.cuh", "map_prep.h")] + [
    os.path.join(os.path.dirname(_HERE), "include", "mcl_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the MCL library is CUDA-only and cannot be built without it")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libmcl_b200.so if missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB_PATH
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    cmd = [_nvcc()] + NVCC_FLAGS + (["-ccbin", host_cxx] if host_cxx else []) + (
        ["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
