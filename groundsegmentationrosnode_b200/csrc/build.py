"""Builds libgs_b200.so (the C-ABI library of include/gs_b200.h) in-tree with nvcc for sm_100a.

    python groundsegmentationrosnode_b200/csrc/build.py [--force] [--verbose]

-fmad=false: no FMA contraction anywhere in device code — the reference's x86-64 build has none and
the parity-critical arithmetic must round like the host's.  Kernels that want an FMA ask for it explicitly (__fmaf_rn).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "libgs_b200.so")
SOURCES = ["gs_capi.cu", "gs_kernels.cu", "gs_fused.cu", "gs_multi.cpp"]
HEADERS = ["gs_internal.h", "gs_math.cuh", os.path.join("..", "..", "include", "gs_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden,-ffp-contract=off", "--shared", "-cudart", "static"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    extra = os.environ.get("GS_EXTRA_NVCC_FLAGS", "").split()  # experiments (e.g. -DGS_UF_DEBUG)
    cmd = [NVCC, *FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), *[os.path.join(HERE, s) for s in SOURCES], "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build("--force" in sys.argv, "--verbose" in sys.argv))
