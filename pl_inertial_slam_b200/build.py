"""In-tree build of the CUDA library (sm_100a only).

``python -m pl_inertial_slam_b200.build`` compiles csrc/plmatch.cu into lib/libplmatch.so with
nvcc; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libplmatch.so")
REFERENCE_INC = "/root/reference/stvo-pl/include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "plmatch.h"))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(LIB, cuda_sources()):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "plmatch.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout, res.stderr, file=sys.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libplmatch.so")
    build_stvo_wrapper(force=force, verbose=verbose)
    return LIB


STVO_LIB = os.path.join(LIBDIR, "libstvo_matching_gpu.so")


def build_stvo_wrapper(force: bool = False, verbose: bool = False):
    """The C++ drop-in for stvo-pl/src/matching.cpp needs the reference's own headers
    (matching.h, gridStructure.h, config.h); it is built only where /root/reference exists and the
    prebuilt .so travels to the GPU box."""
    src = os.path.join(CSRC, "stvo_matching_gpu.cpp")
    if not os.path.exists(src) or not os.path.isdir(REFERENCE_INC):
        return None
    shim = os.path.join(HERE, "..", "oracle", "shim")
    harness = os.path.join(CSRC, "stvo_harness.cpp")
    ref_src = "/root/reference/stvo-pl/src"
    sources = [src, harness]
    if not (force or _stale(STVO_LIB, sources + [LIB])):
        return STVO_LIB
    cmd = ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I", shim, "-I", REFERENCE_INC, "-I", os.path.join(HERE, "..", "include"),
           src, harness, os.path.join(ref_src, "gridStructure.cpp"), os.path.join(ref_src, "lineIterator.cpp"),
           os.path.join(shim, "config_stub.cpp"),
           "-o", STVO_LIB, "-L", LIBDIR, "-lplmatch", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout, res.stderr, file=sys.stderr)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building libstvo_matching_gpu.so")
    return STVO_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
