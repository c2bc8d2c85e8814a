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

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
    # the CUDA runtime as a shared library (libcudart.so.12: the toolkit's under /usr/local/cuda/lib64, or the copy a
    # host process such as PyTorch has already loaded) instead of the default static archive
    "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
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
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".inl"))]
    out.append(os.path.join(HERE, "..", "include", "plmatch.h"))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(LIB, cuda_sources()):
        extra = os.environ.get("PLM_BUILD_DEFINES", "").split()  # e.g. -DPLM_TIMELINE (debug phase stamps)
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "plmatch.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout, res.stderr, file=sys.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libplmatch.so")
    build_tools(force=force)
    return LIB


def build_tools(force: bool = False):
    """tools/latency_bench: C-level latency of the host-buffer entry points (no Python in the loop)."""
    src = os.path.join(HERE, "..", "tools", "latency_bench.cpp")
    out = os.path.join(HERE, "..", "tools", "latency_bench")
    if not os.path.exists(src) or not (force or _stale(out, [src, LIB])):
        return
    cmd = ["g++", "-O2", "-std=c++17", src, "-I", os.path.join(HERE, "..", "include"), "-L", LIBDIR, "-lplmatch",
           "-Wl,-rpath,$ORIGIN/../pl_inertial_slam_b200/lib", "-ldl", "-lpthread", "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout, res.stderr, file=sys.stderr)
        raise RuntimeError("g++ failed building tools/latency_bench")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
