"""Compile the sm_100a CUDA sources into the in-tree C-ABI library.

    python -m wave_mamba_b200.build [--force]

Produces wave_mamba_b200/libwavemamba_b200.so (git-ignored; it travels to the GPU box with
the repo snapshot).  nvcc cross-compiles without a GPU.  No torch headers are involved: the
library's interface is the plain C ABI in include/wavemamba_b200.h.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libwavemamba_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "wavemamba_b200.h")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# per-file extras: the Haar kernels must not contract mul+add (bit-exact with the reference)
EXTRA = {"haar.cu": ["-fmad=false"]}
SOURCES = ["abi.cu", "haar.cu", "ss2d.cu", "ss2d_bwd.cu", "pointwise.cu", "pw_dw_tc5.cu", "spatial32.cu", "pixelwise.cu", "lfss_out_tma.cu", "pw_tma.cu", "gram.cu", "conv3x3_tc5.cu", "skff.cu", "imgio.cu", "metrics.cu", "train.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libwavemamba_b200.so")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    shared_deps = [HEADER, os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc5_common.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "mma_frag.cuh"),
                   os.path.join(CSRC, "ss2d_common.cuh"),
                   os.path.abspath(__file__)]
    objs = []
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC wrapper is not a valid nvcc host compiler override
    env.pop("CXX", None)
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        opath = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(opath)
        if force or _stale(opath, [spath] + shared_deps):
            cmd = [nvcc, "-c", spath, "-o", opath] + ARCH + COMMON + EXTRA.get(src, [])
            if verbose:
                cmd += ["-Xptxas", "-v"]
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True, env=env)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(path)
