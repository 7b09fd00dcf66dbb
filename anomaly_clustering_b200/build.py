"""Build libac_b200.so (hand-written CUDA for sm_100a + the C ABI of include/ac_b200.h) in-tree.

    python -m anomaly_clustering_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  cudart is linked dynamically so that the library shares the CUDA runtime
instance (current device, streams) of the host process (PyTorch loads libcudart.so.12 first).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "csrc", "build")
LIB = os.path.join(PKG, "libac_b200.so")
SOURCES = ["embed.cu", "mindist_simt.cu", "mindist_tc.cu", "refine.cu", "stage3.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    deps = [os.path.join(CSRC, src), os.path.join(CSRC, "common.cuh"), os.path.join(PKG, "..", "include", "ac_b200.h")]
    if _stale(obj, deps):
        flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
        cmd = [_nvcc(), *ARCH, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, os.path.splitext(src)[0] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        if verbose:
            print("compiled", src)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
        if os.path.exists(LIB):
            os.remove(LIB)
    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if _stale(LIB, objs):
        cmd = [_nvcc(), *ARCH, "-shared", "-cudart", "shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
