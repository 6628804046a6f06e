"""Builds libmapf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# MAPF_B200_LIB points the package at another build of the library (A/B runs of two builds on one GPU box); such a
# library is used as it is, never rebuilt
LIB_OVERRIDE = os.environ.get("MAPF_B200_LIB")
LIB_PATH = LIB_OVERRIDE or os.path.join(_HERE, "libmapf_b200.so")
SOURCES = ["mapf_abi.cu", "mapf_env_kernels.cu", "mapf_step_kernels.cu", "mapf_reset_kernels.cu", "mapf_per_kernels.cu", "mapf_replay_kernels.cu"]
HEADERS = ["mapf_common.cuh", os.path.join("..", "..", "include", "mapf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libmapf_b200.so")


def is_stale() -> bool:
    if LIB_OVERRIDE:
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
