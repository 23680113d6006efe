"""Builds csrc/mgicp.cu into libmgicp.so (in-tree, next to this file) for sm_100a with nvcc.

FMA contraction is disabled (-fmad=false): every decision on the path (voxel index, neighbour order,
d^2 < r^2) has to round exactly like the Open3D-CPU path it replaces.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "mgicp.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "mgicp_device.cuh"), os.path.join(HERE, "csrc", "mgicp_math.cuh"),
        os.path.join(HERE, "csrc", "mgicp_fgr.cuh"), os.path.join(HERE, "csrc", "mgicp_fgr_tc.cuh"),
        os.path.join(HERE, "csrc", "fpfh_math.cuh"), os.path.join(HERE, "csrc", "fgr_math.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "mgicp.h")]
LIB = os.path.join(HERE, "libmgicp.so")


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
           "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-o", LIB, SRC, "-lcudart"]
    cmd[1:1] = os.environ.get("MGICP_NVCC_FLAGS", "").split()       # experiments: e.g. -DMGICP_PREFETCH=1
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
