"""Builds alf_b200/libalf_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "alf_b200.cu")
LIB = os.path.join(HERE, "libalf_b200.so")
DEPS = ["alf_b200.cu", "alf_types.cuh", "alf_la.cuh", "alf_la_host.cuh", "alf_ops.cuh", "alf_update.cuh", "alf_update_fast.cuh", "alf_taum.cuh", "alf_obs.cuh"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPS + [os.path.join("..", "..", "include", "alf_b200.h")]:
        p = os.path.join(HERE, "csrc", d)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=False))
