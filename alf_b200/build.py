"""Builds alf_b200/libalf_b200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.
Three translation units (C-ABI, real instantiation, complex instantiation) are compiled in parallel and linked."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libalf_b200.so")
UNITS = ["alf_b200.cu", "alf_inst_real.cu", "alf_inst_cplx.cu"]
OBJ = os.path.join(HERE, "csrc", "_obj")


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
    out.append(os.path.join(HERE, "..", "include", "alf_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        flags.append("-Xptxas=-v")
    if os.environ.get("ALF_QR_PROF"):      # experimental: phase accounting inside k_qrp_reg (alf_qrblk2.cuh)
        flags.append("-DALF_QR_PROF")
    if os.environ.get("ALF_UPD_PROF"):     # experimental: phase accounting inside k_wrapgr_fast (alf_update_fast.cuh)
        flags.append("-DALF_UPD_PROF")

    def cc(u):
        o = os.path.join(OBJ, u.replace(".cu", ".o"))
        subprocess.check_call([nvcc] + flags + ["-c", os.path.join(CSRC, u), "-o", o])
        return o
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(cc, UNITS))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
