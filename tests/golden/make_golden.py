#!/usr/bin/env python
"""Generates tests/golden/sweep_*.npz: first-sweep accept/reject sequences, fields, phases and freshly recomputed Green
functions of small Hubbard runs, produced by the CPU oracle (oracle/alf_oracle.cpp).

The reference itself (Fortran 2008 + LAPACK + MPI) cannot be built or run in this image (no gfortran), so these vectors
pin the ORACLE's output at the commit that generated them, not ALF.out's; the reference's own known answers that are
usable without its binary (test 26 value, closed-form inputs of tests 13/14/15/20/23, ED energies) are checked in
tests/test_oracle_cpu.py.   usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from alf_b200.model import hubbard_square  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    # name: (model kwargs, nwrap, seeds)
    "sweep_hubbard4x4_mz": (dict(L1=4, L2=4, beta=1.0, dtau=0.1, U=4.0), 5, [814748, 2741]),
    "sweep_hubbard4x4_su2": (dict(L1=4, L2=4, beta=1.0, dtau=0.1, U=4.0, Mz=False), 5, [9813457, 351]),
    "sweep_hubbard4x2_ragged": (dict(L1=4, L2=2, beta=0.7, dtau=0.1, U=2.0, symm=False), 3, [77123]),
}


def build(kw):
    kw = dict(kw)
    return hubbard_square(kw.pop("L1"), kw.pop("L2"), **kw)


def run_case(kw, nwrap, seeds):
    model = build(kw)
    out = {"seeds": np.asarray(seeds, dtype=np.int64), "nwrap": np.int64(nwrap)}
    for c, s in enumerate(seeds):
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set()
        out[f"fields0_{c}"] = o.get_fields().real.astype(np.int8)
        o.init()
        out[f"g_init_{c}"] = np.stack([o.green(nf) for nf in range(1, model.N_FL + 1)])
        out[f"phase_init_{c}"] = np.complex128(o.phase())
        o.log(True); o.sweep(0)
        acc, _ = o.get_log()
        out[f"accept_{c}"] = np.packbits(acc.astype(np.uint8))
        out[f"n_accept_{c}"] = np.int64(acc.size)
        out[f"fields1_{c}"] = o.get_fields().real.astype(np.int8)
        out[f"g_sweep_{c}"] = np.stack([o.green(nf) for nf in range(1, model.N_FL + 1)])
        out[f"phase_sweep_{c}"] = np.complex128(o.phase())
        out[f"rng_{c}"] = np.asarray(o.rng_state(), dtype=np.uint64)
    return out


if __name__ == "__main__":
    for name, (kw, nwrap, seeds) in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(kw, nwrap, seeds))
        print("wrote", name)
