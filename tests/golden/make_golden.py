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
from alf_b200.model import hubbard_square, z2_matter_square  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    # name: (model kwargs, nwrap, seeds)
    "sweep_hubbard4x4_mz": (dict(L1=4, L2=4, beta=1.0, dtau=0.1, U=4.0), 5, [814748, 2741]),
    "sweep_hubbard4x4_su2": (dict(L1=4, L2=4, beta=1.0, dtau=0.1, U=4.0, Mz=False), 5, [9813457, 351]),
    "sweep_hubbard4x2_ragged": (dict(L1=4, L2=2, beta=0.7, dtau=0.1, U=2.0, symm=False), 3, [77123]),
    # continuous HS fields (type 3): real-valued field arrays
    "sweep_hubbard4x4_continuous": (dict(L1=4, L2=4, beta=0.8, dtau=0.1, U=4.0, continuous=True), 4, [5501, 90210]),
    # BASELINE config 5 at test size: Ising action tables, restricted sequential range, N_Global_tau = 4 star moves per slice (projective)
    "sweep_z2_matter4x4_projector": (dict(kind="z2_matter", L1=4, L2=4, beta=0.4, dtau=0.1, projector=True, theta=0.3, g=0.8, K=0.5, J=0.7, h=0.9), 5, [424242, 1717]),
}


def build(kw):
    kw = dict(kw); kind = kw.pop("kind", "hubbard")
    if kind == "z2_matter":
        return z2_matter_square(kw.pop("L1"), kw.pop("L2"), **kw)
    return hubbard_square(kw.pop("L1"), kw.pop("L2"), **kw)


def field_array(f):
    """nsigma%f as stored in the vectors: int8 for discrete configurations, float64 when a field is genuinely continuous."""
    r = np.ascontiguousarray(f.real)
    return r.astype(np.int8) if np.array_equal(r, np.rint(r)) and np.abs(r).max() < 100 else r


def run_case(kw, nwrap, seeds):
    model = build(kw)
    out = {"seeds": np.asarray(seeds, dtype=np.int64), "nwrap": np.int64(nwrap)}
    for c, s in enumerate(seeds):
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set()
        out[f"fields0_{c}"] = field_array(o.get_fields())
        o.init()
        out[f"g_init_{c}"] = np.stack([o.green(nf) for nf in range(1, model.N_FL + 1)])
        out[f"phase_init_{c}"] = np.complex128(o.phase())
        o.log(True); o.sweep(0)
        acc, _ = o.get_log()
        out[f"accept_{c}"] = np.packbits(acc.astype(np.uint8))
        out[f"n_accept_{c}"] = np.int64(acc.size)
        out[f"fields1_{c}"] = field_array(o.get_fields())
        out[f"g_sweep_{c}"] = np.stack([o.green(nf) for nf in range(1, model.N_FL + 1)])
        out[f"phase_sweep_{c}"] = np.complex128(o.phase())
        out[f"rng_{c}"] = np.asarray(o.rng_state(), dtype=np.uint64)
    return out


if __name__ == "__main__":
    only = sys.argv[1:]      # optional: names of the cases to (re)generate; existing vectors are left untouched otherwise
    for name, (kw, nwrap, seeds) in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(kw, nwrap, seeds))
        print("wrote", name)
