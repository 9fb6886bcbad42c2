"""Shared helpers of the parity tests: seeds, tolerances and the oracle/GPU drivers."""
import numpy as np

from alf_b200.model import hubbard_square, hubbard_chain, kondo_square

# first lines of Scripts_and_Parameters_files/Start/seeds are arbitrary integers; any fixed list will do for parity
SEEDS = [814748, 2741, 9813457, 351, 77123, 56001, 120033, 4599, 7777, 1234567, 42, 987654]

TOL_G = 1e-10          # north_star check (1): relative Frobenius norm of freshly recomputed G


def relF(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def config1(**kw):     # BASELINE.json configs[0]
    return hubbard_square(4, 4, beta=5.0, dtau=0.1, U=4.0, **kw)


def config2(**kw):     # configs[1]
    return hubbard_square(8, 8, beta=10.0, dtau=0.1, U=4.0, **kw)


def config3(**kw):     # configs[2]
    return hubbard_square(16, 16, beta=10.0, dtau=0.1, U=4.0, **kw)


def make_oracles(model, seeds, nwrap=10, stab3=False):
    from oracle.oracle import Oracle
    out = []
    for s in seeds:
        o = Oracle(model, nwrap=nwrap, stab3=stab3)
        o.ranset(s)
        o.fields_set()
        out.append(o)
    return out
