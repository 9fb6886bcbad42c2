"""Times one batched sweep of Hamiltonian_Z2_Matter 12x12 (BASELINE config 5) on the device and prints the per-category kernel time."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from alf_b200.api import AlfB200, lib
from alf_b200.model import z2_matter_square
import ctypes as C

L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
chains = int(sys.argv[4]) if len(sys.argv) > 4 else 148
m = z2_matter_square(L, L, beta=beta, dtau=0.1, projector=True, theta=theta)
print("Ndim", m.Ndim, "Ltrot", m.Ltrot, "M", m.n_opv, "N_part", m.N_part, m.global_tau, flush=True)
g = AlfB200(m, n_chains=chains, nwrap=10)
g.set_seeds([1000 + i for i in range(chains)]); g.fields_set()
t0 = time.time(); g.init_sweep(); print("init", time.time() - t0, flush=True)
lib().alf_b200_kernel_timing(g.h, 0xff)
t0 = time.time(); g.sweep(1, 0); dt = time.time() - t0
ms = (C.c_double * 8)(); ln = (C.c_long * 8)(); lib().alf_b200_get_kernel_stats(g.h, ms, ln)
print(json.dumps({"sweep_s": dt, "sweeps_per_s": chains / dt, "ms": list(ms), "launches": list(ln), "phase0": str(g.phase()[0]), "ctl": list(g.control()) if hasattr(g, "control") else None}))
