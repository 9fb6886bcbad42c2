"""Short driver for ncu: init + a few WRAPGRUP slices (update kernel + T-conjugation) on the headline workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_model, chain_seed
from alf_b200.api import AlfB200
wl = sys.argv[1] if len(sys.argv) > 1 else "hubbard_16x16_beta10"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
nsl = int(sys.argv[3]) if len(sys.argv) > 3 else 6
model, nwrap, _ = make_model(wl)
g = AlfB200(model, n_chains=C, nwrap=nwrap)
g.set_seeds([chain_seed(c) for c in range(C)]); g.fields_set(); g.init_sweep()
g.kernel_timing(0xff)
t = time.time()
for nt in range(nsl):
    g.wrapgrup(nt)
print("wall per slice ms", 1e3 * (time.time() - t) / nsl, {k: (round(v[0] / max(v[1], 1), 4), v[1]) for k, v in g.kernel_stats().items() if v[1]})
g.close()
