import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_chain, obs_scal_tables
dt = 1.0 / 7.0
for seed0 in (1000, 555555):
    m = hubbard_chain(4, 2.0, dt); C = 4096
    g = AlfB200(m, n_chains=C, nwrap=10); g.set_obs_scal_tables(obs_scal_tables(m))
    g.set_seeds([seed0 + 13 * i for i in range(C)]); g.fields_set(); g.init_sweep(); g.sweep(40, 0)
    b = []
    for _ in range(128):
        g.obs_reset(); g.sweep(4, 0); ob = g.obs(); b.append(ob[8] / ob[1])
    b = np.asarray(b); out = [f"seed0 {seed0}: mean {b.mean():.6f}"]
    for k in (1, 2, 4, 8, 16, 32):
        bb = b[: len(b) // k * k].reshape(-1, k).mean(1); out.append(f"bin {4*k} sweeps: err {bb.std(ddof=1) / np.sqrt(len(bb)):.2e}")
    print("; ".join(out)); g.close()
