import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_chain, obs_scal_tables
for C in (2048, 8192, 16384):
    m = hubbard_chain(4, 1.0, 0.1, projector=True, theta=5.0)
    g = AlfB200(m, n_chains=C, nwrap=10); g.set_obs_scal_tables(obs_scal_tables(m))
    g.set_seeds([1000 + 7 * i for i in range(C)]); g.fields_set(); g.init_sweep()
    g.sweep(2, 0); g.obs_reset()
    g.kernel_timing(0xff)
    t = time.time(); g.sweep(5, 0); dt = time.time() - t
    st = g.kernel_stats()
    ob = g.obs(); print(C, 'chains: 5 sweeps', round(dt, 3), 's; E =', ob[8] / ob[1], {k: round(v[0], 1) for k, v in st.items() if v[1]})
    g.close()
