"""Experimental (library built with ALF_UPD_PROF=1): clock64 accounting of the phases of k_wrapgr_fast over a few slices."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_model, chain_seed
from alf_b200.api import AlfB200, lib
wl = sys.argv[1] if len(sys.argv) > 1 else "hubbard_16x16_beta10"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
nsl = 6
model, nwrap, _ = make_model(wl)
g = AlfB200(model, n_chains=C, nwrap=nwrap)
g.set_seeds([chain_seed(c) for c in range(C)]); g.fields_set(); g.init_sweep()
out = (ctypes.c_ulonglong * 16)()
lib().alf_b200_upd_prof_read(out, 1)
g.kernel_timing(0xff); c0 = g.control()
for nt in range(nsl):
    g.wrapgrup(nt)
st = g.kernel_stats(); c1 = g.control()
lib().alf_b200_upd_prof_read(out, 0)
v = np.array(list(out)[:6], dtype=float); names = ["prologue", "windows before flush (acc.)", "flush", "(a) Gw gather + barrier", "(b) decisions + barrier", "(c) factors + prefetch + barrier"]
per = v / (C * nsl)
print("update avg ms per slice", st["update"][0] / st["update"][1], "accepts/slice/chain", (c1["ACC_up"] - c0["ACC_up"]) / (C * nsl), "flushes/slice/chain", (c1["flushes"] - c0["flushes"]) / (C * nsl))
for nm, x in zip(names, per):
    print(f"  {nm:34s} {x:10.0f} cycles/slice  {100 * x / per.sum():5.1f} %")
print("  total cycles/slice", per.sum(), "=", per.sum() / 1.965e3, "us")
g.close()
