"""Experimental (library built with ALF_QR_PROF=1): clock64 accounting of the phases of k_qrp_reg over one stabilisation interval."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_model, chain_seed
from alf_b200.api import AlfB200, lib
wl = sys.argv[1] if len(sys.argv) > 1 else "hubbard_16x16_beta10"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
model, nwrap, _ = make_model(wl)
g = AlfB200(model, n_chains=C, nwrap=nwrap)
g.set_seeds([chain_seed(c) for c in range(C)]); g.fields_set(); g.init_sweep()
out = (ctypes.c_ulonglong * 16)()
lib().alf_b200_qr_prof_read(out, 1)
g.kernel_timing(0xff)
g.udv_reset(1, "r"); g.wrapur(0, nwrap); g.cgr(1)
st = g.kernel_stats()
lib().alf_b200_qr_prof_read(out, 0)
v = np.array(list(out)[:12], dtype=float); names = ["norms0", "select+swap", "load panel", "PANEL loop rest", "writeback V", "Gram+T", "strips", "finish", "  col: argmax", "  col: owner+bar1", "  col: apply (warp 0)", "  col: bar2 wait"]
v[3] -= v[8:12].sum()      # the kernel-level marker spans the whole column loop
n_launch = st["qrp"][1]; per = v / (2 * C * n_launch)      # per matrix (CTA)
print("qrp avg ms", st["qrp"][0] / n_launch, "launches", n_launch)
for nm, x in zip(names, per):
    print(f"  {nm:20s} {x:10.0f} cycles/matrix  {100 * x / per.sum():5.1f} %")
print("  total cycles/matrix", per.sum(), "=", per.sum() / 1.965e3, "us")
g.close()
