"""Short driver for ncu (--profile-from-start off): init + one sweep without TAU_M, then the profiled range = the first stabilisation
interval of TAU_M is not separable from outside, so the whole TAU_M call is profiled: use -k regex:<kernel> -c <n> to bound the capture."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_model, chain_seed
from alf_b200.api import AlfB200
wl = sys.argv[1] if len(sys.argv) > 1 else "hubbard_16x16_beta10"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
model, nwrap, _ = make_model(wl)
g = AlfB200(model, n_chains=C, nwrap=nwrap)
g.set_seeds([chain_seed(c) for c in range(C)]); g.fields_set(); g.init_sweep()
rt = None
for name in ("libcudart.so", "libcudart.so.12"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
if rt: rt.cudaProfilerStart()
g.kernel_timing(0xff)
t = time.time(); g.tau_m()
st = g.kernel_stats()
if rt: rt.cudaProfilerStop()
print("wall ms", 1e3 * (time.time() - t), {k: (round(v[0] / max(v[1], 1), 4), v[1]) for k, v in st.items() if v[1]})
g.close()
