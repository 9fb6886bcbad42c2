"""One handle alone: per-category device time (CUDA events around every launch) and algorithmic FP64 throughput of one sweep.
usage: time_sweep.py [workload] [chains] [ltau] [obs_tau]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_model, chain_seed
from alf_b200.api import AlfB200, fp64_peak
wl = sys.argv[1] if len(sys.argv) > 1 else "hubbard_16x16_beta10"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
ltau = int(sys.argv[3]) if len(sys.argv) > 3 else 1
obs = int(sys.argv[4]) if len(sys.argv) > 4 else 0
model, nwrap, _ = make_model(wl)
g = AlfB200(model, n_chains=C, nwrap=nwrap)
g.set_seeds([chain_seed(c) for c in range(C)]); g.fields_set(); g.init_sweep()
if obs: g.obs_tau_enable(True)
g.sweep(1, ltau)
g.kernel_timing(0xff); g.sweep(1, ltau)
st = g.kernel_stats(); fl = g.kernel_flops(); g.kernel_timing(0)
dfma, dmma = fp64_peak(0)
out = {"workload": wl, "chains": C, "ltau": ltau, "obs_tau": obs, "total_ms": round(sum(v[0] for v in st.values()), 2), "dmma_peak": round(dmma, 2)}
for k, (ms, n) in st.items():
    if n:
        out[k] = {"ms": round(ms, 2), "n": n}
        if fl.get(k, 0) > 0: out[k]["frac_dmma"] = round(fl[k] / (ms * 1e-3) / 1e12 / dmma, 3)
print(json.dumps(out))
c = g.control(); print("precision", c["XMAXG"], c["XMEANG"] / max(c["NCG"], 1), c["XMAX_tau"])
g.close()
