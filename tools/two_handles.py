"""Experiment: H handles (own stream each) with C/H chains, driven by H host threads, against one handle with C chains.
usage: two_handles.py [chains_total] [handles] [sweeps] [ltau]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_model, chain_seed
from alf_b200.api import AlfB200
C = int(sys.argv[1]) if len(sys.argv) > 1 else 148
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ltau = int(sys.argv[4]) if len(sys.argv) > 4 else 1
model, nwrap, _ = make_model("hubbard_16x16_beta10")
hs = []
for k in range(H):
    n = C // H
    g = AlfB200(model, n_chains=n, nwrap=nwrap); g.set_seeds([chain_seed(k * n + c) for c in range(n)]); g.fields_set(); g.init_sweep(); hs.append(g)
def run(g, n):
    g.sweep(n, ltau)
th = [threading.Thread(target=run, args=(g, 1)) for g in hs]; [t.start() for t in th]; [t.join() for t in th]      # warm-up
t0 = time.time()
th = [threading.Thread(target=run, args=(g, nsw)) for g in hs]; [t.start() for t in th]; [t.join() for t in th]
dt = time.time() - t0
print(f"handles {H} x {C // H} chains: {1e3 * dt / nsw:.1f} ms per sweep, {(C // H) * H * nsw / dt:.1f} sweeps/s")
for g in hs: g.close()
