"""Diagnostic at BASELINE size (Hubbard 16x16, beta=10): first-sweep accept/reject agreement, fresh-G parity and the precision
monitors of the device against the oracle for a few chains (the oracle needs ~15 s per chain-sweep).  usage: config3_parity.py [n_chains]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_square
from oracle.oracle import Oracle
C = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = hubbard_square(16, 16, beta=10.0, dtau=0.1, U=4.0)
seeds = [814748, 2741, 9813457, 351, 77123, 56001, 120033, 4599][:C]
g = AlfB200(model, n_chains=C, nwrap=10); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.accept_log(1)
orcs = []
for s in seeds:
    o = Oracle(model, nwrap=10); o.ranset(s); o.fields_set(); orcs.append(o)
t0 = time.time()
def work(o):
    o.init(); o.log(True); o.sweep(0)
th = [threading.Thread(target=work, args=(o,)) for o in orcs]; [t.start() for t in th]; [t.join() for t in th]
print("oracle: %.1f s for %d chain-sweeps" % (time.time() - t0, C))
g.sweep(1, 0)
log = g.get_accept_log()
relF = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for c, o in enumerate(orcs):
    acc, _ = o.get_log()
    bad = np.nonzero(acc != log[c])[0]
    same_fields = np.array_equal(g.get_fields()[c], o.get_fields())
    e = max(relF(g.green(c, nf), o.green(nf)) for nf in (1, 2))
    oc = o.control()
    print(f"chain {c}: decisions {acc.size}, mismatches {bad.size}" + (f" (first at {bad[0]})" if bad.size else "") +
          f", fields identical {same_fields}, max relF(G) {e:.2e}, oracle XMAXG {oc['XMAXG']:.2e} XMEANG {oc['XMEANG']/max(oc['NCG'],1):.2e}")
cg = g.control()
print(f"device (max over chains): XMAXG {cg['XMAXG']:.2e} XMEANG {cg['XMEANG']/max(cg['NCG'],1):.2e} XMAXP {cg['XMAXP']:.2e}")
g.close()
