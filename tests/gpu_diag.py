"""First-contact diagnostics on a real GPU: prints numbers instead of asserting (run under gpurun)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from alf_b200 import api
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_square
import oracle.oracle as O
from oracle.oracle import Oracle
from common import relF, SEEDS


def step(name, fn):
    t = time.time()
    try:
        r = fn(); print(f"[diag] {name}: {r}  ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[diag] {name}: EXCEPTION {e!r}", flush=True); traceback.print_exc()


def d_peak():
    return "DFMA %.2f TF, DMMA %.2f TF" % api.fp64_peak(0)


def d_gemm():
    rng = np.random.default_rng(0); out = []
    for (m, n, k) in [(16, 16, 16), (70, 33, 45), (256, 256, 256)]:
        for ta, tb in [(0, 0), (1, 0), (0, 1), (1, 1)]:
            A = rng.normal(size=(2, k, m) if ta else (2, m, k)); B = rng.normal(size=(2, n, k) if tb else (2, k, n))
            C = api.test_gemm(A, B, ta, tb, False)
            oa = A.transpose(0, 2, 1) if ta else A; ob = B.transpose(0, 2, 1) if tb else B
            out.append("%.1e" % relF(C, oa @ ob))
    return out


def d_qr():
    out = []
    for n in (16, 64, 256):
        rng = np.random.default_rng(n); A = rng.normal(size=(2, n, n)) * np.exp(rng.normal(size=(2, 1, n)) * 4)
        QR, D, jp, tau, ph = api.test_qdrp(A, False)
        R = np.triu(QR[0]); Q = np.eye(n)
        for j in range(n):
            v = np.zeros(n); v[j] = 1; v[j + 1:] = QR[0][j + 1:, j].real; Q = Q @ (np.eye(n) - tau[0, j].real * np.outer(v, v))
        _, Dref, ip, _ = O.qdrp(A[0])
        out.append((n, "rec %.1e" % relF(Q @ np.diag(D[0]) @ R, A[0][:, jp[0] - 1]), "D %.1e" % relF(D[0], Dref.real), "piv same %s" % np.array_equal(jp[0], ip)))
    return out


def d_udv():
    out = []
    for n in (16, 256):
        for side in "rl":
            rng = np.random.default_rng(n); U0 = rng.normal(size=(1, n, n)); V0 = np.linalg.qr(rng.normal(size=(n, n)))[0][None]; D0 = np.exp(rng.normal(size=(1, n)) * 5)
            U, D, V = api.test_udv_decompose(U0, D0, V0, side, False)
            B0 = U0[0] @ np.diag(D0[0]) @ (V0[0] if side == "r" else V0[0].T); B1 = U[0] @ np.diag(D[0]) @ (V[0] if side == "r" else V[0].conj().T)
            out.append((n, side, "%.1e" % relF(B1, B0), "orth %.1e" % relF(U[0].conj().T @ U[0], np.eye(n))))
    return out


def d_sweep(L, beta, C, nsw=1, Mz=True):
    model = hubbard_square(L, L, beta, Mz=Mz); seeds = SEEDS[:C]
    g = AlfB200(model, n_chains=C, nwrap=10); g.set_seeds(seeds); g.fields_set()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=10); o.ranset(s); o.fields_set(); orcs.append(o)
    res = ["fields eq %s" % all(np.array_equal(g.get_fields()[c], o.get_fields()) for c, o in enumerate(orcs))]
    t = time.time(); g.init_sweep(); res.append("gpu init %.3fs" % (time.time() - t))
    for o in orcs:
        o.init()
    res.append("G0 relF " + " ".join("%.1e" % relF(g.green(c, nf), o.green(nf)) for c, o in enumerate(orcs) for nf in range(1, model.N_FL + 1)))
    res.append("phase " + " ".join("%.1e" % abs(g.phase()[c] - o.phase()) for c, o in enumerate(orcs)))
    g.accept_log(True)
    for o in orcs:
        o.log(True)
    t = time.time(); g.sweep(nsw, 0); res.append("gpu sweep %.3fs" % (time.time() - t))
    t = time.time()
    for o in orcs:
        for _ in range(nsw):
            o.sweep(0)
    res.append("oracle sweeps %.2fs" % (time.time() - t))
    log = g.get_accept_log()
    for c, o in enumerate(orcs):
        acc, _ = o.get_log(); bad = acc != log[c]
        res.append("chain%d: acc mismatches %d/%d first %s rate %.3f" % (c, bad.sum(), acc.size, np.argmax(bad) if bad.any() else None, (acc == 1).mean()))
    res.append("G relF " + " ".join("%.1e" % relF(g.green(c, nf), o.green(nf)) for c, o in enumerate(orcs) for nf in range(1, model.N_FL + 1)))
    res.append("phase " + " ".join("%.1e" % abs(g.phase()[c] - o.phase()) for c, o in enumerate(orcs)))
    res.append(str({k: v for k, v in g.control().items() if k in ("XMEANG", "XMAXG", "NCG", "XMAXP", "nan")}))
    g.close()
    return res


def d_time(L, beta, C, nsw=2):
    model = hubbard_square(L, L, beta)
    g = AlfB200(model, n_chains=C, nwrap=10); g.set_seeds(np.arange(1, C + 1) * 17); g.fields_set(); g.init_sweep()
    g.sweep(1, 0); t = time.time(); g.sweep(nsw, 0); dt = time.time() - t
    c = g.control(); g.close()
    return "%d chains x %d sweeps: %.3fs -> %.1f sweeps/s; acc %.3f XMEANG %.2e XMAXG %.2e" % (C, nsw, dt, C * nsw / dt, c["ACC_up"] / c["NC_up"], c["XMEANG"] / c["NCG"], c["XMAXG"])


if __name__ == "__main__":
    step("fp64 peak", d_peak)
    step("gemm", d_gemm)
    step("qr", d_qr)
    step("udv", d_udv)
    step("sweep 4x4 beta2", lambda: d_sweep(4, 2.0, 2))
    step("sweep 4x4 beta5 su2", lambda: d_sweep(4, 5.0, 2, Mz=False))
    step("sweep 8x8 beta10", lambda: d_sweep(8, 10.0, 2))
    step("time 8x8 beta10 x256", lambda: d_time(8, 10.0, 256))
    step("time 16x16 beta10 x148", lambda: d_time(16, 10.0, 148, 1))
