"""GPU parity tests (-m gpu): every kernel and the full sweep, through the C-ABI, against the CPU oracle.
Tolerances: 1e-10 relative Frobenius on freshly recomputed G (north_star check 1); accept/reject sequences and field
configurations bit-exact (check 2)."""
import os

import numpy as np
import pytest

from alf_b200 import api
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_square, hubbard_chain, kondo_square, z2_gauge_square, z2_matter_square
import oracle.oracle as O
from oracle.oracle import Oracle
from common import relF, SEEDS, TOL_G, config1, config2, config3

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(16, 16, 16), (64, 64, 64), (70, 33, 45), (256, 256, 256)])
def test_gemm(is_complex, ta, tb, m, n, k):
    rng = np.random.default_rng(0); batch = 3
    def mk(r, c):
        x = rng.normal(size=(batch, r, c)); return x + (1j * rng.normal(size=(batch, r, c)) if is_complex else 0)
    A = mk(k, m) if ta else mk(m, k); B = mk(n, k) if tb else mk(k, n)
    Cg = api.test_gemm(A, B, ta, tb, is_complex)
    opA = A.conj().transpose(0, 2, 1) if ta else A; opB = B.conj().transpose(0, 2, 1) if tb else B
    assert relF(Cg, opA @ opB) < 1e-14


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("n", [5, 16, 50, 64, 100, 256])
def test_qdrp_reconstruct(is_complex, n):
    """A P = Q D R with unit-modulus diag(R) (20-qdrp.F90), pivots = descending |R_ii| D, and same D as LAPACK's ZGEQP3."""
    rng = np.random.default_rng(n); batch = 2
    A = rng.normal(size=(batch, n, n)) * np.exp(rng.normal(size=(batch, 1, n)) * 4)
    if is_complex:
        A = A + 1j * rng.normal(size=(batch, n, n))
    if n == 50:
        A[0] = np.array([[1.0 / (i + j + 1) for j in range(n)] for i in range(n)])
    QR, D, jp, tau, ph = api.test_qdrp(A, is_complex)
    for b in range(batch):
        R = np.triu(QR[b]); Q = np.eye(n, dtype=complex)
        for j in range(n):
            v = np.zeros(n, complex); v[j] = 1; v[j + 1:] = QR[b][j + 1:, j]
            Q = Q @ (np.eye(n) - tau[b, j] * np.outer(v, v.conj()))
        assert relF(Q @ np.diag(D[b]) @ R, A[b][:, jp[b] - 1]) < 1e-12
        assert relF(Q.conj().T @ Q, np.eye(n)) < 1e-12
        assert np.all(np.abs(np.abs(np.diag(R)) - 1) < 1e-12)
        _, Dref, _, _ = O.qdrp(A[b])
        if n != 50:
            assert relF(D[b], Dref.real) < 1e-8
        sign = np.linalg.det(np.eye(n)[:, jp[b] - 1])
        assert abs(sign - ph[b, 0]) < 1e-12
        assert abs(np.linalg.det(Q) - complex(ph[b, 3], ph[b, 4])) < 1e-9


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("side", ["r", "l"])
@pytest.mark.parametrize("n", [16, 64, 256])
def test_udv_decompose(is_complex, side, n):
    """decompose: the product U D V (side r) / U D V^H (side l) is unchanged and matches the oracle's factorisation product."""
    rng = np.random.default_rng(3 + n); batch = 2
    U0 = rng.normal(size=(batch, n, n)) + (1j * rng.normal(size=(batch, n, n)) if is_complex else 0)
    V0 = np.stack([np.linalg.qr(rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if is_complex else 0))[0] for _ in range(batch)])
    D0 = np.exp(rng.normal(size=(batch, n)) * 5)
    U, D, V = api.test_udv_decompose(U0, D0, V0, side, is_complex)
    for b in range(batch):
        B0 = U0[b] @ np.diag(D0[b]) @ (V0[b] if side == "r" else V0[b].conj().T)
        B1 = U[b] @ np.diag(D[b]) @ (V[b] if side == "r" else V[b].conj().T)
        assert relF(B1, B0) < 1e-11
        assert relF(U[b].conj().T @ U[b], np.eye(n)) < 1e-12
        Uo, Do, Vo = O.udv_decompose(U0[b], D0[b], V0[b], side)
        assert relF(D[b].real, Do.real) < 1e-8
        assert abs(np.linalg.det(V[b]) / np.linalg.det(Vo) - 1) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("shape", [(16, 16), (24, 9), (64, 64), (144, 72), (256, 256)])
def test_udv_wrap_pivot(is_complex, shape):
    """UDV_Wrap_Pivot (Prog/UDV_WRAP_mod.F90:125-208): U, D, V equal the oracle's (same norm sort, same Householder convention) and
    satisfy the routine's contract A = U D V, U^H U = 1, D > 0, det V = 1 -- on graded products as in testsuite/Prog.tests/24-udv.F90."""
    n1, n2 = shape; rng = np.random.default_rng(17 + n1); batch = 3
    A = np.zeros((batch, n1, n2), dtype=np.complex128)
    for b in range(batch):
        X = np.eye(n1, n2, dtype=np.complex128)
        for _ in range(3 + b):
            X = ((4 * (rng.random((n1, n1)) - 0.5) + (2j * (rng.random((n1, n1)) - 0.5) if is_complex else 0)) / np.sqrt(n1)) @ X
        A[b] = X * np.exp(np.linspace(-6, 6, n2))[rng.permutation(n2)][None, :]
    U, D, V = api.udv_wrap_pivot(A, is_complex)
    for b in range(batch):
        assert relF(U[b] @ np.diag(D[b]) @ V[b], A[b]) < 1e-12
        assert relF(U[b].conj().T @ U[b], np.eye(n2)) < 1e-12
        assert np.all(D[b].real > 0)
        sd, ld = np.linalg.slogdet(V[b]); assert abs(sd - 1) < 1e-8 and abs(ld) < 1e-8
        Uo, Do, Vo = O.udv_wrap_pivot(A[b])
        assert relF(D[b], Do) < 1e-10 and relF(U[b], Uo) < 1e-9 and relF(V[b], Vo) < 1e-9


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("nvar", [1, 2])
@pytest.mark.parametrize("stab", [0, 3])
@pytest.mark.parametrize("n", [5, 16, 64])
def test_cgr_kernel(is_complex, nvar, stab, n):
    """CGR on UDV states with 20 orders of magnitude of scales: G and phase vs the oracle (15-cgr.F90 tolerances)."""
    rng = np.random.default_rng(100 + n); batch = 2
    UR, DR, VR, UL, DL, VL, dR, dL = [], [], [], [], [], [], [], []
    for b in range(batch):
        def mk(side):
            U0 = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if is_complex else 0)
            U0 = U0 * np.exp(np.linspace(-10, 10, n))[None, :]
            return O.udv_decompose(U0, np.ones(n), np.eye(n), side)
        a = mk("r"); c = mk("l")
        UR.append(a[0]); DR.append(a[1]); VR.append(a[2]); UL.append(c[0]); DL.append(c[1]); VL.append(c[2])
        dR.append(np.linalg.det(a[0])); dL.append(np.linalg.det(c[0]))
    st = lambda x: np.stack(x)
    G, ph = api.test_cgr(st(UR), st(DR), st(VR), st(UL), st(DL), st(VL), st(dR), st(dL), nvar, stab, is_complex)
    for b in range(batch):
        Go, pho = O.cgr(UR[b], DR[b], VR[b], UL[b], DL[b], VL[b], nvar=nvar, stab3=(stab == 3))
        assert relF(G[b], Go) < TOL_G
        assert abs(ph[b] - pho) < 1e-9


def _hop_check(model, nf=1):
    g = AlfB200(model, n_chains=2, nwrap=5)
    o = Oracle(model, nwrap=5)
    rng = np.random.default_rng(9)
    A = rng.normal(size=(model.Ndim, model.Ndim)) + (1j * rng.normal(size=(model.Ndim, model.Ndim)) if g.is_complex else 0)
    for which in range(6):
        assert relF(g.hop_apply(which, nf, A), o.hop_apply(which, nf, A)) < 1e-13, which
    g.close()


def test_hop_checkerboard_symm():
    _hop_check(config1())


def test_hop_dense():
    _hop_check(hubbard_square(4, 4, 1.0, checkerboard=False, symm=False))


def test_hop_golden_26_path():
    """The golden-vector operator (testsuite/Prog.tests/26-...) family goes through the dense-T GEMM path on the GPU."""
    _hop_check(hubbard_chain(4, 1.0, 0.1, Mz=False, symm=True))


def _run_parity(model, seeds, nwrap, n_sweeps=1, stab=0, check_udv=True, xmaxg=1e-6):
    C = len(seeds)
    g = AlfB200(model, n_chains=C, nwrap=nwrap, stab=stab)
    g.set_seeds(seeds); g.fields_set()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=nwrap, stab3=(stab == 3)); o.ranset(s); o.fields_set(); orcs.append(o)
    f = g.get_fields()
    for c, o in enumerate(orcs):
        assert np.array_equal(f[c], o.get_fields()), "Fields_set stream differs"
    g.init_sweep()
    for o in orcs:
        o.init()
    ph = g.phase()
    for c, o in enumerate(orcs):
        for nf in range(1, model.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G, ("init", c, nf)
        assert abs(ph[c] - o.phase()) < 1e-9
        if check_udv:
            U, D, V = g.get_udv(0, 0, c, 1); Uo, Do, Vo = o.get_udv(0, 0, 1)
            assert relF(D.real, Do.real) < 1e-7        # scales of the stored left propagation
    g.accept_log(n_sweeps)
    for o in orcs:
        o.log(True)
    for sw in range(n_sweeps):
        g.sweep(1, 0)
        for o in orcs:
            o.sweep(0)
    log = g.get_accept_log(); f = g.get_fields(); ph = g.phase(); st = g.rng_state()
    for c, o in enumerate(orcs):
        acc, _ = o.get_log()
        assert log.shape[1] == acc.size
        nbad = int(np.sum(acc != log[c]))
        assert nbad == 0, f"chain {c}: {nbad} of {acc.size} accept/reject decisions differ (first at {np.argmax(acc != log[c])})"
        assert np.array_equal(f[c], o.get_fields())
        assert np.array_equal(st[c], o.rng_state())
        for nf in range(1, model.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G, ("sweep", c, nf)
        assert abs(ph[c] - o.phase()) < 1e-9
    # device-side scalar observables (N_meas, <sign>, particle number) accumulated where main.F90 calls ham%Obser
    ob = g.obs()[:4]; oo = sum(o.obs() for o in orcs)
    assert ob[0] == oo[0] and ob[1] == oo[1] and ob[0] == len(seeds) * n_sweeps * (2 * model.Ltrot - 1)
    assert abs(ob[2] - oo[2]) <= 1e-9 * abs(oo[2]) + 1e-9 and abs(ob[3] - oo[3]) <= 1e-8
    cg = g.control(); tot = {k: 0.0 for k in ("NC_up", "ACC_up", "NCG")}
    for o in orcs:
        co = o.control()
        for k in tot:
            tot[k] += co[k]
    assert cg["NC_up"] == tot["NC_up"] and cg["ACC_up"] == tot["ACC_up"] and cg["NCG"] == tot["NCG"]
    assert cg["XMAXG"] < xmaxg and cg["nan"] == 0 and cg["unstable"] == 0
    g.close()
    return cg


def test_sweep_config1_real_mz():
    """configs[0]: Hubbard 4x4 Mz, U=4, beta=5, dtau=0.1, Nwrap=10."""
    _run_parity(config1(), SEEDS[:4], nwrap=10, n_sweeps=2)


def test_sweep_config1_stab3():
    _run_parity(config1(), SEEDS[:2], nwrap=10, n_sweeps=1, stab=3)


def test_sweep_config1_su2_complex():
    """configs[0], SU(2) variant: N_FL=1, imaginary coupling -> complex instantiation, non-trivial Op_phase."""
    _run_parity(config1(Mz=False), SEEDS[:3], nwrap=10, n_sweeps=1)


def test_sweep_dense_hopping_nonsymm():
    _run_parity(hubbard_square(4, 4, 2.0, checkerboard=False, symm=False), SEEDS[:2], nwrap=5, n_sweeps=1)


def test_sweep_ragged_stabilisation():
    """Ltrot not a multiple of Nwrap (last interval shorter) and Nwrap > Ltrot (single interval)."""
    _run_parity(hubbard_square(4, 4, 2.3, symm=False), SEEDS[:2], nwrap=7, n_sweeps=1)
    _run_parity(hubbard_square(2, 2, 0.5), SEEDS[:2], nwrap=10, n_sweeps=1)


def test_sweep_config2():
    """configs[1]: Hubbard 8x8, beta=10 (N_dim=64), a few of the 256 chains."""
    _run_parity(config2(), SEEDS[:3], nwrap=10, n_sweeps=1)


def test_sweep_kondo_complex_k2():
    """configs[3] family at test size: Kondo 4x4 bilayer (N_dim=32), rank-1 + rank-2 non-diagonal vertices, complex."""
    _run_parity(kondo_square(4, 4, 2.0), SEEDS[:2], nwrap=5, n_sweeps=1, check_udv=False)


def test_init_config3():
    """configs[2]: Hubbard 16x16 beta=10 (N_dim=256): storage fill + CGR at tau=0 for 2 chains."""
    model = config3(); seeds = SEEDS[:2]
    g = AlfB200(model, n_chains=2, nwrap=10); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
    ph = g.phase()
    for c, s in enumerate(seeds):
        o = Oracle(model, nwrap=10); o.ranset(s); o.fields_set(); o.init()
        for nf in (1, 2):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
    g.close()


def test_many_chains_are_independent():
    """Chains are independent Markov chains: a batch of 40 gives the same per-chain results as batches of 1."""
    model = hubbard_square(4, 4, 1.0); seeds = [1000 + 7 * i for i in range(40)]
    g = AlfB200(model, n_chains=40, nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    f = g.get_fields()
    for c in (0, 17, 39):
        g1 = AlfB200(model, n_chains=1, nwrap=5); g1.set_seeds([seeds[c]]); g1.fields_set(); g1.init_sweep(); g1.sweep(1, 0)
        assert np.array_equal(g1.get_fields()[0], f[c])
        assert relF(g1.green(0, 1), g.green(c, 1)) < 1e-13
        g1.close()
    g.close()


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("stab", [0, 3])
@pytest.mark.parametrize("n", [5, 16, 64])
def test_cgr2_2_kernel(is_complex, stab, n):
    """CGR2_2 (Prog/cgr2_2_mod.F90:196) on UDV states with 20 orders of magnitude of scales vs the oracle; both block orderings
    (D1(1) > D2(1) and the opposite) occur in the batch."""
    rng = np.random.default_rng(200 + n); batch = 3
    S2, S1 = [], []
    for b in range(batch):
        def mk(side, spread):
            U0 = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if is_complex else 0)
            U0 = U0 * np.exp(np.linspace(-spread, spread, n))[None, :]
            return O.udv_decompose(U0, np.ones(n), np.eye(n), side)
        S2.append(mk("r", 10 if b != 1 else 4)); S1.append(mk("l", 10 if b == 1 else 6))
    st = lambda k, S: np.stack([s[k] for s in S])
    out = api.test_cgr2_2(st(0, S2), st(1, S2), st(2, S2), st(0, S1), st(1, S1), st(2, S1), stab, is_complex)
    orders = set()
    for b in range(batch):
        ref = O.cgr2_2(S2[b][0], S2[b][1], S2[b][2], S1[b][0], S1[b][1], S1[b][2], stab3=(stab == 3))
        orders.add(bool(S1[b][1][0].real > S2[b][1][0].real))
        for k in ("GRT0", "GR00", "GRTT", "GR0T"):
            assert relF(out[k][b], ref[k]) < TOL_G, (k, b)
    assert len(orders) == 2


def _run_taum(model, seeds, nwrap, every=1):
    C = len(seeds)
    g = AlfB200(model, n_chains=C, nwrap=nwrap); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
    g.taum_capture(every); g.sweep(1, 1)
    cg = g.control()
    for c, s in enumerate(seeds):
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init(); o.taum_capture(every); o.sweep(1)
        a = g.get_taum(c); b = o.taum_get()
        assert a.shape == b.shape and a.shape[0] >= 2
        # freshly recomputed G(tau,0), G(0,tau), G(0,0), G(tau,tau) after every CGR2_2: north_star check (1) at 1e-10
        af = g.get_taum_fresh(c); bf = o.taum_fresh_get()
        assert af.shape == bf.shape and af.shape[0] == o.nstm()
        for ist in range(af.shape[0]):
            for w in range(4):
                for nf in range(model.N_FL):
                    assert relF(af[ist, w, nf], bf[ist, w, nf]) < TOL_G, (c, ist, w, nf, relF(af[ist, w, nf], bf[ist, w, nf]))
        # what ObserT receives at every slice (wrapped in between: drifts like the reference's own "Precision tau", monitored not matched)
        for it in range(a.shape[0]):
            for w in range(4):
                for nf in range(model.N_FL):
                    assert relF(a[it, w, nf], b[it, w, nf]) < 1e-7, (c, it, w, nf, relF(a[it, w, nf], b[it, w, nf]))
        assert np.array_equal(g.get_fields()[c], o.get_fields())
        co = o.control()
        assert co["NCG_tau"] * C == cg["NCG_tau"]
    assert cg["XMAX_tau"] < 1e-5
    g.close()


def test_taum_config1():
    """configs[0] with Ltau=1: G(tau,0), G(0,tau), G(0,0), G(tau,tau) for every tau vs the oracle's TAU_M."""
    _run_taum(config1(), SEEDS[:3], nwrap=10)


def test_taum_su2_complex():
    _run_taum(config1(Mz=False), SEEDS[:2], nwrap=10, every=5)


def test_taum_ragged():
    _run_taum(hubbard_square(4, 4, 2.3, symm=False), SEEDS[:2], nwrap=7)


def test_taum_config2():
    """configs[1] (N_dim = 64, beta = 10): 2N = 128 extended system."""
    _run_taum(config2(), SEEDS[:2], nwrap=10, every=10)


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("n", [40, 64, 100, 256, 512])
def test_qdrp_blocked_reconstruct(is_complex, n):
    """Blocked windowed-pivoting QR (alf_qrblk.cuh): A P = Q D R with Q from the compact-WY application, R upper triangular with
    unit-modulus diagonal, graded D (same scales as ZGEQP3 within a factor), permutation parity and det(Q) bookkeeping."""
    if is_complex and n == 512:
        pytest.skip("2N = 512 complex exceeds the test's time budget")
    rng = np.random.default_rng(n); batch = 2
    A = rng.normal(size=(batch, n, n)) * np.exp(rng.normal(size=(batch, 1, n)) * 6)
    if is_complex:
        A = A + 1j * rng.normal(size=(batch, n, n)) * np.exp(rng.normal(size=(batch, 1, n)) * 6)
    QR, D, jp, tau, ph, Q = api.test_qdrp_blocked(A, is_complex)
    for b in range(batch):
        R = np.triu(QR[b])
        assert relF(Q[b] @ np.diag(D[b]) @ R, A[b][:, jp[b] - 1]) < 1e-12
        assert relF(Q[b].conj().T @ Q[b], np.eye(n)) < 1e-12
        assert np.all(np.abs(np.abs(np.diag(R)) - 1) < 1e-12)
        assert sorted(jp[b]) == list(range(1, n + 1))
        _, Dref, _, _ = O.qdrp(A[b])
        assert np.max(np.abs(np.log(D[b] / Dref.real))) < np.log(50.0)       # same grading; the pivot order may differ
        sign = np.linalg.det(np.eye(n)[:, jp[b] - 1])
        assert abs(sign - ph[b, 0]) < 1e-12
        assert abs(np.linalg.det(Q[b]) - complex(ph[b, 3], ph[b, 4])) < 1e-8


# ---------------------------------------------------------------------------------- projective algorithm (CGRP, Tau_p)
@pytest.mark.gpu
@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("n,npart", [(5, 2), (16, 8), (64, 32), (256, 128)])
def test_cgrp_kernel(is_complex, n, npart):
    """CGRP on the device (QR-based inverse instead of LU) against the oracle's ZGETRF/ZGETRS restatement (testsuite 23-cgrp)."""
    from oracle.oracle import cgrp
    rng = np.random.default_rng(1000 + n)
    UR = rng.normal(size=(2, n, npart)); UL = rng.normal(size=(2, n, npart))
    if is_complex:
        UR = UR + 1j * rng.normal(size=UR.shape); UL = UL + 1j * rng.normal(size=UL.shape)
    UR = np.linalg.qr(UR)[0]; UL = np.linalg.qr(UL)[0]       # column-orthonormal, as decompose leaves them
    G, ph = api.test_cgrp(UR, UL, is_complex)
    for b in range(2):
        Go, pho = cgrp(UR[b], UL[b])
        assert relF(G[b], Go) < TOL_G
        assert abs(ph[b] - pho) < 1e-9


def _run_projector(model, seeds, nwrap, ltau):
    C = len(seeds)
    g = AlfB200(model, n_chains=C, nwrap=nwrap); g.set_seeds(seeds); g.fields_set()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set(); orcs.append(o)
    g.init_sweep()
    for o in orcs:
        o.init()
    ph = g.phase()
    for c, o in enumerate(orcs):
        for nf in range(1, model.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G, ("init", c, nf)
        assert abs(ph[c] - o.phase()) < 1e-9
    g.accept_log(1)
    if ltau:
        g.taum_capture(1)
    for o in orcs:
        o.log(True)
        if ltau:
            o.taum_capture(1)
    g.sweep(1, ltau)
    for o in orcs:
        o.sweep(ltau)
    log = g.get_accept_log(); f = g.get_fields(); ph = g.phase()
    for c, o in enumerate(orcs):
        acc, _ = o.get_log()
        assert np.array_equal(acc, log[c]), f"chain {c}: accept/reject sequence differs"
        assert np.array_equal(f[c], o.get_fields())
        for nf in range(1, model.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G, ("sweep", c, nf)
        assert abs(ph[c] - o.phase()) < 1e-9
        if ltau:
            a, b = g.get_taum(c), o.taum_get()
            assert a.shape == b.shape and a.shape[0] == model.Ltrot - 2 * model.Thtrot + 1
            assert relF(a, b) < 1e-8                      # propagated (wrapped) time-displaced blocks handed to ObserT
            af, bf = g.get_taum_fresh(c), o.taum_fresh_get()
            assert af.shape == bf.shape
            if af.size:
                assert relF(af[:, 3], bf[:, 3]) < TOL_G   # G(tau,tau) freshly recomputed by CGRP
                assert relF(af, bf) < 1e-8
    ob = g.obs()[:4]; oo = sum(o.obs() for o in orcs)      # measured only inside [Thtrot+1, Ltrot-Thtrot]
    assert ob[0] == oo[0] and ob[1] == oo[1] and ob[0] > 0 and abs(ob[2] - oo[2]) <= 1e-9 * abs(oo[2]) + 1e-9
    cg = g.control()
    assert cg["XMAXG"] < 1e-6 and cg["nan"] == 0 and cg["unstable"] == 0
    if ltau:
        assert cg["NCG_tau"] == sum(o.control()["NCG_tau"] for o in orcs) and cg["XMAX_tau"] < 1e-6
    g.close()


@pytest.mark.gpu
def test_projector_sweep_complex_trial():
    """Hubbard 4x4, projective algorithm with ALF's flux-twisted trial wave function (complex arithmetic)."""
    _run_projector(hubbard_square(4, 4, 1.0, 0.1, 4.0, projector=True, theta=0.5), SEEDS[:2], nwrap=5, ltau=0)


@pytest.mark.gpu
def test_projector_sweep_real_trial_taup():
    """Real trial wave function (real instantiation), Tau_p called inside the down sweep (Stab_nt(NST) <= Thtrot+1 < Stab_nt(NST+1))."""
    _run_projector(hubbard_square(4, 4, 1.0, 0.1, 4.0, projector=True, theta=0.6, trial="dimer"), SEEDS[:2], nwrap=4, ltau=1)


@pytest.mark.gpu
def test_projector_taup_after_sweep_complex():
    """Nwrap > Thtrot+1: Tau_p runs after the sweep with NST_IN = 0 (main.F90:884-886); SU(2) flavor-symmetric model, complex."""
    _run_projector(hubbard_square(4, 4, 0.8, 0.1, 4.0, Mz=False, projector=True, theta=0.3), SEEDS[:2], nwrap=6, ltau=1)


@pytest.mark.gpu
def test_projector_config_like_8x8():
    """8x8 Hubbard projector (N_dim = 64, N_part = 32), the size class of BASELINE configs[4]'s projective runs."""
    _run_projector(hubbard_square(8, 8, 1.0, 0.1, 4.0, projector=True, theta=1.0, trial="dimer"), SEEDS[:2], nwrap=10, ltau=1)


# ---------------------------------------------------------------------------------- global-in-slice moves (a3)
def _other_value(model, n, cur, rng):
    """A new field value for operator n (1-based) different from the current one: type 1 -> flip, type 2 -> one of the other three."""
    cur = int(round(cur.real))
    if model.Op_V[n - 1][0].type == 1:
        return complex(-cur)
    return complex([x for x in (-2, -1, 1, 2) if x != cur][int(rng.integers(0, 3))])


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["hubbard_mz", "hubbard_su2", "kondo"])
def test_wrapgr_random_update_and_placegr(which):
    """Wrapgr_Random_update / Wrapgr_PlaceGR (Prog/Wrapgr_mod.F90:247-433) with host-supplied proposals: single- and multi-field
    moves, accepted and rejected (rollback of G and fields), T0_Proposal_ratio = 0 (not proposed), then PlaceGR back to size(Op_V,1)."""
    model = {"hubbard_mz": lambda: hubbard_square(4, 4, 1.0), "hubbard_su2": lambda: hubbard_square(4, 4, 1.0, Mz=False),
             "kondo": lambda: kondo_square(2, 2, 1.0)}[which]()
    seeds = SEEDS[:3]; C = len(seeds); M = model.n_opv; ntau = 1; nmoves = 6; maxlen = 4
    g = AlfB200(model, n_chains=C, nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=5); o.ranset(s); o.fields_set(); o.init(); orcs.append(o)
    g.wrapgrup(0)
    for o in orcs:
        o.wrapgrup(0)
    g.wrapgr_set_position(M)
    rng = np.random.default_rng(7)
    flen = np.zeros((C, nmoves), dtype=np.int32); flist = np.ones((C, nmoves, maxlen), dtype=np.int32)
    fval = np.ones((C, nmoves, maxlen), dtype=np.complex128); t0 = np.ones((C, nmoves)); s0 = np.ones((C, nmoves))
    acc_ref = np.zeros((C, nmoves), dtype=np.uint8); m_ref = np.zeros(C, dtype=np.int32)
    for c, o in enumerate(orcs):
        m = M
        for mv in range(nmoves):
            L = int(rng.integers(1, maxlen + 1)); ns = rng.choice(np.arange(1, M + 1), size=L, replace=False)
            f = o.get_fields()[ntau - 1]
            vals = [_other_value(model, int(n), f[n - 1], rng) for n in ns]
            flen[c, mv] = L; flist[c, mv, :L] = ns; fval[c, mv, :L] = vals
            t0[c, mv] = 0.0 if mv == 2 else float(rng.uniform(0.5, 2.0)); s0[c, mv] = float(rng.uniform(0.5, 2.0))
            if t0[c, mv] > 1e-7:
                a, m = o.wrapgr_random_update(m, ntau, t0[c, mv], s0[c, mv], ns, vals)
                acc_ref[c, mv] = 1 if a else 0
            else:
                acc_ref[c, mv] = 2
        o.wrapgr_placegr(m, M, ntau); m_ref[c] = M
    # NOTE: the oracle consumed the proposals sequentially with fields evolving; the device gets the same lists
    acc = g.wrapgr_random_update(ntau, flen, flist, fval, t0, s0, place_to=M)
    assert np.array_equal(acc, acc_ref), (acc, acc_ref)
    assert 0 < int(np.sum(acc_ref == 1)) and int(np.sum(acc_ref == 0)) > 0          # both branches exercised
    assert np.array_equal(g.wrapgr_get_position(), m_ref)
    f = g.get_fields(); ph = g.phase(); st = g.rng_state()
    for c, o in enumerate(orcs):
        assert np.array_equal(f[c], o.get_fields())
        assert np.array_equal(st[c], o.rng_state())
        for nf in range(1, model.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G, (c, nf)
        assert abs(ph[c] - o.phase()) < 1e-9
    # moving GR down to position 0 and back up is the identity on G
    g.wrapgr_placegr(0, ntau); g.wrapgr_placegr(M, ntau)
    for c, o in enumerate(orcs):
        assert relF(g.green(c, 1), o.green(1)) < TOL_G
    g.close()


# ---------------------------------------------------------------------------------- north-star check (3): physics
@pytest.mark.gpu
def test_ed_energy_two_site_gpu():
    """Full-run observable against exact diagonalisation (the spirit of testsuite/test_vs_ed): 2-site Hubbard, U = 4, beta = 2,
    64 independent chains x 120 sweeps on the device; energy from the equal-time G of every chain after every sweep.
    Also: the device chains are statistically consistent with the oracle's chains (same estimator, different seeds)."""
    t, U, beta, dtau = 1.0, 4.0, 2.0, 0.05

    def cdag(i, n=4):
        dim = 2 ** n; M = np.zeros((dim, dim))
        for s in range(dim):
            if not (s >> i) & 1:
                M[s | (1 << i), s] = (-1) ** bin(s & ((1 << i) - 1)).count("1")
        return M
    cdg = [cdag(i) for i in range(4)]; cc = [x.T for x in cdg]; nn = [cdg[i] @ cc[i] for i in range(4)]
    H = -t * (cdg[0] @ cc[1] + cdg[1] @ cc[0] + cdg[2] @ cc[3] + cdg[3] @ cc[2])
    H = H + U * ((nn[0] - 0.5 * np.eye(16)) @ (nn[2] - 0.5 * np.eye(16)) + (nn[1] - 0.5 * np.eye(16)) @ (nn[3] - 0.5 * np.eye(16)))
    w = np.linalg.eigvalsh(H); Z = np.exp(-beta * w); E_ed = float((w * Z).sum() / Z.sum())
    m = hubbard_chain(2, beta, dtau, U=U, t=t, Mz=True, symm=False)
    C = 64
    g = AlfB200(m, n_chains=C, nwrap=10); g.set_seeds([1000 + 17 * c for c in range(C)]); g.fields_set(); g.init_sweep()
    Tm = np.array([[0, -t], [-t, 0]], float)

    def energy(Gu, Gd):
        kin = np.sum(Tm * ((np.eye(2) - Gu).T + (np.eye(2) - Gd).T))
        return kin + U * np.sum((1 - np.diag(Gu) - 0.5) * (1 - np.diag(Gd) - 0.5))
    g.sweep(20, 0)                                   # warm-up
    per_chain = np.zeros(C); nsw = 100
    for sw in range(nsw):
        g.sweep(1, 0)
        for c in range(C):
            per_chain[c] += energy(g.green(c, 1).real, g.green(c, 2).real) / nsw
    mean = per_chain.mean(); err = per_chain.std(ddof=1) / np.sqrt(C)      # chains are independent: error from the chain-to-chain spread
    assert abs(mean - E_ed) < max(4 * err, 0.01), (mean, err, E_ed)        # Trotter error O(dtau^2 U t^2) ~ 1e-2 is inside
    cg = g.control()
    assert cg["nan"] == 0 and cg["unstable"] == 0 and cg["XMAXG"] < 1e-8
    ob = g.obs()
    assert abs(ob[2] / ob[0] - 2.0) < 0.02           # half filling: <N> = 2 (particle-hole symmetry), accumulated on the device over all slices
    g.close()


@pytest.mark.gpu
def test_first_sweep_config3_baseline_size():
    """BASELINE configs[2] size (Hubbard 16x16, beta = 10, N_dim = 256, 51200 decisions per chain-sweep): the accept/reject sequence of
    the first sweep is identical to the oracle's, the fields agree bit for bit, the freshly recomputed G agrees to 1e-10 and the
    precision monitors (Control_PrecisionG) are of the same size as the oracle's (the windowed pivoting costs no accuracy)."""
    import threading
    model = config3(); seeds = SEEDS[:2]; C = len(seeds)
    g = AlfB200(model, n_chains=C, nwrap=10); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.accept_log(1)
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=10); o.ranset(s); o.fields_set(); orcs.append(o)

    def work(o):
        o.init(); o.log(True); o.sweep(0)
    th = [threading.Thread(target=work, args=(o,)) for o in orcs]
    [t.start() for t in th]; [t.join() for t in th]
    g.sweep(1, 0)
    log = g.get_accept_log(); f = g.get_fields(); ph = g.phase(); xmax_o = 0.0
    for c, o in enumerate(orcs):
        acc, _ = o.get_log()
        assert acc.size == 2 * model.Ltrot * model.n_opv == log.shape[1]
        assert np.array_equal(acc, log[c]), f"chain {c}: first mismatch at decision {int(np.argmax(acc != log[c]))}"
        assert np.array_equal(f[c], o.get_fields())
        for nf in (1, 2):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
        xmax_o = max(xmax_o, o.control()["XMAXG"])
    cg = g.control()
    assert cg["XMAXG"] < 10 * xmax_o + 1e-9 and cg["nan"] == 0 and cg["unstable"] == 0
    g.close()


# ---------------------------------------------------------------------------------- slice-kernel variants
@pytest.mark.gpu
def test_sweep_k2_vertex_with_one_nonzero_eigenvalue():
    """Pair mode of the fast slice kernel with N_non_zero = 1 < k = 2 (second rank-1 factor vanishes)."""
    m = kondo_square(2, 2, 1.0, jk_vertex="bond_density")
    assert any(op[0].N == 2 and op[0].N_non_zero == 1 for op in m.Op_V)
    _run_parity(m, SEEDS[:2], nwrap=5, n_sweeps=1, check_udv=False)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["hubbard", "kondo"])
def test_sweep_generic_slice_kernel(which, monkeypatch):
    """ALF_B200_GENERIC_UPDATE=1 routes every vertex through the per-visit kernel k_wrapgr (the path of k > 2 vertices)."""
    monkeypatch.setenv("ALF_B200_GENERIC_UPDATE", "1")
    model = hubbard_square(4, 4, 1.0) if which == "hubbard" else kondo_square(2, 2, 1.0)
    _run_parity(model, SEEDS[:2], nwrap=5, n_sweeps=1, check_udv=False)


@pytest.mark.gpu
def test_free_fermions_no_vertices():
    """Edge case: U = 0 leaves size(Op_V,1) = 0 (empty field loop).  G(0) must be the free-fermion (1 + prod e^{-dtau T})^-1 exactly and a
    sweep (with TAU_M) must leave it unchanged; also the oracle agrees."""
    import scipy.linalg as sl
    model = hubbard_square(4, 4, 1.0, U=0.0, checkerboard=False, symm=False)
    assert model.n_opv == 0
    g = AlfB200(model, n_chains=2, nwrap=5); g.set_seeds([1, 2]); g.fields_set(); g.init_sweep()
    op = model.Op_T[0][0]
    Tm = np.zeros((model.Ndim, model.Ndim), dtype=complex)
    for a in range(op.N):
        for b in range(op.N):
            Tm[op.P[a] - 1, op.P[b] - 1] = op.O[a, b]
    B = np.linalg.matrix_power(sl.expm(op.g * Tm), model.Ltrot)
    Gex = np.linalg.inv(np.eye(model.Ndim) + B)
    for c in range(2):
        for nf in (1, 2):
            assert relF(g.green(c, nf), Gex) < 1e-12
    g.sweep(1, 1)
    for c in range(2):
        assert relF(g.green(c, 1), Gex) < 1e-12
    cg = g.control()
    assert cg["NC_up"] == 0 and cg["XMAXG"] < 1e-12 and cg["XMAX_tau"] < 1e-10 and cg["nan"] == 0
    g.close()


@pytest.mark.gpu
def test_too_large_ndim_fails_loudly():
    """Sizes outside what the kernels support are refused at finalize_model with an error code and a message (no silent fallback)."""
    from alf_b200.api import AlfError
    with pytest.raises(AlfError) as ei:
        AlfB200(hubbard_square(26, 24, 0.2, U=0.0, checkerboard=True), n_chains=1, nwrap=2)
    assert "not supported" in str(ei.value)


# ---------------------------------------------------------------------------------- device-side ObserT (time-displaced lattice observables)
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["hubbard_mz_symm", "hubbard_su2", "hubbard_mz_nosymm", "kondo", "projector", "hubbard_mz_8x8_diag", "hubbard_su2_8x8_diag", "kondo_8x4_diag"])
def test_obs_tau_on_device(which):
    """Green / SpinZ / SpinXY / Den time-displaced correlation functions and their backgrounds, accumulated on the device where TAU_M /
    Tau_p call ham%ObserT (with Hop_mod_Symm when Symm), against the oracle's restatement of Predefined_Obs_tau_*_measure."""
    model = {"hubbard_mz_symm": lambda: hubbard_square(4, 4, 0.8), "hubbard_su2": lambda: hubbard_square(4, 4, 0.8, Mz=False),
             "hubbard_mz_nosymm": lambda: hubbard_square(4, 2, 0.6, symm=False), "kondo": lambda: kondo_square(2, 2, 0.6),
             "projector": lambda: hubbard_square(4, 4, 0.6, projector=True, theta=0.3, trial="dimer"),
             # N a multiple of 32 and the numbering invariant under a shift by 32 sites: the register-accumulating kernel k_obs_tau_diag
             "hubbard_mz_8x8_diag": lambda: hubbard_square(8, 8, 0.4), "hubbard_su2_8x8_diag": lambda: hubbard_square(8, 8, 0.4, Mz=False),
             "kondo_8x4_diag": lambda: kondo_square(8, 4, 0.4)}[which]()
    seeds = SEEDS[:2]; nwrap = 4
    g = AlfB200(model, n_chains=len(seeds), nwrap=nwrap); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.obs_tau_enable()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init(); o.obs_tau_enable(); orcs.append(o)
    for sw in range(2):
        g.sweep(1, 1)
        for o in orcs:
            o.sweep(1)
    acc, bg, n, sg = g.obs_tau()
    ro = [o.obs_tau() for o in orcs]
    acc_o = sum(r[0] for r in ro); bg_o = sum(r[1] for r in ro); n_o = sum(r[2] for r in ro); sg_o = sum(r[3] for r in ro)
    assert n == n_o == 2 * len(seeds) and sg == sg_o
    assert acc.shape == acc_o.shape and acc.shape[1] == model.Ltrot - 2 * model.Thtrot + 1
    for ch in range(4):
        if np.abs(acc_o[ch]).max() > 0:
            assert relF(acc[ch], acc_o[ch]) < 1e-9, ch
        else:
            assert np.abs(acc[ch]).max() == 0
    assert np.abs(bg - bg_o).max() < 1e-9 * max(1.0, np.abs(bg_o).max())
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["hubbard_mz_symm", "hubbard_su2", "kondo", "hubbard_mz_8x8_diag", "kondo_8x4_diag"])
def test_obs_eq_on_device(which):
    """Equal-time Green / SpinZ / SpinXY / Den correlation functions accumulated on the device at every measured slice (on the
    symmetrised G, as main.F90:761-764 hands GR_Tilde to ham%Obser) against the oracle's restatement of Predefined_Obs_eq_*_measure."""
    model = {"hubbard_mz_symm": lambda: hubbard_square(4, 4, 0.8), "hubbard_su2": lambda: hubbard_square(4, 4, 0.8, Mz=False),
             "kondo": lambda: kondo_square(2, 2, 0.6), "hubbard_mz_8x8_diag": lambda: hubbard_square(8, 8, 0.4),
             "kondo_8x4_diag": lambda: kondo_square(8, 4, 0.4)}[which]()
    seeds = SEEDS[:2]; nwrap = 4
    g = AlfB200(model, n_chains=len(seeds), nwrap=nwrap); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.obs_eq_enable()
    orcs = []
    for s in seeds:
        o = Oracle(model, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init(); o.obs_eq_enable(); orcs.append(o)
    g.sweep(1, 0)
    for o in orcs:
        o.sweep(0)
    acc, bg, n, sg = g.obs_eq()
    ro = [o.obs_eq() for o in orcs]
    acc_o = sum(r[0] for r in ro); bg_o = sum(r[1] for r in ro)
    assert n == sum(r[2] for r in ro) == len(seeds) * (2 * model.Ltrot - 1) and sg == sum(r[3] for r in ro)
    for ch in range(4):
        if np.abs(acc_o[ch]).max() > 0:
            assert relF(acc[ch], acc_o[ch]) < 1e-9, ch
        else:
            assert np.abs(acc[ch]).max() == 0
    assert np.abs(bg - bg_o).max() < 1e-9 * max(1.0, np.abs(bg_o).max())
    g.close()


@pytest.mark.gpu
def test_example_writes_alf_bin_files(tmp_path):
    """examples/hubbard_bins.py: sweeps + device-side measurements + bin files in ALF's text layout that parse back (what Analysis reads)."""
    import importlib.util
    from alf_b200.bins import read_latt, read_scal
    spec = importlib.util.spec_from_file_location("hubbard_bins", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "hubbard_bins.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = str(tmp_path / "run")
    assert mod.main(["--L1", "4", "--L2", "2", "--beta", "1.0", "--chains", "8", "--bins", "2", "--sweeps", "3", "--warmup", "2", "--nwrap", "5", "--out", out]) == 0
    obs, sign = read_scal(os.path.join(out, "Part_scal"))
    assert obs.shape == (2, 1) and np.all(np.abs(obs[:, 0].real - 8.0) < 0.5) and np.allclose(sign, 1.0)      # half filling: <N> = 8 sites
    for nm in ("Green", "SpinZ", "SpinXY", "Den"):
        beq = read_latt(os.path.join(out, nm + "_eq")); btau = read_latt(os.path.join(out, nm + "_tau"))
        assert len(beq) == 2 and len(btau) == 2 and beq[0][3].shape == (8, 1, 1, 1) and btau[0][3].shape == (8, 11, 1, 1)
    # the k-sum of the k-space function is its r = 0 value: Green_eq(r = 0) = sum_i N_SUN sum_nf <c^dag_i c_i> = 8 particles at half filling
    g_eq = read_latt(os.path.join(out, "Green_eq"))[1][3][:, 0, 0, 0]
    assert abs(g_eq.sum().real - 8.0) < 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["hubbard_continuous", "z2_matter"])
def test_example_other_models_and_restart(tmp_path, model):
    """The same end-to-end example with continuous fields and with Hamiltonian_Z2_Matter (S0 / Global_move_tau tables), including the restart
    from the confin_<chain> files the first run leaves behind: bins are appended, particle number stays at half filling."""
    import importlib.util
    from alf_b200.bins import read_scal
    spec = importlib.util.spec_from_file_location("hubbard_bins", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "hubbard_bins.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = str(tmp_path / "run")
    args = ["--model", model, "--L1", "4", "--L2", "4", "--beta", "0.6", "--chains", "4", "--bins", "1", "--sweeps", "2", "--warmup", "1", "--nwrap", "3", "--ltau", "0", "--out", out]
    assert mod.main(args) == 0 and os.path.exists(os.path.join(out, "confin_3"))
    assert mod.main(args + ["--restart"]) == 0
    obs, sign = read_scal(os.path.join(out, "Part_scal"))
    assert obs.shape == (2, 1) and np.all(np.abs(obs[:, 0].real - 16.0) < 1.5)


@pytest.mark.gpu
def test_ed_correlations_plaquette_device_observables():
    """Full chain of the device-side measurements against exact diagonalisation (testsuite/test_vs_ed in spirit): 2x2 Hubbard plaquette,
    U = 4, beta = 2: equal-time spin and density correlation functions accumulated ON THE DEVICE over 64 chains vs the 256-state Fock space."""
    beta, dtau, U = 2.0, 0.05, 4.0
    m = hubbard_square(2, 2, beta, dtau, U)
    N = m.Ndim
    # single-particle hopping matrix exactly as the operator list defines it: sum over the checkerboard operators of O * (g / -dtau)
    Tm = np.zeros((N, N))
    for ops in m.Op_T:
        op = ops[0]
        for a in range(op.N):
            for b in range(op.N):
                Tm[op.P[a] - 1, op.P[b] - 1] += (op.O[a, b] * (op.g / (-dtau))).real
    norb = 2 * N; dim = 2 ** norb

    def cdag(i):
        M = np.zeros((dim, dim))
        for s in range(dim):
            if not (s >> i) & 1:
                M[s | (1 << i), s] = (-1) ** bin(s & ((1 << i) - 1)).count("1")
        return M
    cd_ = [cdag(i) for i in range(norb)]; c_ = [x.T for x in cd_]; n_ = [cd_[i] @ c_[i] for i in range(norb)]
    H = np.zeros((dim, dim))
    for sp in range(2):
        for i in range(N):
            for j in range(N):
                if Tm[i, j] != 0.0:
                    H += Tm[i, j] * cd_[sp * N + i] @ c_[sp * N + j]
    I = np.eye(dim)
    for i in range(N):
        H += U * (n_[i] - 0.5 * I) @ (n_[N + i] - 0.5 * I)
    w, v = np.linalg.eigh(H); p = np.exp(-beta * (w - w.min())); p /= p.sum()

    def ev(Op):
        return float(np.einsum("i,ji,jk,ki->", p, v, Op, v))
    imj = m.latt.imj_table() - 1
    spin_ed = np.zeros(N); den_ed = np.zeros(N)
    for i in range(N):
        for j in range(N):
            mi, mj = n_[N + i] - n_[i], n_[N + j] - n_[j]
            spin_ed[imj[i, j]] += ev(mi @ mj); den_ed[imj[i, j]] += ev((n_[i] + n_[N + i]) @ (n_[j] + n_[N + j]))
    C = 64
    g = AlfB200(m, n_chains=C, nwrap=10); g.set_seeds([31 + 101 * c for c in range(C)]); g.fields_set(); g.init_sweep()
    g.sweep(30, 0); g.obs_eq_enable(True); g.sweep(150, 0)
    acc, bg, n, sg = g.obs_eq()
    spin = acc[1, 0, 0, 0].real / n; den = acc[3, 0, 0, 0].real / n
    assert sg == n                                       # no sign problem at half filling
    assert np.abs(spin - spin_ed).max() < 0.06 * max(1.0, np.abs(spin_ed).max()), (spin, spin_ed)
    assert np.abs(den - den_ed).max() < 0.03 * max(1.0, np.abs(den_ed).max()), (den, den_ed)
    g.close()


@pytest.mark.gpu
def test_handles_are_reentrant_across_threads():
    """Two handles with DIFFERENT models (real Hubbard 8x8 with blocked kernels' small cousins, complex Kondo) advanced concurrently by two
    host threads on their own streams give the same chains as the oracle: no shared state between handles (bench.py runs two per GPU)."""
    import threading
    models = [hubbard_square(8, 8, 1.0), kondo_square(2, 2, 0.6), hubbard_square(4, 4, 0.8, projector=True, theta=0.3, trial="dimer")]
    seeds = SEEDS[:2]
    gs = []
    for m in models:
        g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); gs.append(g)
    errs = []

    def run(g):
        try:
            g.sweep(2, 1)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(g,)) for g in gs]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    for m, g in zip(models, gs):
        f = g.get_fields(); ph = g.phase()
        for c, s in enumerate(seeds):
            o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.sweep(1); o.sweep(1)
            assert np.array_equal(f[c], o.get_fields()), m.name
            for nf in range(1, m.N_FL + 1):
                assert relF(g.green(c, nf), o.green(nf)) < TOL_G, m.name
            assert abs(ph[c] - o.phase()) < 1e-9
        g.close()


def test_confout_restart_continues_the_chain(tmp_path):
    """Checkpoint / restart through ALF's confout_<rank> text files (Prog/Fields_mod.F90:631-662,750-774): a handle restarted from the
    files written after sweep 1 reproduces sweep 2 of the uninterrupted run bit for bit (fields) and G to 1e-10."""
    from alf_b200 import conf
    m = hubbard_square(4, 4, 1.0); seeds = SEEDS[:3]
    a = AlfB200(m, n_chains=3, nwrap=5); a.set_seeds(seeds); a.fields_set(); a.init_sweep(); a.sweep(1, 0)
    files = conf.write_confs(a, str(tmp_path)); assert [os.path.basename(f) for f in files] == ["confout_0", "confout_1", "confout_2"]
    a.sweep(1, 0)
    for f in files:
        os.rename(f, os.path.join(os.path.dirname(f), os.path.basename(f).replace("confout", "confin")))
    b = AlfB200(m, n_chains=3, nwrap=5); b.set_seeds([1, 2, 3]); conf.read_confs(b, str(tmp_path)); b.init_sweep(); b.sweep(1, 0)
    assert np.array_equal(a.get_fields(), b.get_fields()) and np.array_equal(a.rng_state(), b.rng_state())
    for c in range(3):
        for nf in (1, 2):
            assert relF(b.green(c, nf), a.green(c, nf)) < TOL_G
    a.close(); b.close()


@pytest.mark.parametrize("variant", ["gauge", "gauge_propose_s0", "gauge_hubbard", "gauge_no_action"])
def test_ising_action_s0_tables_sweep_parity(variant):
    """ham%S0 as device-side tables: the Z2-gauge sector of Hamiltonian_Z2_Matter (Ising bond vertices, type 1, k = 2; Ising action
    in time and on plaquettes; Hamiltonian_Z2_Matter_smod.F90:439-512) swept on the device gives the oracle's accept / reject / not-proposed
    sequence, fields, G and phase -- also with main.F90's Propose_S0 (data-dependent number of random draws, Wrapgr_mod.F90:127-140) and
    with Hubbard vertices mixed in."""
    kw = dict(g=0.8, K=0.5)
    if variant == "gauge_hubbard":
        kw.update(U=2.0)
    if variant == "gauge_no_action":
        kw.update(g=0.0, K=0.0)
    m = z2_gauge_square(4, 4, beta=1.0, dtau=0.1, propose_s0=(variant == "gauge_propose_s0"), **kw)
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.accept_log(2); g.sweep(2, 0)
    log = g.get_accept_log(); f = g.get_fields(); ph = g.phase()
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.log(True); o.sweep(0); o.sweep(0)
        acc, _ = o.get_log()
        assert np.array_equal(acc, log[c]), variant
        if variant == "gauge_propose_s0":
            assert (acc == 2).any() and (acc == 1).any()               # some visits are not proposed at all
        assert np.array_equal(f[c], o.get_fields())
        assert relF(g.green(c, 1), o.green(1)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
    g.close()


@pytest.mark.parametrize("variant", ["finite_T", "projector", "matter_only", "with_hubbard"])
def test_z2_matter_batched_sweep_parity(variant):
    """BASELINE config 5 at test size: Hamiltonian_Z2_Matter (gauge + matter Ising bond vertices, Ising action, N_Global_tau = N/4 star moves per
    slice through ham%Global_move_tau, restricted sequential range) swept entirely on the device -- S0 and Global_move_tau from tables, no host
    in the loop -- reproduces the oracle: identical fields and random-number state after two sweeps (hence identical proposals, draws and
    decisions of every sequential visit and every global move), G and phase within tolerance.  Finite temperature and projective algorithm."""
    kw = dict(g=0.8, K=0.5, J=0.7, h=0.9)
    if variant == "projector":
        kw.update(projector=True, theta=0.5)
    if variant == "matter_only":
        kw.update(t_z2=0.0)
    if variant == "with_hubbard":
        kw.update(U=2.0)
    m = z2_matter_square(4, 4, beta=0.5 if variant == "projector" else 1.0, dtau=0.1, **kw)
    assert m.global_tau["n_global_tau"] == 4
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(2, 0)
    f = g.get_fields(); ph = g.phase(); rs = g.rng_state(); ctr = g.counters() if hasattr(g, "counters") else None
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.log(True); o.sweep(0); o.sweep(0)
        gm = o.get_gm_log(); assert (gm == 1).any() and (gm == 2).any(), np.bincount(gm)
        assert np.array_equal(f[c], o.get_fields()), variant
        assert np.array_equal(rs[c], o.rng_state()), variant
        assert relF(g.green(c, 1), o.green(1)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
    g.close()


@pytest.mark.parametrize("mz", [True, False])
def test_continuous_hs_fields_sweep_parity(mz):
    """Continuous Hubbard-Stratonovich fields (Op_V%type = 3; Hamiltonian_Hubbard with Continuous = .true.: Predefined_Int_U_MZ_continuous_HS /
    _U_SUN_continuous_HS, Gaussian ham%S0, proposal f + Amplitude (ranf - 1/2), exponentials evaluated on the fly): the batched sweep gives the
    oracle's accept / reject sequence, bit-identical real-valued fields, G and phase.  Mz: real arithmetic; SU(2): imaginary coupling, complex."""
    m = hubbard_square(4, 4, 1.0, Mz=mz, continuous=True)
    assert all(op[0].type == 3 for op in m.Op_V) and m.s0_gaussian
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.accept_log(2); g.sweep(2, 1)
    assert g.is_complex == (not mz)
    log = g.get_accept_log(); f = g.get_fields(); ph = g.phase()
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.log(True); o.sweep(1); o.sweep(1)
        acc, _ = o.get_log()
        assert np.array_equal(acc, log[c])
        fo = o.get_fields()
        assert np.array_equal(f[c], fo) and np.abs(fo.real - np.rint(fo.real)).max() > 1e-3      # genuinely continuous values
        for nf in range(1, m.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
    # host round trip of real-valued fields
    g.set_fields(f); assert np.array_equal(g.get_fields(), f)
    g.close()


@pytest.mark.parametrize("variant", ["mz", "su2_continuous", "kondo", "projector", "z2_matter"])
def test_compute_fermion_det(variant):
    """alf_b200_compute_fermion_det = Compute_Fermion_Det with storage = "Empty" (Prog/Global_mod.F90:792-1000) for every chain: log|det| (the sum
    of Det_Vec, all that Compute_Ratio_Global uses) and Phase_det against the oracle's restatement (QR + SVD as UDV_WRAP), after a sweep has
    changed the fields; the rebuilt storage then continues the Markov chain exactly as the oracle's does."""
    m = {"mz": lambda: hubbard_square(4, 4, 1.0), "su2_continuous": lambda: hubbard_square(4, 4, 1.0, Mz=False, continuous=True),
         "kondo": lambda: kondo_square(2, 2, 0.6), "projector": lambda: hubbard_square(4, 4, 0.4, projector=True, theta=0.3, trial="dimer"),
         "z2_matter": lambda: z2_matter_square(4, 4, 0.5)}[variant]()
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    ld, ph = g.compute_fermion_det()
    g.sweep(1, 0); f = g.get_fields()
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.sweep(0)
        pho, dv = o.compute_fermion_det()
        for nf in range(m.N_FL):
            assert abs(ld[c, nf] - dv[nf].sum()) < 1e-8 * max(1.0, abs(dv[nf].sum())), (variant, ld[c, nf], dv[nf].sum())
            assert abs(ph[c, nf] - pho[nf]) < 1e-8, (variant, ph[c, nf], pho[nf])
        o.sweep(0)
        assert np.array_equal(f[c], o.get_fields())
    g.close()


@pytest.mark.parametrize("mz", [True, False])
def test_langevin_forces_and_update(mz):
    """Langevin scheme for continuous fields (Prog/Langevin_HMC_mod.F90): the fermionic forces of every chain equal the oracle's (which are checked
    against a finite difference of log det in the CPU suite), and three Langevin updates -- adaptive step, Box-Muller noise from the chain's
    stream, storage reset -- leave the same fields (to rounding), the same random-number state, G and phase."""
    m = hubbard_square(4, 4, 1.0, Mz=mz, continuous=True)
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    F = g.langevin_forces()
    orcs = []
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.sweep(0); orcs.append(o)
        Fo = o.langevin_forces()
        assert np.abs(F[c] - Fo).max() < 1e-9 * max(1.0, np.abs(Fo).max()) and np.abs(Fo).max() > 1e-2
    for it in range(3):
        dt = g.langevin_update(0.02, 0.4)             # max_force small enough to make the step adaptive
        for c, o in enumerate(orcs):
            dto = o.langevin_update(0.02, 0.4)
            assert abs(dt[c] - dto) < 1e-9 * dto
    assert (dt < 0.02).any()
    f = g.get_fields(); rs = g.rng_state(); ph = g.phase()
    for c, o in enumerate(orcs):
        assert np.abs(f[c] - o.get_fields()).max() < 1e-8
        assert np.array_equal(rs[c], o.rng_state())
        for nf in range(1, m.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < 1e-7      # G of slightly different fields (forces agree to 1e-9)
        assert abs(ph[c] - o.phase()) < 1e-6
    g.close()


@pytest.mark.parametrize("mz", [True, False])
def test_hmc_update(mz):
    """Scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571): leapfrog trajectories of all chains on the device, Compute_Fermion_Det before and after,
    Compute_Ratio_Global, Metropolis test per chain.  Weight, acceptance, fields (to rounding), random-number state, G and phase follow the oracle;
    a small step conserves the Hamiltonian (Weight ~ 1), a large one does not."""
    m = hubbard_square(4, 4, 1.0, Mz=mz, continuous=True)
    seeds = SEEDS[:3]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    orcs = []
    for s in seeds:
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); o.init(); o.sweep(0); orcs.append(o)
    for (dt, nl) in ((0.02, 5), (0.6, 2), (0.1, 3)):
        acc, w = g.hmc_update(dt, nl)
        for c, o in enumerate(orcs):
            acco, wo = o.hmc_update(dt, nl)
            assert abs(w[c] - wo) < 1e-6 * max(1.0, wo), (dt, nl, w[c], wo)
            assert bool(acc[c]) == acco
        if dt == 0.02:
            assert np.all(np.abs(w - 1.0) < 5e-3)
        if dt == 0.6:
            assert np.any(np.abs(w - 1.0) > 5e-2)
    f = g.get_fields(); rs = g.rng_state(); ph = g.phase()
    for c, o in enumerate(orcs):
        assert np.abs(f[c] - o.get_fields()).max() < 1e-7
        assert np.array_equal(rs[c], o.rng_state())
        for nf in range(1, m.N_FL + 1):
            assert relF(g.green(c, nf), o.green(nf)) < 1e-6
        assert abs(ph[c] - o.phase()) < 1e-6
    g.close()


def _with_g_t(model, amp=0.3):
    """Gives every interaction vertex a time-dependent coupling g_t(nt) = g (1 + amp cos(2 pi nt / Ltrot + n)) (Operator_mod.F90:66)."""
    L = model.Ltrot
    for n, row in enumerate(model.Op_V):
        for op in row:
            op.g_t = np.array([op.g * (1.0 + amp * np.cos(2.0 * np.pi * (nt + 1) / L + 0.7 * n)) for nt in range(L)], dtype=np.complex128)
    return model


@pytest.mark.parametrize("which", ["mz_real", "su2_complex", "kondo_k2", "dense_nonsymm"])
def test_sweep_time_dependent_coupling(which):
    """Op_V%g_t: the device builds its vertex tables per time slice (alf_b200_set_op_v_gt), the oracle evaluates Op_exp on the fly as the reference does
    (Operator_mod.F90:583-604, 677-699, 768-791, 885-908): identical accept / reject sequences, fields, G and phase after a sweep."""
    if which == "mz_real":
        _run_parity(_with_g_t(config1()), SEEDS[:3], nwrap=10, n_sweeps=2, xmaxg=1e-5)      # couplings up to 1.3 g: larger wrap error than config 1
    elif which == "su2_complex":
        _run_parity(_with_g_t(config1(Mz=False)), SEEDS[:2], nwrap=10, n_sweeps=1, xmaxg=1e-5)
    elif which == "kondo_k2":
        _run_parity(_with_g_t(kondo_square(4, 4, 2.0)), SEEDS[:2], nwrap=5, n_sweeps=1, check_udv=False, xmaxg=1e-5)
    else:
        _run_parity(_with_g_t(hubbard_square(4, 4, 2.0, checkerboard=False, symm=False)), SEEDS[:2], nwrap=5, n_sweeps=1, xmaxg=1e-5)


def test_time_dependent_coupling_changes_the_chain_and_is_rejected_where_unsupported():
    m0 = config1(); m1 = _with_g_t(config1())
    ga = AlfB200(m0, n_chains=1, nwrap=10); gb = AlfB200(m1, n_chains=1, nwrap=10)
    for g in (ga, gb):
        g.set_seeds(SEEDS[:1]); g.fields_set(); g.init_sweep()
    assert relF(ga.green(0, 1), gb.green(0, 1)) > 1e-3            # g_t is actually used
    ga.close(); gb.close()
    mc = _with_g_t(hubbard_square(4, 4, 1.0, continuous=True))
    g = AlfB200(mc, n_chains=1, nwrap=5); g.set_seeds(SEEDS[:1]); g.fields_set(); g.init_sweep()
    with pytest.raises(api.AlfError):
        g.langevin_update(0.01, 1.5)
    g.close()


def test_taum_time_dependent_coupling():
    """TAU_M with Op_V%g_t: PROPR / PROPRM1 (Prog/tau_m_mod.F90:215-263) use the coupling of the slice they propagate over."""
    _run_taum(_with_g_t(config1()), SEEDS[:2], nwrap=10)
    _run_taum(_with_g_t(config1(Mz=False)), SEEDS[:2], nwrap=10)


@pytest.mark.parametrize("shape", [(4, 4), (8, 4), (6, 6), (12, 12), (16, 16), (16, 8)])
def test_hop_ring_groups_bond_dependent_amplitudes(shape):
    """Checkerboard hopping with a different amplitude on every bond operator (no translation invariance): the ring-group op kernel takes its per-bond
    matrices path (rings of 4 / 8 / 6 / 12 / 16 sites; 16 x 8: two ring lengths in one list), compared with the oracle for the six Hop_mod entry points."""
    model = hubbard_square(shape[0], shape[1], 0.3)
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    for row in model.Op_T:
        for op in row:
            op.g = op.g * (1.0 + 0.4 * rng.random())
    _hop_check(model)
    _hop_check(model, nf=2)


def test_sweep_bond_dependent_hopping():
    """A full sweep (wraps with the diagonal vertices fused into the ring passes, TAU_M) with bond-dependent hopping amplitudes: per-bond ring matrices."""
    for shape, nw in (((4, 4), 5), ((8, 8), 10)):
        model = hubbard_square(shape[0], shape[1], 1.0)
        rng = np.random.default_rng(11)
        for row in model.Op_T:
            for op in row:
                op.g = op.g * (1.0 + 0.4 * rng.random())
        _run_parity(model, SEEDS[:2], nwrap=nw, n_sweeps=1)
        _run_taum(model, SEEDS[:1], nwrap=nw)


def test_projector_time_dependent_coupling():
    """Projective algorithm + Tau_p with Op_V%g_t (wraps, CGRP and the time-displaced propagation all use the coupling of their slice)."""
    _run_projector(_with_g_t(hubbard_square(4, 4, 1.0, 0.1, 4.0, projector=True, theta=0.6, trial="dimer")), SEEDS[:2], nwrap=4, ltau=1)
    _run_projector(_with_g_t(hubbard_square(4, 4, 0.8, 0.1, 4.0, Mz=False, projector=True, theta=0.3)), SEEDS[:2], nwrap=6, ltau=1)
