"""CPU suite (-m "not gpu"): pins the oracle against the reference's own golden vector / closed-form test inputs and
against brute-force dense algebra, checks the host model tables, and that the C-ABI library loads and exports every
symbol include/alf_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.linalg as sl

from alf_b200.model import Model, Op_make, Op_set, HamiltonianError, hubbard_square, hubbard_chain, kondo_square, Lattice
from oracle.oracle import Oracle
import oracle.oracle as O
from common import relF, SEEDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHI2 = {-2: -np.sqrt(2 * (3 + np.sqrt(6))), -1: -np.sqrt(2 * (3 - np.sqrt(6))), 1: np.sqrt(2 * (3 - np.sqrt(6))), 2: np.sqrt(2 * (3 + np.sqrt(6)))}


def dense_op(op, N):
    A = np.zeros((N, N), complex); P = op.P - 1
    A[np.ix_(P, P)] = op.O
    return A


def phi(op, s):
    return float(s) if op.type == 1 else PHI2[int(s)]


def test_golden_vector_26():
    """testsuite/Prog.tests/26-Test-Polymorphic-Fortran.F90:79-86: res(5,4) after 5 real + 5 complex OpT applied right and left."""
    N = 5; ops = []
    for i in range(5):
        a = Op_make(N); b = Op_make(N)
        a.P[:] = np.arange(1, N + 1); b.P[:] = np.arange(1, N + 1)
        g = float(np.float32(0.2))          # "Op_T(i)%g = 0.2" is a single-precision literal in the reference test
        a.g = g; b.g = g; a.type = 2; b.type = 2
        for j in range(1, N + 1):
            if j + 1 <= N:
                b.O[j - 1, j] = complex(j, j); b.O[j, j - 1] = complex(j, -j)
        Op_set(a); Op_set(b); ops += [[a], [b]]
    res = np.eye(N, dtype=complex)
    for k in range(10):
        mk = Model(name="t26", Ndim=N, N_FL=1, N_SUN=1, Ltrot=1, Dtau=0.1, Symm=False, Op_V=[], Op_T=[ops[k]])
        ok = Oracle(mk, nwrap=1)
        res = ok.hop_apply(2, 1, res)       # rmult
        res = ok.hop_apply(4, 1, res)       # lmult
    assert abs(res[4, 3].real - 559995.58637168515) <= 559995.58637168515 * 1e-13
    assert abs(res[4, 3].imag + 559995.58637168526) <= 559995.58637168526 * 1e-13


def _vertex_model(N, k, typ, diag, seed=0):
    rng = np.random.default_rng(seed)
    op = Op_make(k)
    op.P[:] = np.sort(rng.choice(np.arange(1, N + 1), size=k, replace=False))
    if diag:
        for i in range(k):
            op.O[i, i] = rng.normal()
    else:
        A = rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k)); op.O[:, :] = A + A.conj().T
    op.g = complex(0.3, 0.1); op.alpha = 0.2; op.type = typ
    Op_set(op)
    t = Op_make(1); t.P[0] = 1; t.g = 0.0; Op_set(t)
    return Model(name="v", Ndim=N, N_FL=1, N_SUN=1, Ltrot=1, Dtau=0.1, Symm=False, Op_V=[[op]], Op_T=[[t]]), op


@pytest.mark.parametrize("k", [1, 2, 3, 4])
@pytest.mark.parametrize("typ,s", [(1, 1), (1, -1), (2, 2), (2, -1)])
@pytest.mark.parametrize("diag", [True, False])
def test_op_wrapup_wrapdo_mmult_vs_dense(k, typ, s, diag):
    """Inputs in the spirit of testsuite/Prog.tests/13-Op-Wrapup.F90 (N=5, k=1..4, types 1,2; tol 5e-14) and
    14-Op-Wrapdo / 10,11,27-Op-mmult*: the oracle's sparse routines against dense e^{g phi O} algebra."""
    N = 5
    m, op = _vertex_model(N, k, typ, diag, seed=k * 10 + typ)
    o = Oracle(m, nwrap=1)
    rng = np.random.default_rng(1)
    A = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    E = sl.expm(op.g * phi(op, s) * dense_op(op, N)); Ei = np.linalg.inv(E)
    assert relF(o.op_mmultR(1, 1, A, s, "n"), E @ A) < 2e-12
    assert relF(o.op_mmultR(1, 1, A, s, "c"), E.conj().T @ A) < 2e-12
    assert relF(o.op_mmultL(1, 1, A, s, "n", 1), A @ E) < 2e-12
    assert relF(o.op_mmultL(1, 1, A, s, "n", -1), A @ Ei) < 2e-12
    up = o.op_wrapup(1, 1, o.op_wrapup(1, 1, A, s, 1), s, 2)
    assert relF(up, E @ A @ Ei) < 5e-12
    dn = o.op_wrapdo(1, 1, o.op_wrapdo(1, 1, A, s, 2), s, 1)
    assert relF(dn, Ei @ A @ E) < 5e-12


@pytest.mark.parametrize("n", [2, 5, 10, 20, 50])
def test_qdrp_hilbert(n):
    """testsuite/Prog.tests/20-qdrp.F90: Hilbert matrices n=2..50, reconstruct A P = Q D R to 1e-14 (relative)."""
    A = np.array([[1.0 / (i + j + 1) for j in range(n)] for i in range(n)], dtype=complex)
    QR, D, ipvt, tau = O.qdrp(A)
    R = np.triu(QR)
    Q = np.eye(n, dtype=complex)
    for j in range(n):          # Q = H_1 ... H_n
        v = np.zeros(n, complex); v[j] = 1; v[j + 1:] = QR[j + 1:, j]
        Q = Q @ (np.eye(n) - tau[j] * np.outer(v, v.conj()))
    rec = Q @ np.diag(D) @ R
    assert relF(rec, A[:, ipvt - 1]) < 1e-13
    assert np.all(np.abs(np.abs(np.diag(R)) - 1) < 1e-12)


@pytest.mark.parametrize("side", ["r", "l"])
def test_udv_decompose_invariants(side):
    """testsuite/Prog.tests/24-udv.F90 (N=16): U D V (side r) / U D V^H (side l) is unchanged, U unitary, det V = 1 (real case)."""
    rng = np.random.default_rng(5); n = 16
    U0 = rng.normal(size=(n, n)) + 0j; V0 = np.linalg.qr(rng.normal(size=(n, n)))[0] + 0j; D0 = np.exp(rng.normal(size=n) * 3) + 0j
    if np.linalg.det(V0).real < 0:
        V0[:, 0] = -V0[:, 0]
    B = U0 @ np.diag(D0) @ (V0 if side == "r" else V0.conj().T)
    U, D, V = O.udv_decompose(U0, D0, V0, side)
    B2 = U @ np.diag(D) @ (V if side == "r" else V.conj().T)
    assert relF(B2, B) < 1e-12
    assert relF(U.conj().T @ U, np.eye(n)) < 1e-13
    assert abs(np.linalg.det(V) - 1) < 1e-10


@pytest.mark.parametrize("shape", [(16, 16), (24, 9)])
def test_udv_wrap_pivot_invariants(shape):
    """UDV_Wrap_Pivot (Prog/UDV_WRAP_mod.F90:125-208) on the products of testsuite/Prog.tests/24-udv.F90 (entries 4(u-1/2) + 2i(u-1/2), ten
    factors, N = 16) and on a rectangular (projector) matrix: A = U D V, U column-orthonormal, D > 0, det V = 1."""
    rng = np.random.default_rng(4782347); n1, n2 = shape
    A = np.eye(n1, n2, dtype=np.complex128)
    for _ in range(10):
        A = (4 * (rng.random((n1, n1)) - 0.5) + 2j * (rng.random((n1, n1)) - 0.5)) @ A
        U, D, V = O.udv_wrap_pivot(A)
        assert relF(U @ np.diag(D) @ V, A) < 1e-12
        assert relF(U.conj().T @ U, np.eye(n2)) < 1e-12
        assert np.all(D.real > 0) and np.all(D.imag == 0)
        assert abs(np.linalg.det(V) - 1) < 1e-9
        # V is an upper triangular matrix with unit-modulus diagonal up to the column permutation of the norm sort
        order = np.argsort(-np.sum(np.abs(A) ** 2, axis=0), kind="stable")
        Vp = V[:, order]
        assert np.abs(np.tril(Vp, -1)).max() < 1e-13 and np.allclose(np.abs(np.diag(Vp)), 1.0)


@pytest.mark.parametrize("nvar", [1, 2])
@pytest.mark.parametrize("stab3", [False, True])
def test_cgr_vs_direct_inverse(nvar, stab3):
    """testsuite/Prog.tests/15-cgr.F90 (N=5, NVAR=1,2; G to 1e-10, phase to 1e-13) with UDV states produced by decompose."""
    rng = np.random.default_rng(11); n = 5
    def mk(side):
        U0 = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)); return O.udv_decompose(U0, np.ones(n), np.eye(n), side)
    UR, DR, VR = mk("r"); UL, DL, VL = mk("l")
    BR = UR @ np.diag(DR) @ VR; BL = VL @ np.diag(DL) @ UL.conj().T
    G, ph = O.cgr(UR, DR, VR, UL, DL, VL, nvar=nvar, stab3=stab3)
    M = np.eye(n) + BR @ BL
    assert relF(G, np.linalg.inv(M)) < 1e-10
    d = np.linalg.det(M)
    assert abs(ph - d / abs(d)) < 1e-12


@pytest.mark.parametrize("cfg", [dict(L=4, beta=1.0, cb=True, symm=True, Mz=True), dict(L=4, beta=2.0, cb=True, symm=False, Mz=True),
                                 dict(L=4, beta=1.0, cb=False, symm=False, Mz=True), dict(L=2, beta=2.0, cb=True, symm=True, Mz=False)])
def test_green_vs_bruteforce(cfg):
    """Independent cross-check of the oracle: G(0) = (1 + B_L ... B_1)^-1 by direct dense products (small beta)."""
    m = hubbard_square(cfg["L"], cfg["L"], cfg["beta"], 0.1, 4.0, checkerboard=cfg["cb"], symm=cfg["symm"], Mz=cfg["Mz"])
    o = Oracle(m, nwrap=5); o.ranset(12345); o.fields_set(); f = o.get_fields(); o.init()
    for nf in range(m.N_FL):
        Bt = np.eye(m.Ndim, dtype=complex)
        for nt in range(m.Ltrot):
            B = np.eye(m.Ndim, dtype=complex)
            for nc in range(m.n_opt - 1, -1, -1):
                op = m.Op_T[nc][nf]; B = sl.expm(op.g * dense_op(op, m.Ndim)) @ B
            for n in range(m.n_opv):
                op = m.Op_V[n][nf]; B = sl.expm(op.g * phi(op, round(f[nt, n].real)) * dense_op(op, m.Ndim)) @ B
            Bt = B @ Bt
        Mx = np.eye(m.Ndim) + Bt
        assert relF(o.green(nf + 1), np.linalg.inv(Mx)) < 1e-11
    # the sweep keeps the wrapped G consistent with the recomputed one and the time-displaced blocks with CGR2_2
    o.sweep(1)
    c = o.control()
    assert c["XMAXG"] < 1e-9 and c["XMAX_tau"] < 1e-8 and c["nan"] == 0 and c["NC_up"] == 2 * m.Ltrot * m.n_opv


def test_rng_and_fields_set_deterministic():
    m = hubbard_square(4, 4, 1.0)
    a = Oracle(m, 5); b = Oracle(m, 5)
    a.ranset(SEEDS[0]); b.ranset(SEEDS[0])
    assert [a.ranf() for _ in range(5)] == [b.ranf() for _ in range(5)]
    a.fields_set(); b.fields_set()
    f = a.get_fields()
    assert np.array_equal(f, b.get_fields()) and set(np.unique(f.real)) <= {-1.0, 1.0}
    u = np.array([a.ranf() for _ in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01


def test_ed_energy_two_site():
    """Physics anchor in the spirit of testsuite/test_vs_ed: 2-site Hubbard (dense hopping, Mz), energy against exact
    diagonalisation of the 16-state Fock space; Trotter error O(dtau^2) and Monte Carlo error bound the tolerance."""
    t, U, beta, dtau = 1.0, 4.0, 2.0, 0.05
    # ED: basis of 2 sites x 2 spins
    import itertools
    def cdag(i, n=4):
        dim = 2 ** n; M = np.zeros((dim, dim))
        for s in range(dim):
            if not (s >> i) & 1:
                sign = (-1) ** bin(s & ((1 << i) - 1)).count("1"); M[s | (1 << i), s] = sign
        return M
    cd = [cdag(i) for i in range(4)]; c = [x.T for x in cd]
    nn = [cd[i] @ c[i] for i in range(4)]
    H = -t * sum(cd[2 * s] @ c[2 * s + 1] + cd[2 * s + 1] @ c[2 * s] for s in range(2)) * 1.0
    H = -t * (cd[0] @ c[1] + cd[1] @ c[0] + cd[2] @ c[3] + cd[3] @ c[2])      # site0/1 spin up = 0,1 ; spin down = 2,3
    H = H + U * ((nn[0] - 0.5 * np.eye(16)) @ (nn[2] - 0.5 * np.eye(16)) + (nn[1] - 0.5 * np.eye(16)) @ (nn[3] - 0.5 * np.eye(16)))
    w, v = np.linalg.eigh(H); Z = np.exp(-beta * w)
    E_ed = float((w * Z).sum() / Z.sum())
    m = hubbard_chain(2, beta, dtau, U=U, t=t, Mz=True, symm=False)
    o = Oracle(m, nwrap=10); o.ranset(4711); o.fields_set(); o.init()
    Tm = np.array([[0, -t], [-t, 0]], float)
    es = []
    for sw in range(600):
        o.sweep(0)
        if sw >= 100:
            Gu = o.green(1).real; Gd = o.green(2).real
            kin = np.sum(Tm * ((np.eye(2) - Gu).T + (np.eye(2) - Gd).T))
            nu = 1 - np.diag(Gu); ndn = 1 - np.diag(Gd)
            pot = U * np.sum((nu - 0.5) * (ndn - 0.5))
            es.append(kin + pot)
    es = np.array(es); nb = 10; bins = es[: len(es) // nb * nb].reshape(nb, -1).mean(1)
    err = bins.std(ddof=1) / np.sqrt(nb)
    assert abs(bins.mean() - E_ed) < max(4 * err, 0.03), (bins.mean(), err, E_ed)


def test_op_set_validation():
    """The WILL_FAIL cases of testsuite/Prog.tests/CMakeLists.txt:132-145 (Op_set validation)."""
    op = Op_make(2); op.P[:] = [1, 2]; op.O[0, 1] = 1.0; op.O[1, 0] = 2.0; op.type = 2
    with pytest.raises(HamiltonianError):
        Op_set(op)
    op = Op_make(1); op.P[0] = 0; op.O[0, 0] = 1.0
    with pytest.raises(HamiltonianError):
        Op_set(op)
    op = Op_make(1); op.P[0] = 1; op.O[0, 0] = 1.0; op.type = 7
    with pytest.raises(HamiltonianError):
        Op_set(op)
    op = Op_make(1); op.P[0] = 1; op.O[0, 0] = 1j
    with pytest.raises(HamiltonianError):
        Op_set(op)


def test_model_tables():
    m = hubbard_square(4, 4, 5.0)
    assert m.Ndim == 16 and m.N_FL == 2 and m.N_SUN == 1 and m.Ltrot == 50 and m.n_opv == 16
    assert m.n_opt == 7 * 8            # Symm: 2*4-1 families of N/2 bonds (Predefined_Hop_mod.F90:1422-1495)
    # every site appears exactly once per family
    for fam in range(7):
        sites = sorted(int(p) for row in m.Op_T[fam * 8:(fam + 1) * 8] for p in row[0].P)
        assert sites == list(range(1, 17))
    gs = sorted(set(round(abs(row[0].g), 12) for row in m.Op_T))
    assert gs == [0.05, 0.1]
    assert abs(m.Op_V[0][0].g - np.sqrt(0.1 * 4 / 2)) < 1e-15 and abs(m.Op_V[0][1].g + np.sqrt(0.2)) < 1e-15
    k = kondo_square(4, 4, 2.0)
    assert k.Ndim == 32 and k.n_opv == 32 and k.Op_V[16][0].N == 2 and not k.Op_V[16][0].diag and k.Op_V[0][0].g.imag != 0
    latt = Lattice(4, 4)
    assert latt.N == 16 and latt.nnlist(1, 0, 1) == 2 and latt.nnlist(4, 0, 1) == 1


def test_cabi_library_exports_all_symbols():
    """The C-ABI shared library loads here (no GPU needed) and exports every symbol the header declares."""
    import __graft_entry__ as ge
    ge.build()
    from alf_b200 import api
    lib = ctypes.CDLL(api.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "alf_b200.h")).read()
    names = sorted(set(re.findall(r"\b(alf_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
    # without a device creation must fail loudly, never fall back
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert lib.alf_b200_create(ctypes.byref(h), 16, 2, 1, 50, 10, 16, 56, 1, 0, 4, 0) == 100


# ---------------------------------------------------------------------------------- projective algorithm (CGRP, Tau_p)
def test_cgrp_vs_direct():
    """CGRP (Prog/cgr1_mod.F90:464-515): G = 1 - U_R (U_L^H U_R)^-1 U_L^H and the phase of det(U_L^H U_R); inputs in the
    spirit of testsuite/Prog.tests/23-cgrp.F90 (N = 5) plus a random complex case."""
    from oracle.oracle import cgrp
    rng = np.random.default_rng(23)
    for n, npart in ((5, 2), (12, 6), (16, 16)):
        UR = rng.normal(size=(n, npart)) + 1j * rng.normal(size=(n, npart))
        UL = rng.normal(size=(n, npart)) + 1j * rng.normal(size=(n, npart))
        G, ph = cgrp(UR, UL)
        S = UL.conj().T @ UR
        assert relF(G, np.eye(n) - UR @ np.linalg.solve(S, UL.conj().T)) < 1e-11
        d = np.linalg.det(S)
        assert abs(ph - d / abs(d)) < 1e-11


@pytest.mark.parametrize("trial,Mz", [("flux", True), ("dimer", True), ("flux", False)])
def test_projector_green_vs_bruteforce(trial, Mz):
    """Projective sweep of the oracle against dense products: G(0) = 1 - P_R (P_L^H B_L..B_1 P_R)^-1 P_L^H B_L..B_1."""
    m = hubbard_square(4, 4, 0.4, 0.1, 4.0, Mz=Mz, projector=True, theta=0.3, trial=trial)
    assert m.Projector and m.Thtrot == 3 and m.Ltrot == 10 and m.N_part == 8
    o = Oracle(m, nwrap=4); o.ranset(4711); o.fields_set(); f = o.get_fields(); o.init()

    def Bmat(nf, nt):
        B = np.eye(m.Ndim, dtype=complex)
        for nc in range(m.n_opt - 1, -1, -1):
            op = m.Op_T[nc][nf]; B = sl.expm(op.g * dense_op(op, m.Ndim)) @ B
        for n in range(m.n_opv):
            op = m.Op_V[n][nf]; B = sl.expm(op.g * phi(op, round(f[nt, n].real)) * dense_op(op, m.Ndim)) @ B
        return B
    for nf in range(m.N_FL):
        Bt = np.eye(m.Ndim, dtype=complex)
        for nt in range(m.Ltrot):
            Bt = Bmat(nf, nt) @ Bt
        PL, PR = m.WF_L[nf], m.WF_R[nf]
        left = PL.conj().T @ Bt
        G0 = np.eye(m.Ndim) - PR @ np.linalg.solve(left @ PR, left)
        assert relF(o.green(nf + 1), G0) < 1e-10
    # one sweep with Tau_p: wrapped and recomputed Green functions stay consistent (Control_PrecisionG / _tau)
    o.sweep(1)
    c = o.control()
    assert c["XMAXG"] < 1e-8 and c["XMAX_tau"] < 1e-8 and c["NCG_tau"] > 0 and c["nan"] == 0


def test_s0_tables_of_z2_gauge_model_match_the_ising_action():
    """The Ising action of Hamiltonian_Z2_Matter_smod.F90 (gauge sector: transverse-field coupling in time, :841-845; plaquette flux term,
    :846-850) evaluated by brute force: for random single flips the ratio of Boltzmann weights equals the table-driven S0(n, nt)
    (:439-512) that the oracle and the CUDA path use."""
    from alf_b200.model import z2_gauge_square
    dtau, g, K = 0.1, 0.7, 0.4
    m = z2_gauge_square(4, 4, beta=0.6, dtau=dtau, g=g, K=K)
    o = Oracle(m, nwrap=3); o.ranset(77); o.fields_set()
    f = o.get_fields().real.copy()                                     # (Ltrot, n_opv)
    L, M = f.shape; latt = m.latt; inv = m.params["field_inv"]; FL = {(I, no): n for n, (I, no, ty) in enumerate(inv)}
    gam = -0.5 * np.log(np.tanh(dtau * g))

    def log_weight(c):
        s = gam * np.sum(c * np.roll(c, -1, axis=0))
        for I in range(1, latt.N + 1):                                   # plaquette with lower-left corner I
            Ix, Iy = latt.nnlist(I, 1, 0), latt.nnlist(I, 0, 1)
            s += -dtau * K * np.sum(c[:, FL[(I, 1)]] * c[:, FL[(I, 2)]] * c[:, FL[(Iy, 1)]] * c[:, FL[(Ix, 2)]])
        return s
    rng = np.random.default_rng(0); w0 = log_weight(f)
    for _ in range(40):
        n, nt = int(rng.integers(0, M)), int(rng.integers(0, L))
        c = f.copy(); c[nt, n] = -c[nt, n]
        assert abs(np.exp(log_weight(c) - w0) / o.s0(n + 1, nt + 1) - 1) < 1e-12, (n, nt)


@pytest.mark.parametrize("projector", [False, True])
def test_z2_matter_tables_match_the_ising_action(projector):
    """Full Hamiltonian_Z2_Matter action by brute force (gauge: time + flux + coupling to matter bonds; matter: tau_I(nt) tau_I(nt+1) with the site
    variables of Hamiltonian_set_Z2_matter, :1288-1317, rebuilt here by walking the lattice): the Boltzmann-weight ratio of a single gauge flip
    equals the table-driven S0(n, nt) (:439-512) and that of a star move equals S0_Matter of Global_move_tau (:535-643), with periodic
    (finite temperature) and open (projective) boundaries in imaginary time."""
    from alf_b200.model import z2_matter_square
    dtau, g, K, J, h = 0.1, 0.7, 0.4, 0.6, 0.9
    m = z2_matter_square(4, 4, beta=0.4, dtau=dtau, g=g, K=K, J=J, h=h, projector=projector, theta=0.2)
    o = Oracle(m, nwrap=3); o.ranset(91); o.fields_set()
    f = o.get_fields().real.copy(); L, M = f.shape; latt = m.latt; inv = m.params["field_inv"]
    FL = {(I, no, ty): n for n, (I, no, ty) in enumerate(inv)}
    assert M == 4 * latt.N + 1 and m.global_tau == dict(nt_seq_start=1, nt_seq_end=2 * latt.N, n_global_tau=latt.N // 4)
    gg, gh = -0.5 * np.log(np.tanh(dtau * g)), -0.5 * np.log(np.tanh(dtau * h))

    def tau(c):                                                         # Hamiltonian_set_Z2_matter on every slice
        t = np.zeros((L, latt.N + 1)); t[:, latt.N] = c[:, FL[(latt.N, 3, 4)]]; I = latt.N
        for _ in range(latt.L1):
            for _ in range(latt.L2):
                I1 = latt.nnlist(I, 0, 1); t[:, I1] = t[:, I] * c[:, FL[(I, 2, 2)]]; I = I1
            I1 = latt.nnlist(I, 1, 0); t[:, I1] = t[:, I] * c[:, FL[(I, 1, 2)]]; I = I1
        return t[:, 1:]

    def log_weight(c):
        pairs = (lambda a: a[:-1] * a[1:]) if projector else (lambda a: a * np.roll(a, -1, axis=0))
        gauge = c[:, [FL[(I, no, 1)] for I in range(1, latt.N + 1) for no in (1, 2)]]
        s = gg * np.sum(pairs(gauge)) + gh * np.sum(pairs(tau(c)))
        for I in range(1, latt.N + 1):
            Ix, Iy = latt.nnlist(I, 1, 0), latt.nnlist(I, 0, 1)
            s += -dtau * K * np.sum(c[:, FL[(I, 1, 1)]] * c[:, FL[(I, 2, 1)]] * c[:, FL[(Iy, 1, 1)]] * c[:, FL[(Ix, 2, 1)]])
            for no in (1, 2):
                s += -dtau * J * np.sum(c[:, FL[(I, no, 1)]] * c[:, FL[(I, no, 2)]])
        return s
    rng = np.random.default_rng(1); w0 = log_weight(f)
    for nt in list(rng.integers(0, L, size=12)) + [0, L - 1]:
        n = int(rng.integers(0, 2 * latt.N))                             # a gauge field
        c = f.copy(); c[nt, n] = -c[nt, n]
        assert abs(np.exp(log_weight(c) - w0) / o.s0(n + 1, nt + 1) - 1) < 1e-11, (n, nt)
        I = int(rng.integers(1, latt.N + 1)) if nt not in (0, L - 1) else latt.N
        star = [FL[(I, 1, 2)], FL[(I, 2, 2)], FL[(latt.nnlist(I, -1, 0), 1, 2)], FL[(latt.nnlist(I, 0, -1), 2, 2)]] + ([FL[(latt.N, 3, 4)]] if I == latt.N else [])
        c = f.copy(); c[nt, star] = -c[nt, star]
        assert abs(np.exp(log_weight(c) - w0) / o.global_move_s0(I, nt + 1) - 1) < 1e-11, (I, nt)
        gm = m.global_move_tau_ising
        assert sorted(x + 1 for x in star) == list(gm["move_fields"][gm["move_start"][I - 1]:gm["move_start"][I]])


def test_compute_fermion_det_matches_brute_force_determinant():
    """Compute_Fermion_Det (Prog/Global_mod.F90:792-1000; what Global_Updates and the tempering exchange weigh configurations with): sum of Det_Vec
    and Phase_det equal log|det| and the phase of det(1 + B(beta, 0)) built slice by slice with PROPR -- real (Mz), complex with a
    non-trivial phase (SU(2) continuous fields, Kondo), and the projector formula det(P_L^H B P_R) up to the normalisation the UDV steps keep."""
    from alf_b200.model import hubbard_square, kondo_square
    for m in (hubbard_square(2, 2, 0.5), hubbard_square(4, 2, 0.6, Mz=False), kondo_square(2, 2, 0.4), hubbard_square(4, 4, 2.0, Mz=False, continuous=True)):
        o = Oracle(m, nwrap=2); o.ranset(11); o.fields_set(); o.init()
        ph, dv = o.compute_fermion_det()
        for nf in range(1, m.N_FL + 1):
            B = np.eye(m.Ndim, dtype=complex, order="F")
            for nt in range(1, m.Ltrot + 1):
                B = o.propr(nf, B, nt)
            d = np.linalg.det(np.eye(m.Ndim) + B)
            assert abs(dv[nf - 1].sum() - np.log(abs(d))) < 1e-9 * max(1.0, abs(np.log(abs(d)))), m.name
            assert abs(ph[nf - 1] - d / abs(d)) < 1e-9, m.name
    m = hubbard_square(4, 4, 0.4, projector=True, theta=0.3, trial="dimer")
    o = Oracle(m, nwrap=2); o.ranset(3); o.fields_set(); o.init()
    ph, dv = o.compute_fermion_det()
    for nf in range(1, m.N_FL + 1):
        B = np.asfortranarray(m.WF_R[nf - 1].astype(complex)); np_ = B.shape[1]
        Bf = np.zeros((m.Ndim, m.Ndim), dtype=complex, order="F"); Bf[:, :np_] = B
        for nt in range(1, m.Ltrot + 1):
            Bf = o.propr(nf, Bf, nt)
        d = np.linalg.det(m.WF_L[nf - 1].conj().T @ Bf[:, :np_])
        assert abs(dv[nf - 1].sum() - np.log(abs(d))) < 1e-8 and abs(ph[nf - 1] - d / abs(d)) < 1e-8


def test_langevin_forces_are_the_gradient_of_the_fermion_action():
    """Langevin_HMC_Forces (Prog/Langevin_HMC_mod.F90:107-226) against a finite difference: Force(n, nt) = -d/dphi(n, nt) [N_SUN sum_nf log det(1 + B)]
    with the determinant from Compute_Fermion_Det -- ties the two restatements together."""
    from alf_b200.model import hubbard_square
    m = hubbard_square(4, 2, 0.6, continuous=True)
    o = Oracle(m, nwrap=3); o.ranset(7); o.fields_set(); o.init()
    F = o.langevin_forces(); f = o.get_fields().copy()

    def logdet(fields):
        o.set_fields(fields); ph, dv = o.compute_fermion_det()
        return m.N_SUN * sum(dv[nf].sum() for nf in range(m.N_FL))
    eps = 1e-5
    for (nt, n) in ((0, 0), (2, 5), (5, 7)):
        fp, fm = f.copy(), f.copy(); fp[nt, n] += eps; fm[nt, n] -= eps
        fd = -(logdet(fp) - logdet(fm)) / (2 * eps)
        assert abs(fd - F[nt, n].real) < 1e-6 * max(1.0, abs(fd)), (nt, n, fd, F[nt, n])


def test_hmc_conserves_the_hamiltonian_to_second_order():
    """Scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571) in the oracle: the Metropolis weight exp(-Delta H) of a leapfrog trajectory of fixed length tends to 1
    like dt^2 -- forces (Langevin_HMC_Forces), determinants (Compute_Fermion_Det), Gaussian action and phase factors (Compute_Ratio_Global) are
    mutually consistent.  Same random momenta for every step size."""
    from alf_b200.model import hubbard_square
    m = hubbard_square(4, 2, 0.6, Mz=False, continuous=True)
    errs = []
    for dt, nl in ((0.1, 2), (0.05, 4), (0.025, 8)):
        o = Oracle(m, nwrap=3); o.ranset(99); o.fields_set(); o.init(); o.sweep(0)
        acc, w = o.hmc_update(dt, nl)
        errs.append(abs(np.log(w)))
    assert errs[0] < 0.2 and errs[1] < 0.4 * errs[0] and errs[2] < 0.4 * errs[1], errs
