"""GPU parity tests (-m gpu) at the sizes bench.py TIMES: the time-displaced path of the headline
configuration (TAU_M + CGR2_2 at 2N = 512), CGR2_2 at n = 256 real / n = 288 complex, one full sweep of the Kondo 12x12 (N = 288,
complex) and Z2_Matter 12x12 (N = 144, projective) workloads, and sweeps over k = 3 and k = 4 non-diagonal vertices
(testsuite/Prog.tests/13-Op-Wrapup.F90 exercises Op%N = 1..4).  Every case goes through the C-ABI and is compared with the CPU oracle:
accept/reject sequences and fields bit-exact, freshly recomputed Green functions to 1e-10 relative Frobenius norm."""
import threading

import numpy as np
import pytest

from alf_b200 import api
from alf_b200.api import AlfB200
from alf_b200.model import Model, Op_make, Op_set, hubbard_square, kondo_square, z2_matter_square
import oracle.oracle as O
from oracle.oracle import Oracle
from common import relF, SEEDS, TOL_G, config3
from test_gpu_parity import _run_taum, _run_parity

pytestmark = pytest.mark.gpu


def test_taum_config3_headline_size():
    """BASELINE configs[2] exactly as bench.py times it (Hubbard 16x16, beta = 10, Nwrap = 10, Ltau = 1): the four time-displaced Green
    functions right after every CGR2_2 (2N = 512 extended system, Prog/cgr2_2_mod.F90:196, Prog/tau_m_mod.F90:179-198) to 1e-10, and what
    ObserT receives every 10th slice to 1e-7."""
    _run_taum(config3(), SEEDS[:1], nwrap=10, every=10)


def _cgr2_2_case(is_complex, stab, n, batch=2):
    rng = np.random.default_rng(300 + n)
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
            assert relF(out[k][b], ref[k]) < TOL_G, (k, b, relF(out[k][b], ref[k]))
    assert len(orders) == 2


@pytest.mark.parametrize("is_complex,n", [(False, 256), (True, 288), (False, 144), (True, 128)])
@pytest.mark.parametrize("stab", [0, 3])
def test_cgr2_2_kernel_bench_sizes(is_complex, n, stab):
    """CGR2_2 at the instantiations the bench runs: n = 256 real (config 3: 512 x 512 pivoted QR, k_qrp_reg / k_apply_q2 / k_trsm_blk) and
    n = 288 complex (config 4: 576 x 576), 20 orders of magnitude of scales, both block orderings."""
    _cgr2_2_case(is_complex, stab, n)


@pytest.mark.parametrize("is_complex,n,nvar", [(False, 256, 1), (False, 256, 2), (True, 288, 1), (True, 288, 2)])
def test_cgr_kernel_bench_sizes(is_complex, n, nvar):
    """CGR (Prog/cgr1_mod.F90:36) at n = 256 real and n = 288 complex, both NVAR branches, graded scales."""
    rng = np.random.default_rng(400 + n); batch = 2
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
    G, ph = api.test_cgr(st(UR), st(DR), st(VR), st(UL), st(DL), st(VL), st(dR), st(dL), nvar, 0, is_complex)
    for b in range(batch):
        Go, pho = O.cgr(UR[b], DR[b], VR[b], UL[b], DL[b], VL[b], nvar=nvar)
        assert relF(G[b], Go) < TOL_G
        assert abs(ph[b] - pho) < 1e-9


def test_first_sweep_kondo_12x12_config4():
    """BASELINE configs[3] at full size: SU(2) Kondo lattice 12x12 (N_dim = 288, complex, 144 rank-1 + 144 rank-2 vertices), beta = 20,
    Nwrap = 5 (40 stabilisation intervals each way): first sweep of one chain -- 115 200 decisions bit-exact, fresh G to 1e-10, phase."""
    _run_parity(kondo_square(12, 12, 20.0), SEEDS[:1], nwrap=5, n_sweeps=1, check_udv=False)


def test_taum_kondo_6x6_complex_k2():
    """TAU_M on the Kondo lattice (complex, pair groups with basis rotations inside PROPR / PROPRM1): 6x6, N = 72, 2N = 144."""
    _run_taum(kondo_square(6, 6, 2.0), SEEDS[:1], nwrap=5, every=5)


def test_sweep_z2_matter_12x12_config5():
    """BASELINE configs[4] at full lattice size: Hamiltonian_Z2_Matter 12x12 (N_dim = 144, N_part = 72, 577 Ising fields, 36 star moves per
    slice), projective algorithm with a short projection (theta = 0.5, beta = 1: 20 slices): one sweep, fields and random-number state
    bit-exact (every proposal, draw and decision of the sequential visits and the global moves), G to 1e-10."""
    m = z2_matter_square(12, 12, beta=1.0, dtau=0.1, g=0.8, K=0.5, J=0.7, h=0.9, projector=True, theta=0.5)
    assert m.Ndim == 144 and m.global_tau["n_global_tau"] == 36
    seeds = SEEDS[:2]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    f = g.get_fields(); ph = g.phase(); rs = g.rng_state()
    orcs = []
    for s in seeds:
        o = Oracle(m, nwrap=5); o.ranset(s); o.fields_set(); orcs.append(o)

    def work(o):
        o.init(); o.log(True); o.sweep(0)
    th = [threading.Thread(target=work, args=(o,)) for o in orcs]
    [t.start() for t in th]; [t.join() for t in th]
    for c, o in enumerate(orcs):
        gm = o.get_gm_log(); assert (gm == 1).any()
        assert np.array_equal(f[c], o.get_fields())
        assert np.array_equal(rs[c], o.rng_state())
        assert relF(g.green(c, 1), o.green(1)) < TOL_G
        assert abs(ph[c] - o.phase()) < 1e-9
    g.close()


def _multi_site_vertex_model(k, typ, imag_g, L1=4, L2=4, beta=1.0, dtau=0.1):
    """Hubbard 4x4 hopping (N_FL = 1, N_SUN = 2) with non-diagonal k-site vertices: a fixed Hermitian k x k matrix on the sites
    (I, I+a1, I+a2, I+a1+a2)[:k] of every unit cell, coupling g real or imaginary, fields of type 1 or 2."""
    base = hubbard_square(L1, L2, beta, dtau, 4.0, Mz=False)
    latt = base.latt; rng = np.random.default_rng(7 * k + typ)
    A = rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k)); Omat = (A + A.conj().T) / 2
    if not imag_g:
        Omat = Omat.real                          # real symmetric vertex, real coupling: stays in the real instantiation
    Op_V = []
    for I in range(1, latt.N + 1):
        sites = [I, latt.nnlist(I, 1, 0), latt.nnlist(I, 0, 1), latt.nnlist(I, 1, 1)][:k]
        op = Op_make(k); op.P[:] = np.array(sites, dtype=np.int32); op.O[:, :] = Omat
        op.g = (1j if imag_g else 1.0) * np.sqrt(dtau * 0.7); op.alpha = -0.3 if imag_g else 0.0; op.type = typ
        Op_set(op); assert not op.diag and op.N_non_zero == k
        Op_V.append([op])
    return Model(name=f"k{k}", Ndim=base.Ndim, N_FL=1, N_SUN=2, Ltrot=base.Ltrot, Dtau=dtau, Symm=base.Symm, Op_V=Op_V, Op_T=base.Op_T, latt=latt)


@pytest.mark.parametrize("k", [3, 4])
@pytest.mark.parametrize("typ,imag_g", [(1, False), (2, True)])
def test_sweep_k3_k4_nondiagonal_vertices(k, typ, imag_g):
    """Vertices with Op%N = 3 and 4 (non-diagonal, overlapping supports between neighbouring cells): rank-k Woodbury updates through the k x k
    determinant (LU for k > 2, Prog/upgrade_mod.F90:168-193), Op_Wrapup / Op_Wrapdo with both N_type passes.  The reference covers Op%N = 1..4
    in testsuite/Prog.tests/13-Op-Wrapup.F90 and 14-Op-Wrapdo.F90."""
    m = _multi_site_vertex_model(k, typ, imag_g)
    _run_parity(m, SEEDS[:2], nwrap=5, n_sweeps=1, check_udv=False)
    _run_taum(m, SEEDS[:1], nwrap=5, every=5)
