"""GPU tests (-m gpu) of the compat-mode boundary: reference routines called with the reference's own arguments on host arrays --
CGR(PHASE, NVAR, GRUP, udvr, udvl) without extra inputs, WRAPUR(NTAU, NTAU1, UDVR) / WRAPUL on the caller's UDV states -- and of the NCCL bin
reduction entry points with a single-rank communicator (the multi-rank path is exercised by bench.py under torchrun)."""
import numpy as np
import pytest

from alf_b200 import api
from alf_b200.api import AlfB200
from alf_b200.model import hubbard_square
import oracle.oracle as O
from oracle.oracle import Oracle
from common import relF, SEEDS, TOL_G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("is_complex", [False, True])
@pytest.mark.parametrize("nvar", [1, 2])
def test_cgr_standalone_reference_arguments(is_complex, nvar):
    """alf_b200_test_cgr with detUR = detUL = NULL: the phase comes out right with nothing but (udvr, udvl), as Prog/cgr1_mod.F90:36."""
    rng = np.random.default_rng(5); n = 32; batch = 2
    S = []
    for b in range(batch):
        def mk(side):
            U0 = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if is_complex else 0)
            return O.udv_decompose(U0 * np.exp(np.linspace(-6, 6, n))[None, :], np.ones(n), np.eye(n), side)
        S.append((mk("r"), mk("l")))
    st = lambda side, k: np.stack([s[side][k] for s in S])
    G, ph = api.test_cgr(st(0, 0), st(0, 1), st(0, 2), st(1, 0), st(1, 1), st(1, 2), None, None, nvar, 0, is_complex)
    for b in range(batch):
        Go, pho = O.cgr(*S[b][0], *S[b][1], nvar=nvar)
        assert relF(G[b], Go) < TOL_G and abs(ph[b] - pho) < 1e-9


def test_wrapur_wrapul_on_host_udv_states():
    """Compat mode: set_udv -> WRAPUR / WRAPUL -> get_udv reproduces the oracle's states (D to 1e-7, the product U D V to 1e-10), starting from
    states the caller owns (here: the oracle's udvr / udvl after a sweep)."""
    m = hubbard_square(4, 4, 1.0); seeds = SEEDS[:2]; nwrap = 5
    g = AlfB200(m, n_chains=len(seeds), nwrap=nwrap); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init()
        for nf in (1, 2):
            Ul, Dl, Vl = o.get_udv(0, 0, nf)                     # udvl after the storage fill = B(beta, 0)^H decomposed
            g.set_udv(1, 0, c, nf, np.eye(m.Ndim), np.ones(m.Ndim), np.eye(m.Ndim))     # udvr = 1
            g.set_udv(0, 0, c, nf, Ul, Dl, Vl)
    g.wrapur(0, nwrap)                                           # WRAPUR(0, Nwrap, udvr) on the states just handed over
    g.cgr(1)
    for c, s in enumerate(seeds):
        o = Oracle(m, nwrap=nwrap); o.ranset(s); o.fields_set(); o.init()
        f = o.get_fields()
        for nf in (1, 2):
            U, D, V = g.get_udv(1, 0, c, nf)
            B = np.eye(m.Ndim, dtype=complex)
            for nt in range(1, nwrap + 1):
                B = o.propr(nf, B, nt)                           # B(nt) ... B(1)
            assert relF(U @ np.diag(D) @ V, B) < 1e-10
    g.close()


def test_reduce_bins_single_rank_communicator():
    """alf_b200_comm_unique_id / comm_init / reduce_bins / reduce_control with nranks = 1: the NCCL plumbing (dlopen, communicator, grouped in-place
    reductions of the scalar, equal-time and time-displaced accumulators) runs and leaves the single rank's accumulators unchanged."""
    m = hubbard_square(4, 4, 1.0); seeds = SEEDS[:2]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
    g.obs_eq_enable(True); g.obs_tau_enable(True); g.sweep(1, 1)
    before = (g.obs().copy(), [np.array(x) for x in g.obs_eq()[:2]], [np.array(x) for x in g.obs_tau()[:2]], g.control())
    g.comm_init(1, 0, AlfB200.comm_unique_id())
    g.reduce_bins(0); ctl = g.reduce_control(0)
    assert np.array_equal(g.obs(), before[0])
    for a, b in zip(g.obs_eq()[:2], before[1]):
        assert np.array_equal(a, b)
    for a, b in zip(g.obs_tau()[:2], before[2]):
        assert np.array_equal(a, b)
    assert ctl["NC_up"] == before[3]["NC_up"] and ctl["XMAXG"] == before[3]["XMAXG"]
    g.close()


def test_measure_interval_lobs():
    """alf_b200_set_measure_interval = LOBS_ST / LOBS_EN of VAR_QMC (Prog/QMC_runtime_var_mod.F90:156-189): ham%Obser is called for lobs_st <= NTAU1 <= lobs_en
    on the way up (NTAU1 = 1 .. Ltrot, Prog/main.F90:757-773) and on the way down (NTAU1 = Ltrot - 1 .. 0, :789-802); the projective window is checked."""
    m = hubbard_square(4, 4, 1.0); L = m.Ltrot; seeds = SEEDS[:2]
    for (a, b) in ((0, 0), (3, 7), (1, 1), (L, L)):
        g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep()
        g.set_measure_interval(a, b); g.sweep(1, 0)
        lo, hi = (a or 1), (b or L)
        n_up = sum(1 for nt1 in range(1, L + 1) if lo <= nt1 <= hi); n_dn = sum(1 for nt1 in range(L - 1, -1, -1) if lo <= nt1 <= hi)
        assert g.obs()[0] == len(seeds) * (n_up + n_dn), (a, b)
        g.close()
    g = AlfB200(m, n_chains=1, nwrap=5)
    with pytest.raises(api.AlfError):
        g.set_measure_interval(L + 1, 0)
    g.close()
    from alf_b200.model import hubbard_chain
    mp = hubbard_chain(4, 1.0, 0.1, projector=True, theta=0.5)
    gp = AlfB200(mp, n_chains=1, nwrap=5)
    with pytest.raises(api.AlfError):
        gp.set_measure_interval(1, 0)            # LOBS_ST < Thtrot + 1
    gp.set_measure_interval(mp.Thtrot + 1, mp.Ltrot - mp.Thtrot)
    gp.close()
