"""GPU tests (-m gpu) of the tempering exchange (alf_b200/tempering.py = Exchange_Step of Prog/Global_mod.F90:108-420 over handles)."""
import numpy as np
import pytest

from alf_b200.api import AlfB200
from alf_b200.model import hubbard_square
from alf_b200 import tempering
from oracle.oracle import Oracle
from common import relF, SEEDS, TOL_G

pytestmark = pytest.mark.gpu


def _handles(models, seeds, nwrap=5, sweeps=1):
    hs = []
    for k, m in enumerate(models):
        g = AlfB200(m, n_chains=len(seeds), nwrap=nwrap); g.set_seeds([s + 1000 * k for s in seeds]); g.fields_set(); g.init_sweep()
        g.sweep(sweeps, 0); hs.append(g)
    return hs


def test_exchange_between_identical_parameter_sets_always_accepts_and_swaps():
    m = [hubbard_square(4, 4, 1.0, U=4.0), hubbard_square(4, 4, 1.0, U=4.0)]
    hs = _handles(m, SEEDS[:3])
    f0 = [h.get_fields() for h in hs]; g0 = [[h.green(c, 1) for c in range(3)] for h in hs]
    acc, w, pairs = tempering.exchange_step(hs, np.random.default_rng(1), n_exchange_steps=1)
    assert np.allclose(w, 1.0, rtol=1e-8) and acc.all() and pairs[0] in ([(0, 1)], [(0, 1)])
    assert np.array_equal(hs[0].get_fields(), f0[1]) and np.array_equal(hs[1].get_fields(), f0[0])
    for c in range(3):       # the rebuilt Green function of a handle is the one its partner had (same Hamiltonian): main.F90's GR moves with the configuration
        assert relF(hs[0].green(c, 1), g0[1][c]) < 1e-6 and relF(hs[1].green(c, 1), g0[0][c]) < 1e-6
    for h in hs:
        h.close()


def test_exchange_weight_against_the_oracle_and_restore_on_reject():
    """Two interaction strengths: the weight of a pair is W_1(conf_2) W_2(conf_1) / (W_1(conf_1) W_2(conf_2)); the fermion determinants are checked against the
    oracle's Compute_Fermion_Det of the same configurations, rejected pairs keep their configurations, accepted ones swap."""
    models = [hubbard_square(4, 4, 1.0, U=4.0), hubbard_square(4, 4, 1.0, U=4.4)]
    seeds = SEEDS[:4]
    hs = _handles(models, seeds)
    f0 = [h.get_fields() for h in hs]
    want = np.zeros(len(seeds))
    for c in range(len(seeds)):
        lw = 0.0
        for g, m in enumerate(models):
            o = Oracle(m, nwrap=5); o.ranset(SEEDS[0]); o.fields_set(); o.init()
            o.set_fields(f0[g][c]); ph_own, dv_own = o.compute_fermion_det()
            o.set_fields(f0[1 - g][c]); ph_new, dv_new = o.compute_fermion_det()
            r1, r2 = tempering.compute_ratio_global(m, dv_own.sum(axis=1)[None, :], ph_own[None, :], dv_new.sum(axis=1)[None, :], ph_new[None, :], f0[g][c][None], f0[1 - g][c][None])
            lw += np.log(np.abs(r1[0])) + r2[0]
        want[c] = np.exp(lw)
    acc, w, pairs = tempering.exchange_step(hs, np.random.default_rng(7), n_exchange_steps=1)
    assert np.allclose(w[0, 0], want, rtol=1e-6), (w[0, 0], want)
    f1 = [h.get_fields() for h in hs]
    for c in range(len(seeds)):
        if acc[0, 0, c]:
            assert np.array_equal(f1[0][c], f0[1][c]) and np.array_equal(f1[1][c], f0[0][c])
        else:
            assert np.array_equal(f1[0][c], f0[0][c]) and np.array_equal(f1[1][c], f0[1][c])
    # the rebuilt state is a valid starting point: one more sweep keeps the precision monitors quiet
    for h in hs:
        h.sweep(1, 0); ctl = h.control(); assert ctl["nan"] == 0 and ctl["XMAXG"] < 1e-5
        h.close()


def test_exchange_needs_an_even_ring():
    hs = _handles([hubbard_square(2, 2, 0.5)] * 3, SEEDS[:1])
    with pytest.raises(ValueError):
        tempering.exchange_step(hs, np.random.default_rng(0))
    for h in hs:
        h.close()


def test_global_updates_site_flip():
    """Global_Updates (Prog/Global_mod.F90:450-639) with a proposal that flips the fields of one site on every time slice: weights against the oracle's
    determinants, accepted / rejected bookkeeping, detailed balance of a move and its inverse, and the carried phase against the rebuilt one."""
    m = hubbard_square(4, 4, 1.0, U=4.0, Mz=False); seeds = SEEDS[:4]
    g = AlfB200(m, n_chains=len(seeds), nwrap=5); g.set_seeds(seeds); g.fields_set(); g.init_sweep(); g.sweep(1, 0)
    f0 = g.get_fields(); ph0 = np.asarray(g.phase()).copy()
    sites = [3, 7, 0, 12]

    def propose(f, rng):
        for c, s in enumerate(sites):
            f[c, :, s] = -f[c, :, s]
        lt0 = np.zeros(len(sites)); lt0[3] = -np.inf               # chain 3: LOG_T0_REJECTED, no proposal
        return f, lt0
    want = np.zeros(len(seeds))
    for c in range(3):
        o = Oracle(m, nwrap=5); o.ranset(SEEDS[0]); o.fields_set(); o.init()
        o.set_fields(f0[c]); pa, da = o.compute_fermion_det()
        fn = f0[c].copy(); fn[:, sites[c]] *= -1
        o.set_fields(fn); pb, db = o.compute_fermion_det()
        r1, r2 = tempering.compute_ratio_global(m, da.sum(axis=1)[None], pa[None], db.sum(axis=1)[None], pb[None], f0[c][None], fn[None])
        rt = r1[0] * np.exp(r2[0]); want[c] = abs((ph0[c] * rt).real / ph0[c].real)
    acc, w = tempering.global_updates(g, propose, np.random.default_rng(5), n_global=1)
    assert np.allclose(w[0, :3], want[:3], rtol=1e-6) and w[0, 3] == 0.0 and not acc[0, 3]
    f1 = g.get_fields()
    for c in range(4):
        fn = f0[c].copy()
        if acc[0, c]:
            fn[:, sites[c]] *= -1
        assert np.array_equal(f1[c], fn)
    # the reverse move has the inverse weight (compare on the chains that accepted; phases are +-1-ish complex numbers of modulus 1 here)
    g2 = AlfB200(m, n_chains=len(seeds), nwrap=5); g2.set_seeds(seeds); g2.fields_set(); g2.set_fields(f1); g2.init_sweep()
    acc2, w2 = tempering.global_updates(g2, propose, np.random.default_rng(6), n_global=1, rebuild=False)
    for c in range(3):
        if acc[0, c] and abs(ph0[c].imag) < 1e-9:
            assert abs(w[0, c] * w2[0, c] - 1.0) < 1e-6
    g.close(); g2.close()
