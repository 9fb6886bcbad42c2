"""Host logic of the tempering exchange (alf_b200/tempering.py): Compute_Ratio_Global (Prog/Global_mod.F90:651-760) on synthetic determinants."""
import numpy as np

from alf_b200 import tempering
from alf_b200.model import hubbard_square


def test_field_tables_match_fields_init():
    """Phi_st / Gama_st of Prog/Fields_mod.F90:270-285: the four-valued HS fields reproduce the moments of 4 exp(x^2): sum gamma = 4, sum gamma phi^2 = 8,
    sum gamma phi^4 = 48, sum gamma phi^6 = 480 (so that sum_l gamma(l) exp(a phi(l) O) = 4 exp(a^2 O^2) up to a^8)."""
    p, g = tempering.PHI_ST[2], tempering.GAMA_ST[2]
    idx = [0, 1, 3, 4]
    assert abs(g[idx].sum() - 4.0) < 1e-14 and abs((g[idx] * p[idx] ** 2).sum() - 8.0) < 1e-13 and abs((g[idx] * p[idx] ** 4).sum() - 48.0) < 1e-12
    assert abs((g[idx] * p[idx] ** 6).sum() - 480.0) < 1e-10
    assert np.array_equal(tempering.PHI_ST[1], [-2, -1, 0, 1, 2]) and np.all(tempering.GAMA_ST[1] == 1.0)


def test_compute_ratio_global_pieces():
    m = hubbard_square(4, 4, 1.0, Mz=False)          # SU(2): N_FL = 1, N_SUN = 2, imaginary g, alpha != 0
    C = 3; rng = np.random.default_rng(3)
    f = rng.choice([-2.0, -1.0, 1.0, 2.0], size=(C, m.Ltrot, m.n_opv)).astype(np.complex128)
    ld = rng.normal(size=(C, m.N_FL)); ph = np.exp(1j * rng.normal(size=(C, m.N_FL)))
    r1, r2 = tempering.compute_ratio_global(m, ld, ph, ld, ph, f, f)
    assert np.allclose(r1, 1.0) and np.allclose(r2, 0.0)                                  # nothing changed
    r1, r2 = tempering.compute_ratio_global(m, ld, ph, ld + 0.25, ph * np.exp(0.3j), f, f)
    assert np.allclose(r2, m.N_SUN * 0.25 * m.N_FL) and np.allclose(r1, np.exp(0.3j * m.N_SUN) ** m.N_FL)
    f2 = f.copy(); f2[0, 2, 5] = 2.0 if abs(f[0, 2, 5].real) != 2.0 else 1.0                # one field of chain 0 changes its modulus
    r1, r2 = tempering.compute_ratio_global(m, ld, ph, ld, ph, f, f2)
    op = m.Op_V[5][0]
    s_o, s_n = int(f[0, 2, 5].real) + 2, int(f2[0, 2, 5].real) + 2
    want = tempering.GAMA_ST[2][s_n] / tempering.GAMA_ST[2][s_o] * np.exp(m.N_SUN * (tempering.PHI_ST[2][s_n] - tempering.PHI_ST[2][s_o]) * op.g * op.alpha)
    assert np.allclose(r1[0], want) and np.allclose(r1[1:], 1.0) and np.allclose(r2, 0.0)
    # the weight of a pair is symmetric under exchanging the roles of the two groups
    ra = tempering.compute_ratio_global(m, ld, ph, ld + 0.1, ph, f, f2); rb = tempering.compute_ratio_global(m, ld + 0.1, ph, ld, ph, f2, f)
    assert np.allclose(ra[0] * rb[0], 1.0) and np.allclose(ra[1] + rb[1], 0.0)
