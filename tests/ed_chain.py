"""Exact diagonalisation of the 4-site Hubbard chain of testsuite/test_vs_ed (256 Fock states): the reference's ED anchors, and the energy the
Trotterised, discretely Hubbard-Stratonovich-decoupled path integral of the QMC has at FINITE dtau -- what a correct sweep must reproduce within
its error bar at every dtau, a sharper statement than the reference's dtau^2 extrapolation.  Test infrastructure (numpy / scipy only)."""
import numpy as np
import scipy.linalg as sl


class Chain:
    def __init__(self, L=4, t=1.0, U=4.0):
        self.L, self.t, self.U = L, t, U
        n = 2 * L; dim = 2 ** n; self.dim = dim

        def cdag(i):
            M = np.zeros((dim, dim))
            for s in range(dim):
                if not (s >> i) & 1:
                    M[s | (1 << i), s] = (-1) ** bin(s & ((1 << i) - 1)).count("1")
            return M
        self.C = [cdag(i) for i in range(n)]                       # orbital index = site + L * spin
        num = [c @ c.T for c in self.C]; self.num = num
        Ht = np.zeros((dim, dim))
        for s in range(2):
            for i in range(L):
                j = (i + 1) % L
                if L == 2 and i == 1:
                    continue
                Ht += -t * (self.C[i + L * s] @ self.C[j + L * s].T + self.C[j + L * s] @ self.C[i + L * s].T)
        eye = np.eye(dim)
        self.Ht = Ht
        self.Hv = sum(U * (num[i] - 0.5 * eye) @ (num[i + L] - 0.5 * eye) for i in range(L))
        self.Pot = sum(U * num[i] @ num[i + L] for i in range(L))
        self.O = Ht + self.Pot                                      # Ener = Kin + Pot of Hamiltonian_Hubbard_smod.F90:738-772
        self.A = [np.diag(num[i] - num[i + L]) for i in range(L)]    # n_up - n_down per site (diagonal)

    def ed_finite_t(self, beta):
        E, V = np.linalg.eigh(self.Ht + self.Hv); w = np.exp(-beta * (E - E.min()))
        return float(np.sum(w * np.einsum("ij,ij->j", V, self.O @ V)) / w.sum())

    def ed_ground_state(self, n_up, n_dn):
        L = self.L
        nu = np.diag(sum(self.num[i] for i in range(L))); nd = np.diag(sum(self.num[i + L] for i in range(L)))
        idx = np.where((np.abs(nu - n_up) < 1e-9) & (np.abs(nd - n_dn) < 1e-9))[0]
        E, V = np.linalg.eigh((self.Ht + self.Hv)[np.ix_(idx, idx)]); g = V[:, 0]
        return float(g @ self.O[np.ix_(idx, idx)] @ g)

    def v_slice(self, dtau, discrete_hs=True):
        """e^{-dtau H_V} as the QMC realises it: exactly, or summed over the four values of the type-2 HS field of Predefined_Int_U_MZ
        (Prog/Predefined_Int_mod.F90:107-133, Fields_mod.F90:270-287): prod_i e^{-dtau U/4} (1/4) sum_l gamma_l e^{sqrt(dtau U/2) eta_l (n_up - n_dn)_i}."""
        if not discrete_hs:
            return np.exp(-dtau * np.diag(self.Hv))
        s6 = np.sqrt(6.0); gam = [1 + s6 / 3, 1 - s6 / 3]; eta = [np.sqrt(2 * (3 - s6)), np.sqrt(2 * (3 + s6))]
        lam = np.sqrt(dtau * self.U / 2.0); v = np.ones(self.dim)
        for a in self.A:
            v = v * np.exp(-dtau * self.U / 4.0) * 0.25 * sum(g * (np.exp(lam * e * a) + np.exp(-lam * e * a)) for g, e in zip(gam, eta))
        return v

    def trotter_finite_t(self, beta, dtau, discrete_hs=True):
        """Tr[(e^{-dtau T/2} e^{-dtau V} e^{-dtau T/2})^L O] / Tr[...]: symmetric Trotter decomposition, i.e. B = e^{-dtau T} e^{V} measured on
        Hop_mod_Symm(G) (Symm = .true.)."""
        Lt = int(round(beta / dtau)); Th = sl.expm(-0.5 * dtau * self.Ht)
        B = Th @ np.diag(self.v_slice(dtau, discrete_hs)) @ Th
        rho = np.linalg.matrix_power(B, Lt)
        return float(np.trace(rho @ self.O) / np.trace(rho))


def fit_e0(dtaus, e, de):
    """curve_fit(y0 + a x^2, sigma = dy, absolute_sigma = True) of testsuite/test_vs_ed/analysis.py:22-29 as weighted linear least squares."""
    x2 = np.asarray(dtaus) ** 2; w = 1.0 / np.asarray(de) ** 2
    A = np.stack([np.ones_like(x2), x2], axis=1)
    cov = np.linalg.inv(A.T @ (A * w[:, None]))
    p = cov @ (A.T @ (w * np.asarray(e)))
    return float(p[0]), float(np.sqrt(cov[0, 0]))


def trotter_projector(chain, theta, beta, dtau, wf, discrete_hs=True):
    """The projective algorithm's estimate at finite dtau: |psi_T> = prod_spin prod_k (sum_i wf[i, k] c^dag_{i,spin}) |0>, Thtrot = nint(theta / dtau),
    Ltrot = nint(beta / dtau) + 2 Thtrot slices B = e^{-dtau T} e^{V}; the energy is measured on the slices Thtrot + 1 .. Ltrot - Thtrot
    (main.F90:757-773, 789-802; Hamiltonian_Hubbard_smod.F90:236-239) on the symmetrised Green function, i.e. with the operator sandwiched as
    e^{-dtau T/2} O e^{+dtau T/2} after slice tau."""
    L = chain.L; dim = chain.dim
    th = int(round(theta / dtau)); lt = int(round(beta / dtau)) + 2 * th
    psi = np.zeros(dim, dtype=complex); psi[0] = 1.0
    for s in range(2):
        for k in range(wf.shape[1]):
            op = sum(wf[i, k] * chain.C[i + L * s] for i in range(L)); psi = op @ psi
    Tf = sl.expm(-dtau * chain.Ht); Th = sl.expm(-0.5 * dtau * chain.Ht); Thi = sl.expm(0.5 * dtau * chain.Ht)
    B = Tf @ np.diag(chain.v_slice(dtau, discrete_hs))        # one slice: e^{V} first, then e^{-dtau T}
    Ot = Th @ chain.O @ Thi
    right = [psi]                                             # right[tau] = B^tau |psi_T>
    for _ in range(lt):
        x = B @ right[-1]; right.append(x / np.linalg.norm(x))
    BH = B.conj().T; left = [psi]                             # left[k] = (B^H)^k |psi_T>
    for _ in range(lt):
        x = BH @ left[-1]; left.append(x / np.linalg.norm(x))
    es = []
    for tau in range(th + 1, lt - th + 1):
        l, r = left[lt - tau], right[tau]
        es.append((np.vdot(l, Ot @ r) / np.vdot(l, r)).real)
    return float(np.mean(es))
