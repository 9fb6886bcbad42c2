"""Host-side stand-in for ALF's model plugin surface (L1): ``Operator`` (Op_make / Op_set),
the Bravais lattice, the predefined hoppings (checkerboard / symmetric Trotter) and the
predefined interactions that the shipped Hamiltonians use.

In a real drop-in this layer is ALF's own, unchanged Fortran (``Hamiltonian_main_mod``,
``Operator_mod``, ``Predefined_*``); the Fortran toolchain is absent in this image, so the
tables that the Fortran shim would flatten and hand to the C-ABI (``include/alf_b200.h``)
are produced here with the same conventions.  Nothing in this file is on the hot path.

Reference: Prog/Operator_mod.F90:56-90,193-215,259-473; Libraries/Modules/lattices_v3_mod.F90:88-330;
Prog/Predefined_Hop_mod.F90:257-362,1422-1495,1624-1839; Prog/Predefined_Int_mod.F90:59-219;
Prog/Hamiltonians/Hamiltonian_Hubbard_smod.F90:207-330,477-542; Hamiltonian_Kondo_smod.F90:464-530.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

EPS_SMALL = 1.0e-10          # Libraries/Modules/natural_constants_mod.F90:8-10 (Eps_small)
EPS_MACHINE = np.finfo(np.float64).eps


class HamiltonianError(RuntimeError):
    """Mirrors Terminate_on_error(ERROR_HAMILTONIAN, ...) (Libraries/Modules/runtime_error_mod.F90:56-68)."""


@dataclass
class Operator:
    """Prog/Operator_mod.F90:56-90.  ``P`` is 1-based, as in Fortran."""
    N: int
    O: np.ndarray = None
    P: np.ndarray = None
    g: complex = 0.0
    alpha: complex = 0.0
    type: int = 0
    flip_protocol: int = 1
    N_non_zero: int = 0
    diag: bool = False
    U: Optional[np.ndarray] = None
    E: Optional[np.ndarray] = None
    g_t: Optional[np.ndarray] = None      # time-dependent coupling g_t(1..Ltrot) of an interaction vertex (Operator_mod.F90:66), or None


def Op_make(N: int) -> Operator:
    """Prog/Operator_mod.F90:193-215."""
    op = Operator(N=N)
    op.O = np.zeros((N, N), dtype=np.complex128)
    op.P = np.zeros(N, dtype=np.int32)
    op.N_non_zero = N
    return op


def Op_set(op: Operator) -> None:
    """Prog/Operator_mod.F90:259-473: validate, detect diagonal operators, diagonalise, order the
    non-zero eigenvalues first, normalise det(U)=1.  (The exponential tables E_exp/M_exp are private
    in ALF and are rebuilt on the C side from U, E, g -- Operator_mod.F90:400-470.)"""
    if op.type < 0 or op.type > 4:
        raise HamiltonianError(f"Op_set: Invalid operator type: {op.type}")
    if np.any(op.P < 1):
        raise HamiltonianError("Op_set: Projector index out of bounds")
    N = op.N
    O = op.O
    tol = EPS_MACHINE
    for i in range(N):
        for j in range(i + 1, N):
            dev = abs(O[i, j] - np.conj(O[j, i]))
            if dev > tol * max(abs(O[i, j]), abs(O[j, i]), 1e-30):
                raise HamiltonianError("Op_set: Operator matrix Op%O is not Hermitian.")
        if abs(O[i, i].imag) > tol * max(abs(O[i, i]), 1e-30):
            raise HamiltonianError("Op_set: complex diagonal element (not Hermitian).")
    op.U = np.zeros((N, N), dtype=np.complex128)
    op.E = np.zeros(N, dtype=np.float64)
    if N > 1:
        offdiag = O - np.diag(np.diag(O))
        op.diag = not np.any(offdiag != 0)
        if op.diag:
            op.E[:] = np.real(np.diag(O))
            op.U[:] = np.eye(N)
            op.N_non_zero = N
        else:
            E, U = np.linalg.eigh(O)          # Diag -> ZHEEV, ascending eigenvalues
            npz, nz = 0, 0
            for i in range(N):
                if abs(E[i]) > EPS_SMALL:
                    op.U[:, npz] = U[:, i]
                    op.E[npz] = E[i]
                    npz += 1
                else:
                    op.U[:, N - 1 - nz] = U[:, i]
                    op.E[N - 1 - nz] = E[i]
                    nz += 1
            op.N_non_zero = npz
            Z = np.linalg.det(op.U)
            if abs(Z) < 1e-30:
                raise HamiltonianError("Op_set: Eigenvector matrix has near-zero determinant")
            op.U[:, 0] = op.U[:, 0] / Z        # scale into SU(N)
    else:
        op.E[0] = O[0, 0].real
        op.U[0, 0] = 1.0
        op.N_non_zero = 1
        op.diag = True


# ----------------------------------------------------------------------------- lattice
class Lattice:
    """Square-type Bravais lattice, Libraries/Modules/lattices_v3_mod.F90:88-330 (Make_lattice):
    unit cells inside the Wigner-Seitz cell of (L1_p, L2_p), enumerated i1 outer / i2 inner."""

    def __init__(self, L1: int, L2: int):
        self.L1, self.L2 = L1, L2

        def rng(L):
            # x.L <= L^2/2 + 0  and  x.L >= -L^2/2 + Zero  (upper edge included, lower excluded)
            return [i for i in range(-L, L + 1) if (i * L <= L * L / 2.0 + EPS_SMALL) and (i * L >= -L * L / 2.0 + EPS_SMALL)]

        r1, r2 = rng(L1), rng(L2)
        self.list = []
        self.invlist = {}
        for i1 in r1:
            for i2 in r2:
                self.list.append((i1, i2))
                self.invlist[(i1, i2)] = len(self.list)      # 1-based
        self.N = len(self.list)
        self._r1, self._r2 = r1, r2

    def _wrap(self, i, r, L):
        lo = r[0]
        return (i - lo) % L + lo

    def imj(self, I: int, J: int) -> int:
        """Latt%imj(I, J): the lattice point r_I - r_J folded back (lattices_v3_mod.F90:316-330); 1-based."""
        i1, i2 = self.list[I - 1]; j1, j2 = self.list[J - 1]
        return self.invlist[(self._wrap(i1 - j1, self._r1, self.L1), self._wrap(i2 - j2, self._r2, self.L2))]

    def imj_table(self) -> np.ndarray:
        """imj as an (N, N) int32 array, [I-1, J-1] -> 1-based index of r_I - r_J."""
        t = np.zeros((self.N, self.N), dtype=np.int32)
        for I in range(1, self.N + 1):
            for J in range(1, self.N + 1):
                t[I - 1, J - 1] = self.imj(I, J)
        return t

    def nnlist(self, I: int, n1: int, n2: int) -> int:
        i1, i2 = self.list[I - 1]
        return self.invlist[(self._wrap(i1 + n1, self._r1, self.L1), self._wrap(i2 + n2, self._r2, self.L2))]


# ----------------------------------------------------------------------------- model container
@dataclass
class Model:
    """What ``Hamiltonian_main`` exposes publicly to the sweep (Prog/Hamiltonian_main_mod.F90:181-197)."""
    name: str
    Ndim: int
    N_FL: int
    N_SUN: int
    Ltrot: int
    Dtau: float
    Symm: bool
    Op_V: List[List[Operator]]          # Op_V[n][nf]
    Op_T: List[List[Operator]]          # Op_T[nc][nf]
    latt: Optional[Lattice] = None
    params: dict = field(default_factory=dict)
    # projective algorithm (Prog/Hamiltonian_main_mod.F90:181-197: Projector, Thtrot, WF_L, WF_R); Ltrot already includes 2*Thtrot
    Projector: bool = False
    Thtrot: int = 0
    WF_L: Optional[List[np.ndarray]] = None     # per flavor, (Ndim, N_part) complex
    WF_R: Optional[List[np.ndarray]] = None

    # ham%S0 of Ising actions as tables (alf_b200_set_s0_ising) and main.F90's Propose_S0
    s0_ising: Optional[dict] = None
    propose_s0: bool = False
    # Nt_sequential_start / _end, N_Global_tau (Hamiltonian_main_mod.F90 Overide_global_tau_sampling_parameters) and ham%Global_move_tau as tables
    global_tau: Optional[dict] = None
    s0_gaussian: bool = False           # ham%S0 = exp(-(f'^2 - f^2)/2) of the continuous HS transformation (Hubbard_smod.F90:880-882)
    global_move_tau_ising: Optional[dict] = None

    # List(I1, 1:2) of the Hamiltonians (unit cell, orbital) per site, 1-based, and the number of orbitals per unit cell
    # (Prog/Predefined_Latt_mod.F90:250-260); None -> one orbital per cell, site I1 = cell I1
    site_cell: Optional[np.ndarray] = None
    site_orb: Optional[np.ndarray] = None
    n_orb: int = 1
    # Latt_unit%N_coord and Latt_unit%Orb_pos_p (Prog/Predefined_Latt_mod.F90:117-237); they only enter the `_info` files of the lattice bins
    n_coord: int = 2
    orb_pos: Optional[np.ndarray] = None

    def lattice_tables(self):
        """(n_unit, n_orb, site_cell, site_orb, imj) as the C-ABI's alf_b200_set_lattice wants them (all 1-based)."""
        n_unit = self.latt.N
        cell = self.site_cell if self.site_cell is not None else np.arange(1, self.Ndim + 1, dtype=np.int32)
        orb = self.site_orb if self.site_orb is not None else np.ones(self.Ndim, dtype=np.int32)
        return n_unit, int(self.n_orb), np.ascontiguousarray(cell, dtype=np.int32), np.ascontiguousarray(orb, dtype=np.int32), self.latt.imj_table()

    @property
    def N_part(self):
        return 0 if not self.Projector else int(self.WF_L[0].shape[1])

    @property
    def n_opv(self):
        return len(self.Op_V)

    @property
    def n_opt(self):
        return len(self.Op_T)


def _square_hopping_families(latt: Lattice, symm: bool):
    """Set_Default_hopping_parameters_square (Predefined_Hop_mod.F90:331-360) + Symmetrize_Families (:1422-1495).
    Returns list of (family members [(I, bond)], Prop_Fam)."""
    fam = [[], [], [], []]
    for I in range(1, latt.N + 1):
        i1, i2 = latt.list[I - 1]
        if (i1 + i2) % 2 == 0:
            fam[0].append((I, 1)); fam[1].append((I, 2))
        else:
            fam[2].append((I, 1)); fam[3].append((I, 2))
    prop = [1.0] * 4
    if not symm:
        return list(zip(fam, prop))
    nfam_c = 4
    order = list(range(nfam_c)) + list(range(nfam_c - 2, -1, -1))
    lens = [len(f) for f in fam]
    n_f_max = int(np.argmax(lens))            # first longest family (strict > in the reference)
    if n_f_max != nfam_c - 1:
        order[nfam_c - 1] = n_f_max
        order[0] = nfam_c - 1
        order[-1] = nfam_c - 1
    props = [0.5] * (2 * nfam_c - 1)
    props[nfam_c - 1] = 1.0
    return [(fam[k], props[i]) for i, k in enumerate(order)]


def trial_wave_function_square(latt: Lattice, n_part: int, N_FL: int, kind: str = "flux", t: float = 1.0):
    """Predefined_TrialWaveFunction for the square lattice (Prog/Predefined_Trial_mod.F90:112-398): the N_part lowest
    single-particle states of a non-interacting H0 whose degeneracy at the Fermi level is lifted.
    kind = "flux": ALF's choice, a twist Phi_X = 0.01 of the boundary condition along L1 (complex orbitals);
    kind = "dimer": a real alternative (bond dimerisation delta = 0.01 along L1) that keeps the sweep in real arithmetic.
    Returns (WF_L, WF_R, degeneracy gap E(N_part+1) - E(N_part))."""
    N = latt.N
    H = np.zeros((N, N), dtype=np.complex128)
    for I in range(1, N + 1):
        i1, i2 = latt.list[I - 1]
        for (n1, n2) in ((0, 1), (1, 0)):
            J = latt.nnlist(I, n1, n2)
            amp = -t
            if n1 == 1:
                if kind == "flux":
                    amp = -t * np.exp(2j * np.pi * 0.01 / latt.L1)
                elif kind == "dimer":
                    amp = -t * (1.0 + 0.01 * (-1) ** (i1 % 2))
            H[I - 1, J - 1] += amp
            H[J - 1, I - 1] += np.conj(amp)
    E, Uv = np.linalg.eigh(H)
    P = np.asfortranarray(Uv[:, :n_part])
    if kind == "dimer":
        P = np.asfortranarray(P.real.astype(np.complex128))
    return [P.copy(order="F") for _ in range(N_FL)], [P.copy(order="F") for _ in range(N_FL)], float(E[n_part] - E[n_part - 1])


def hubbard_square(L1: int, L2: int, beta: float, dtau: float = 0.1, U: float = 4.0, t: float = 1.0, mu: float = 0.0,
                   Mz: bool = True, checkerboard: bool = True, symm: bool = True, N_SUN: int = 2,
                   projector: bool = False, theta: float = 10.0, trial: str = "flux", continuous: bool = False) -> Model:
    """Hubbard model on the square lattice as set up by Hamiltonian_Hubbard_smod.F90 (Ham_Set :207-330,
    Ham_Hop, Ham_V :477-542) with the shipped defaults (Scripts_and_Parameters_files/Start/parameters).
    projector=True: the projective algorithm (:236-239 Thtrot = nint(theta/dtau), Ltrot += 2 Thtrot; Ham_Trial :455-470,
    N_part = Ndim/2)."""
    latt = Lattice(L1, L2)
    Ndim = latt.N
    Ltrot = int(round(beta / dtau))
    Thtrot = 0
    if projector:
        Thtrot = int(round(theta / dtau))
        Ltrot = Ltrot + 2 * Thtrot
    N_FL = 2 if Mz else 1
    n_sun = N_SUN // 2 if N_FL == 2 else N_SUN
    # bonds: (no_I, no_J, n_1, n_2) ; List(1) = (1,1,0,1), List(2) = (1,1,1,0) ; T = -t ; T_loc = -mu
    bonds = {1: (0, 1), 2: (1, 0)}
    Op_T: List[List[Operator]] = []
    if checkerboard:
        fams = _square_hopping_families(latt, symm)
        mult = 4                                   # Multiplicity of orbital 1 (Predefined_Hop_mod.F90:1773-1781)
        for members, prop in fams:
            for (I, nb) in members:
                n1, n2 = bonds[nb]
                J = latt.nnlist(I, n1, n2)
                row = []
                for nf in range(N_FL):
                    op = Op_make(2)
                    op.P[0], op.P[1] = I, J
                    op.O[0, 1] = -t
                    op.O[1, 0] = -t
                    op.O[0, 0] = -mu / mult
                    op.O[1, 1] = -mu / mult
                    op.g = -dtau * prop
                    Op_set(op)
                    row.append(op)
                Op_T.append(row)
    else:
        row = []
        for nf in range(N_FL):
            op = Op_make(Ndim)
            for I in range(1, latt.N + 1):
                for nb in (1, 2):
                    n1, n2 = bonds[nb]
                    J = latt.nnlist(I, n1, n2)
                    op.O[I - 1, J - 1] = -t
                    op.O[J - 1, I - 1] = -t
                op.O[I - 1, I - 1] = -mu
            op.P[:] = np.arange(1, Ndim + 1)
            op.g = -dtau
            Op_set(op)
            row.append(op)
        Op_T.append(row)
    Op_V: List[List[Operator]] = []
    if abs(U) > EPS_SMALL:
        for I in range(1, latt.N + 1):
            if Mz and continuous:
                # Predefined_Int_U_MZ_continuous_HS (Predefined_Int_mod.F90:160-181), called with Ham_U_vec/N_SUN (Hubbard_smod.F90:510-512)
                ops = []
                for sgn in (+1.0, -1.0):
                    op = Op_make(1); op.P[0] = I; op.O[0, 0] = 1.0; op.alpha = 0.0
                    op.g = sgn * np.sqrt(complex(dtau * U / float(n_sun), 0.0)); op.type = 3
                    Op_set(op); ops.append(op)
                Op_V.append(ops)
            elif continuous:
                # Predefined_Int_U_SUN_continuous_HS (Predefined_Int_mod.F90:90-105)
                op = Op_make(1); op.P[0] = I; op.O[0, 0] = 1.0; op.alpha = -0.5
                op.g = np.sqrt(complex(-dtau * U * 2.0 / float(N_FL * n_sun), 0.0)); op.type = 3
                Op_set(op); Op_V.append([op])
            elif Mz:
                # Predefined_Int_U_MZ (Predefined_Int_mod.F90:107-133); Ham_U_vec/N_SUN with N_SUN already halved
                Ueff = U / float(n_sun)
                ops = []
                for sgn in (+1.0, -1.0):
                    op = Op_make(1)
                    op.P[0] = I
                    op.O[0, 0] = 1.0
                    op.g = sgn * np.sqrt(complex(dtau * Ueff / 2.0, 0.0))
                    op.alpha = -0.5
                    op.type = 2
                    Op_set(op)
                    ops.append(op)
                Op_V.append(ops)
            else:
                # Predefined_Int_U_SUN (Predefined_Int_mod.F90:59-77)
                op = Op_make(1)
                op.P[0] = I
                op.O[0, 0] = 1.0
                op.alpha = -0.5
                op.g = np.sqrt(complex(-dtau * U / float(N_FL * n_sun), 0.0))
                op.type = 2
                Op_set(op)
                Op_V.append([op])
    m = Model(name="Hubbard", Ndim=Ndim, N_FL=N_FL, N_SUN=n_sun, Ltrot=Ltrot, Dtau=dtau, Symm=bool(symm and checkerboard),
              Op_V=Op_V, Op_T=Op_T, latt=latt,
              params=dict(L1=L1, L2=L2, beta=beta, dtau=dtau, U=U, t=t, mu=mu, Mz=Mz, checkerboard=checkerboard, symm=symm,
                          projector=projector, theta=theta, trial=trial))
    m.s0_gaussian = bool(continuous)
    if projector:
        m.Projector, m.Thtrot = True, Thtrot
        m.WF_L, m.WF_R, m.params["wf_degen"] = trial_wave_function_square(latt, Ndim // 2, N_FL, trial, t)
    return m


def trial_wave_function_chain(L: int, n_part: int, N_FL: int, t: float = 1.0):
    """Predefined_TrialWaveFunction for Lattice_type = N_leg_ladder with one leg (Prog/Predefined_Trial_mod.F90:319-325, 352-364): the N_part lowest
    eigenvectors of the non-interacting ring with Ham_T = 1 and a twist Phi_X = 0.01 applied as a BOUNDARY condition (Bulk = .false.: only the bond
    that crosses the boundary carries exp(2 pi i Phi_X N1), Generic_hopping, Prog/Predefined_Hop_mod.F90:2105-2115), which lifts the degeneracy at
    half filling.  Returns (WF_L, WF_R, E(N_part+1) - E(N_part))."""
    H = np.zeros((L, L), dtype=np.complex128)
    for I in range(L):
        J = (I + 1) % L
        if L == 2 and I == 1:
            continue
        z = -t * (np.exp(2j * np.pi * 0.01) if J < I else 1.0)
        H[I, J] += z
        H[J, I] += np.conj(z)
    E, Uv = np.linalg.eigh(H)
    P = np.asfortranarray(Uv[:, :n_part])
    return [P.copy(order="F") for _ in range(N_FL)], [P.copy(order="F") for _ in range(N_FL)], float(E[n_part] - E[n_part - 1])


def hubbard_chain(L: int, beta: float, dtau: float, U: float = 4.0, t: float = 1.0, Mz: bool = True, symm: bool = True,
                  projector: bool = False, theta: float = 5.0) -> Model:
    """N_leg_ladder with L2=1 and Checkerboard=.false. as in testsuite/test_vs_ed/test_specs.yaml:24-94
    (dense hopping matrix, periodic chain; with Symm the half-step similarity is applied before measuring).
    projector=True: Hamiltonian_Hubbard_smod.F90:236-239 (Thtrot = nint(theta/dtau), Ltrot += 2 Thtrot) with the trial wave function of Ham_Trial
    (N_part = Ndim/2)."""
    Ndim = L
    Ltrot = int(round(beta / dtau))
    Thtrot = 0
    if projector:
        Thtrot = int(round(theta / dtau)); Ltrot += 2 * Thtrot
    N_FL = 2 if Mz else 1
    row = []
    for nf in range(N_FL):
        op = Op_make(Ndim)
        for I in range(L):
            J = (I + 1) % L
            if L == 2 and I == 1:
                continue
            op.O[I, J] += -t
            op.O[J, I] += -t
        op.P[:] = np.arange(1, Ndim + 1)
        op.g = -dtau
        Op_set(op)
        row.append(op)
    Op_V = []
    for I in range(1, L + 1):
        if Mz:
            ops = []
            for sgn in (+1.0, -1.0):
                op = Op_make(1); op.P[0] = I; op.O[0, 0] = 1.0
                op.g = sgn * np.sqrt(complex(dtau * U / 2.0, 0.0)); op.alpha = -0.5; op.type = 2
                Op_set(op); ops.append(op)
            Op_V.append(ops)
        else:
            op = Op_make(1); op.P[0] = I; op.O[0, 0] = 1.0; op.alpha = -0.5
            op.g = np.sqrt(complex(-dtau * U / 2.0, 0.0)); op.type = 2
            Op_set(op); Op_V.append([op])
    m = Model(name="Hubbard_chain", Ndim=Ndim, N_FL=N_FL, N_SUN=1 if Mz else 2, Ltrot=Ltrot, Dtau=dtau, Symm=symm,
              Op_V=Op_V, Op_T=[row], params=dict(L=L, beta=beta, dtau=dtau, U=U, t=t, Mz=Mz, symm=symm, projector=projector, theta=theta))
    if projector:
        m.Projector, m.Thtrot = True, Thtrot
        m.WF_L, m.WF_R, m.params["wf_degen"] = trial_wave_function_chain(L, Ndim // 2, N_FL, t)
    return m


def kondo_square(L1: int, L2: int, beta: float, dtau: float = 0.1, t: float = 1.0, J: float = 2.0, Uf: float = 1.0,
                 Uc: float = 0.0, symm: bool = True, N_SUN: int = 2, jk_vertex: str = "hop") -> Model:
    """SU(N) Kondo lattice on the bilayer square lattice, Hamiltonian_Kondo_smod.F90:464-530: orbital 1 = conduction
    (hops), orbital 2 = f (no hopping); vertices U_f (k=1, imaginary g: Predefined_Int_U_SUN, Ham_V :497-505) and
    J_K (k=2 vertex, Predefined_Int_V_SUN with Ham_JK/2 for N_SUN=2, :507-514).  The conduction-layer hopping uses the
    checkerboard families of Set_Default_hopping_parameters_Bilayer_square for Ham_T2 = Ham_Tperp = 0
    (Predefined_Hop_mod.F90:1126-1175): bond 1 = (0, 1), bond 2 = (1, 0) between conduction orbitals, families 1 / 2 = bonds 1 / 2 of
    the unit cells with even i1 + i2, families 3 / 4 = those of the odd ones -- the same grouping and order as the square lattice
    (:336-360), so _square_hopping_families serves both; Symmetrize_Families (:1422-1495) is applied on top of it."""
    latt = Lattice(L1, L2)
    Nc = latt.N
    Ndim = 2 * Nc                                  # Invlist(I,no) = (I-1)*Norb + no (Predefined_Latt_mod.F90:250-260)
    Ltrot = int(round(beta / dtau))
    inv = lambda I, no: (I - 1) * 2 + no
    bonds = {1: (0, 1), 2: (1, 0)}
    fams = _square_hopping_families(latt, symm)
    Op_T = []
    for members, prop in fams:
        for (I, nb) in members:
            n1, n2 = bonds[nb]
            Jn = latt.nnlist(I, n1, n2)
            op = Op_make(2)
            op.P[0], op.P[1] = inv(I, 1), inv(Jn, 1)
            op.O[0, 1] = -t; op.O[1, 0] = -t
            op.g = -dtau * prop
            Op_set(op)
            Op_T.append([op])
    Op_V = []
    for I in range(1, Nc + 1):                    # U_f on the f orbital
        if abs(Uf) > EPS_SMALL:
            op = Op_make(1); op.P[0] = inv(I, 2); op.O[0, 0] = 1.0; op.alpha = -0.5
            op.g = np.sqrt(complex(-dtau * Uf / float(N_SUN), 0.0)); op.type = 2
            Op_set(op); Op_V.append([op])
    for I in range(1, Nc + 1):                    # N_SUN == 2: Predefined_Int_V_SUN(OP, I1, I2, N_SUN, DTAU, Ham_JK/2)
        op = Op_make(2)
        op.P[0], op.P[1] = inv(I, 1), inv(I, 2)
        op.O[0, 1] = 1.0; op.O[1, 0] = 1.0
        if jk_vertex == "bond_density":          # test variant: (c^dag + f^dag)(c + f), eigenvalues (2, 0): a k = 2 vertex with ONE non-zero eigenvalue
            op.O[0, 0] = 1.0; op.O[1, 1] = 1.0
        op.g = np.sqrt(complex(dtau * (J / 2.0) / float(N_SUN), 0.0)); op.alpha = 0.0; op.type = 2
        Op_set(op); Op_V.append([op])
    m = Model(name="Kondo", Ndim=Ndim, N_FL=1, N_SUN=N_SUN, Ltrot=Ltrot, Dtau=dtau, Symm=symm, Op_V=Op_V, Op_T=Op_T, latt=latt,
              params=dict(L1=L1, L2=L2, beta=beta, dtau=dtau, t=t, J=J, Uf=Uf, symm=symm))
    m.n_orb = 2
    m.n_coord = 2; m.orb_pos = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, -1.0]])      # Bilayer_square: Orb_pos_p(no, 3) = 1 - no
    m.site_cell = np.repeat(np.arange(1, Nc + 1, dtype=np.int32), 2)
    m.site_orb = np.tile(np.array([1, 2], dtype=np.int32), Nc)
    return m


def z2_matter_square(L1: int, L2: int, beta: float, dtau: float = 0.1, t: float = 1.0, t_z2: float = 1.0, chem: float = 0.0, U: float = 0.0,
                     J: float = 1.0, K: float = 1.0, h: float = 1.0, g: float = 1.0, N_SUN: int = 2, propose_s0: bool = False,
                     projector: bool = False, theta: float = 10.0, n_part: int = -1) -> Model:
    """Hamiltonian_Z2_Matter_smod.F90 (Z2 lattice gauge theory coupled to fermions and Z2 matter on the square lattice) as Ham_Set builds it
    (:143-240) with the shipped defaults of Scripts_and_Parameters_files/Start/parameters:156-167.
      * Fields (Setup_Ising_action_and_field_list :738-832): [Hubbard, one per site, if U] + [Z2 gauge bonds, if t_z2] + [matter bonds and ONE
        site matter field at site Latt%N, if t]; bonds are enumerated per site I of even Ix + Iy as (I,x), (I,y), (I - a_x, x), (I - a_y, y).
      * Ham_V (:313-371): Predefined_Int_Ising_SUN bond vertices (type 1, k = 2) with Xi = -t_z2 / -t; the site matter field is a k = 1 operator
        with g = 0 (no fermionic weight); Ham_Hop (:274-296): one dense Op_T carrying only the chemical potential.
      * Ising action as tables: ham%S0 (:439-512) of the gauge fields = coupling to the matter bond field (DW_Ising_Matter), to the two
        neighbouring time slices (DW_Ising_tau; only the existing neighbour(s) in the projective algorithm) and the two plaquettes
        (DW_Ising_Flux); matter fields are not visited sequentially (Overide_global_tau_sampling_parameters :1330-1343) but through
        N_Global_tau = Latt%N/4 star moves per slice, ham%Global_move_tau (:535-643): the four matter bonds around a random site (+ the site
        field for I = Latt%N), S0_Matter = prod DW_Ising_Matter * DW_Matter_tau(tau_I(nt) tau_I(nt +- 1)) where the site variable tau_I is the
        product of the bond fields along the walk of Hamiltonian_set_Z2_matter (:1288-1317) -- again a product of fields, hence a table term.
      * projector=True: Thtrot = nint(theta/dtau), Ltrot += 2 Thtrot, N_part = L1 L2 / 2, trial wave function of Ham_Trial (:380-425)."""
    latt = Lattice(L1, L2)
    if L1 == 1 or L2 == 1:
        raise HamiltonianError("Ham_Latt: One dimensional systems are not included")
    if abs(t) < EPS_SMALL:
        J = 0.0; h = 0.0
    if abs(t_z2) < EPS_SMALL:
        J = 0.0; K = 0.0; g = 0.0
    Ndim = latt.N
    Ltrot = int(round(beta / dtau))
    Thtrot = int(round(theta / dtau)) if projector else 0
    Ltrot += 2 * Thtrot
    op = Op_make(Ndim)
    for I in range(1, Ndim + 1):
        op.O[I - 1, I - 1] = -chem
    op.P[:] = np.arange(1, Ndim + 1); op.g = -dtau
    Op_set(op)
    Op_T = [[op]]
    field_list = {}                 # (site, orientation, type) -> field index (1-based)
    field_inv = []                  # field -> (site, orientation, type)

    def add(I, no, ty):
        field_inv.append((I, no, ty)); field_list[(I, no, ty)] = len(field_inv)
    if abs(U) > EPS_SMALL:
        for I in range(1, latt.N + 1):
            add(I, 3, 3)
    for ty, amp in ((1, t_z2), (2, t)):
        if abs(amp) > EPS_SMALL:
            for I in range(1, latt.N + 1):
                ix, iy = latt.list[I - 1]
                if (ix + iy) % 2 == 0:
                    for (I1, no) in ((I, 1), (I, 2), (latt.nnlist(I, -1, 0), 1), (latt.nnlist(I, 0, -1), 2)):
                        add(I1, no, ty)
    if abs(t) > EPS_SMALL:
        add(latt.N, 3, 4)
    Op_V = []
    I1 = 1
    for (I, no, ty) in field_inv:
        if ty == 3:                # Predefined_Int_U_SUN (Predefined_Int_mod.F90:59-77)
            o = Op_make(1); o.P[0] = I; o.O[0, 0] = 1.0; o.alpha = -0.5
            o.g = np.sqrt(complex(-dtau * U / float(N_SUN), 0.0)); o.type = 2
        elif ty in (1, 2):         # Predefined_Int_Ising_SUN(OP, I, I1, DTAU, -Ham_TZ2 | -Ham_T) (Predefined_Int_mod.F90:264-283)
            I1 = latt.nnlist(I, 1, 0) if no == 1 else latt.nnlist(I, 0, 1)
            o = Op_make(2); o.P[0], o.P[1] = I, I1; o.O[0, 1] = 1.0; o.O[1, 0] = 1.0
            o.g = complex(-dtau * (-(t_z2 if ty == 1 else t)), 0.0); o.alpha = 0.0; o.type = 1
        else:                      # site matter field (:358-367): P(1) = I1 is whatever the previous bond left there; g = 0
            o = Op_make(1); o.P[0] = I1; o.O[0, 0] = 1.0; o.g = 0.0; o.alpha = 0.0; o.type = 1
        Op_set(o); Op_V.append([o])
    # ---- tables (product -1, product +1)
    dw_tau = (1.0 / np.tanh(dtau * g), float(np.tanh(dtau * g))) if g > EPS_SMALL else (1.0, 1.0)
    dw_flux = (float(np.exp(-2.0 * dtau * K)), float(np.exp(2.0 * dtau * K)))
    dw_im = (float(np.exp(-2.0 * dtau * J)), float(np.exp(2.0 * dtau * J)))
    dw_mtau = (1.0 / np.tanh(dtau * h), float(np.tanh(dtau * h))) if h > EPS_SMALL else (1.0, 1.0)
    FL = lambda I, no, ty=1: field_list[(I, no, ty)]

    class Terms:
        def __init__(self):
            self.owner_start, self.term_start, self.e_op, self.e_dt, self.w = [0], [0], [], [], []

        def term(self, entries, tab):
            for (m_, dt) in entries:
                self.e_op.append(m_); self.e_dt.append(dt)
            self.term_start.append(len(self.e_op)); self.w.extend(tab)

        def close_owner(self):
            self.owner_start.append(len(self.term_start) - 1)

        def table(self, open_bc):
            return dict(n_terms=len(self.term_start) - 1, op_start=np.array(self.owner_start, np.int32), term_start=np.array(self.term_start, np.int32),
                        e_op=np.array(self.e_op if self.e_op else [1], np.int32), e_dt=np.array(self.e_dt if self.e_dt else [0], np.int32),
                        w=np.array(self.w if self.w else [1.0, 1.0], np.float64), open_bc=int(open_bc))
    s0 = Terms()
    for n, (I1, no, ty) in enumerate(field_inv, start=1):
        if ty == 1:
            if abs(t) > EPS_SMALL:
                s0.term([(n, 0), (FL(I1, no, 2), 0)], dw_im)
            s0.term([(n, 0), (n, 1)], dw_tau); s0.term([(n, 0), (n, -1)], dw_tau)
            if no == 1:
                I2, I3 = latt.nnlist(I1, 0, 1), latt.nnlist(I1, 1, 0)
                s0.term([(n, 0), (FL(I1, 2), 0), (FL(I2, 1), 0), (FL(I3, 2), 0)], dw_flux)
                I2, I3 = latt.nnlist(I1, 0, -1), latt.nnlist(I1, 1, -1)
                s0.term([(n, 0), (FL(I2, 1), 0), (FL(I2, 2), 0), (FL(I3, 2), 0)], dw_flux)
            else:
                I2, I3 = latt.nnlist(I1, -1, 0), latt.nnlist(I1, -1, 1)
                s0.term([(n, 0), (FL(I2, 1), 0), (FL(I2, 2), 0), (FL(I3, 1), 0)], dw_flux)
                I2, I3 = latt.nnlist(I1, 0, 1), latt.nnlist(I1, 1, 0)
                s0.term([(n, 0), (FL(I1, 1), 0), (FL(I2, 1), 0), (FL(I3, 2), 0)], dw_flux)
        s0.close_owner()
    m = Model(name="Z2_Matter", Ndim=Ndim, N_FL=1, N_SUN=N_SUN, Ltrot=Ltrot, Dtau=dtau, Symm=False, Op_V=Op_V, Op_T=Op_T, latt=latt,
              params=dict(L1=L1, L2=L2, beta=beta, dtau=dtau, t=t, t_z2=t_z2, g=g, K=K, J=J, h=h, chem=chem, U=U, projector=projector, theta=theta))
    if abs(t_z2) > EPS_SMALL:
        m.s0_ising = s0.table(projector)
    m.propose_s0 = bool(propose_s0)
    m.params["field_inv"] = field_inv
    # ---- Overide_global_tau_sampling_parameters (:1330-1343) and the star moves
    nt_seq_end = (latt.N if abs(U) > EPS_SMALL else 0) + (2 * latt.N if abs(t_z2) > EPS_SMALL else 0)
    m.global_tau = dict(nt_seq_start=1, nt_seq_end=nt_seq_end, n_global_tau=(latt.N // 4 if abs(t) > EPS_SMALL else 0))
    if abs(t) > EPS_SMALL:
        # tau_I as a set of fields: the walk of Hamiltonian_set_Z2_matter (:1303-1315), later assignments overwrite earlier ones
        site_f = FL(latt.N, 3, 4)
        sets = {latt.N: frozenset([site_f])}
        I = latt.N
        for _nx in range(L1):
            for _ny in range(L2):
                I1 = latt.nnlist(I, 0, 1); sets[I1] = sets[I] ^ frozenset([FL(I, 2, 2)]); I = I1
            I1 = latt.nnlist(I, 1, 0); sets[I1] = sets[I] ^ frozenset([FL(I, 1, 2)]); I = I1
        assert len(sets) == latt.N
        gm = Terms(); move_start, move_fields = [0], []
        for I in range(1, latt.N + 1):
            bonds = [(I, 1), (I, 2), (latt.nnlist(I, -1, 0), 1), (latt.nnlist(I, 0, -1), 2)]
            fl = [FL(b[0], b[1], 2) for b in bonds]
            if abs(t_z2) > EPS_SMALL:
                for b, n_op in zip(bonds, fl):
                    gm.term([(n_op, 0), (FL(b[0], b[1], 1), 0)], dw_im)
            path = sorted(sets[I])
            gm.term([(f_, 0) for f_ in path] + [(f_, 1) for f_ in path], dw_mtau)       # tau_I(nt) tau_I(nt + 1)
            gm.term([(f_, 0) for f_ in path] + [(f_, -1) for f_ in path], dw_mtau)      # tau_I(nt) tau_I(nt - 1)
            gm.close_owner()
            if I == latt.N:
                fl = fl + [site_f]
            move_fields.extend(sorted(fl)); move_start.append(len(move_fields))          # Wrapgr_sort: ascending field index
        tb = gm.table(projector)
        tb.update(n_sites=latt.N, move_start=np.array(move_start, np.int32), move_fields=np.array(move_fields, np.int32))
        m.global_move_tau_ising = tb
        m.params["tau_sets"] = sets
    if projector:
        npart = (L1 * L2) // 2 if n_part < 0 else n_part
        H0 = np.zeros((Ndim, Ndim)); Delta = 0.01
        for I in range(1, latt.N + 1):
            ix, iy = latt.list[I - 1]
            Ix = latt.nnlist(I, 1, 0)
            H0[I - 1, Ix - 1] = H0[Ix - 1, I - 1] = -(1.0 + Delta * np.cos(np.pi * float(ix + iy)))
            if L2 > 1:
                Iy = latt.nnlist(I, 0, 1)
                H0[I - 1, Iy - 1] = H0[Iy - 1, I - 1] = -(1.0 - Delta)
        E0, U0 = np.linalg.eigh(H0)
        P = np.asfortranarray(U0[:, :npart].astype(np.complex128))
        m.Projector, m.Thtrot = True, Thtrot
        m.WF_L, m.WF_R = [P.copy()], [P.copy()]
        m.params["wf_degen"] = float(E0[npart] - E0[npart - 1])
    return m


def z2_gauge_square(L1: int, L2: int, beta: float, dtau: float = 0.1, t_z2: float = 1.0, g: float = 1.0, K: float = 1.0, chem: float = 0.0,
                    U: float = 0.0, N_SUN: int = 2, propose_s0: bool = False) -> Model:
    """The Z2-gauge sector of Hamiltonian_Z2_Matter (Ham_T = 0, hence Ham_J = Ham_h = 0, :163-166): only the Ising bond fields the fermions
    hop through (+ optional Hubbard vertices); all fields are visited sequentially, N_Global_tau = 0."""
    return z2_matter_square(L1, L2, beta, dtau, t=0.0, t_z2=t_z2, chem=chem, U=U, J=0.0, K=K, h=0.0, g=g, N_SUN=N_SUN, propose_s0=propose_s0)


def obs_scal_tables(model: Model) -> dict:
    """Tables for alf_b200_set_obs_scal_tables: what ham%Obser of the Hubbard Hamiltonian accumulates as Kin, Pot, Ener
    (Prog/Hamiltonians/Hamiltonian_Hubbard_smod.F90:738-772).  Kin = N_SUN sum_{nf,i,j} T(i,j,nf) GRC(i,j,nf) with the hopping matrix
    T = sum_nc Op_T(nc)%O g_nc / (-dtau) (Predefined_Hoppings_Compute_Kin, Predefined_Hop_mod.F90:1850-1959: per bond Z GRC(I1,J1) + conj(Z) GRC(J1,I1),
    per site T_Loc GRC(I1,I1); the checkerboard / symmetric decomposition only splits the same matrix over several Op_T);
    Pot = ham_U sum_i GRC(i,i,1) GRC(i,i,dec), dec = 2 for Mz (N_FL = 2) else 1.  Indices 1-based."""
    N, F = model.Ndim, model.N_FL
    ki, kj, kf, kc = [], [], [], []
    for nf in range(F):
        Tm = np.zeros((N, N), dtype=np.complex128)
        for row in model.Op_T:
            op = row[nf]
            P = np.asarray(op.P) - 1
            Tm[np.ix_(P, P)] += op.O * (op.g / (-model.Dtau))
        for i, j in zip(*np.nonzero(np.abs(Tm) > 1e-14)):
            ki.append(i + 1); kj.append(j + 1); kf.append(nf + 1); kc.append(Tm[i, j])
    p1, f1, p2, f2, pc = [], [], [], [], []
    U = model.params.get("U", 0.0) if model.name.startswith("Hubbard") else 0.0
    if abs(U) > EPS_SMALL:
        dec = 2 if F == 2 else 1
        for i in range(1, N + 1):
            p1.append(i); f1.append(1); p2.append(i); f2.append(dec); pc.append(complex(U))
    return dict(kin_i=ki, kin_j=kj, kin_nf=kf, kin_coef=kc, pot_i1=p1, pot_nf1=f1, pot_i2=p2, pot_nf2=f2, pot_coef=pc)


def flatten_ops(model: Model):
    """Flattened operator tables, one entry per Fortran-shim call into the C-ABI
    (alf_b200_set_op_v / alf_b200_set_op_t)."""
    out_v, out_t = [], []
    for n, row in enumerate(model.Op_V):
        for nf, op in enumerate(row):
            out_v.append(dict(n=n + 1, nf=nf + 1, N=op.N, nnz=op.N_non_zero, diag=int(op.diag), type=op.type,
                              P=np.ascontiguousarray(op.P, dtype=np.int32),
                              U=np.asfortranarray(op.U, dtype=np.complex128),
                              E=np.ascontiguousarray(op.E, dtype=np.float64), g=complex(op.g), alpha=complex(op.alpha),
                              g_t=None if op.g_t is None else np.ascontiguousarray(op.g_t, dtype=np.complex128)))
    for nc, row in enumerate(model.Op_T):
        for nf, op in enumerate(row):
            out_t.append(dict(nc=nc + 1, nf=nf + 1, N=op.N, diag=int(op.diag),
                              P=np.ascontiguousarray(op.P, dtype=np.int32),
                              U=np.asfortranarray(op.U, dtype=np.complex128),
                              E=np.ascontiguousarray(op.E, dtype=np.float64), g=complex(op.g)))
    return out_v, out_t
