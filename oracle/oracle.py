"""ctypes binding of the CPU oracle (oracle/alf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (alf_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "alf_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_ranf.restype = C.c_double
        _lib.orc_s0.restype = C.c_double
        _lib.orc_langevin_update.restype = C.c_double
        _lib.orc_global_move_s0.restype = C.c_double
        _lib.orc_get_gm_log.restype = C.c_long
        _lib.orc_get_log.restype = C.c_long
        _lib.orc_taum_get.restype = C.c_long
        _lib.orc_taum_fresh_get.restype = C.c_long
        _lib.orc_eq_get.restype = C.c_long
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _cplx(a):
    return np.asfortranarray(a, dtype=np.complex128)


class Oracle:
    """One Markov chain of the reference algorithm (one MPI rank of ALF)."""

    def __init__(self, model, nwrap: int = 10, stab3: bool = False):
        from alf_b200.model import flatten_ops
        L = lib()
        self.m = model
        self.N = model.Ndim
        self.nwrap = nwrap
        self.h = C.c_void_p(L.orc_create(model.Ndim, model.N_FL, model.N_SUN, model.Ltrot, nwrap, model.n_opv, model.n_opt,
                                         int(model.Symm), int(stab3)))
        ov, ot = flatten_ops(model)
        for o in ov:
            L.orc_set_op_v(self.h, o["n"], o["nf"], o["N"], o["nnz"], o["diag"], o["type"], o["P"].ctypes.data_as(_ip),
                           _d(o["U"]), _d(o["E"]), C.c_double(o["g"].real), C.c_double(o["g"].imag),
                           C.c_double(o["alpha"].real), C.c_double(o["alpha"].imag))
            if o.get("g_t") is not None:
                assert o["g_t"].size == model.Ltrot
                L.orc_set_op_v_gt(self.h, o["n"], o["nf"], _d(o["g_t"]))
        for o in ot:
            L.orc_set_op_t(self.h, o["nc"], o["nf"], o["N"], o["diag"], o["P"].ctypes.data_as(_ip), _d(o["U"]), _d(o["E"]),
                           C.c_double(o["g"].real), C.c_double(o["g"].imag))

        if getattr(model, "Projector", False):
            L.orc_set_projector(self.h, int(model.Thtrot), int(model.N_part))
            for nf in range(model.N_FL):
                L.orc_set_trial_wf(self.h, nf + 1, _d(_cplx(model.WF_L[nf])), _d(_cplx(model.WF_R[nf])))

        s0 = getattr(model, "s0_ising", None)
        if s0 is not None or getattr(model, "propose_s0", False):
            t = s0 if s0 is not None else dict(n_terms=0, op_start=np.zeros(model.n_opv + 1, np.int32), term_start=np.zeros(1, np.int32),
                                               e_op=np.zeros(1, np.int32), e_dt=np.zeros(1, np.int32), w=np.zeros(2), open_bc=0)
            arr = {k: np.ascontiguousarray(t[k], dtype=np.int32) for k in ("op_start", "term_start", "e_op", "e_dt")}
            w = np.ascontiguousarray(t["w"], dtype=np.float64)
            if int(t["n_terms"]) > 0:
                L.orc_set_s0_ising(self.h, int(t["n_terms"]), arr["op_start"].ctypes.data_as(_ip), arr["term_start"].ctypes.data_as(_ip),
                                   arr["e_op"].ctypes.data_as(_ip), arr["e_dt"].ctypes.data_as(_ip), _d(w), int(t["open_bc"]),
                                   int(bool(getattr(model, "propose_s0", False))))
            elif getattr(model, "propose_s0", False):
                L.orc_set_propose_s0(self.h, 1)

        if getattr(model, "s0_gaussian", False):
            L.orc_set_s0_gaussian(self.h, 1)
        gt = getattr(model, "global_tau", None)
        if gt is not None and (gt["n_global_tau"] > 0 or gt["nt_seq_end"] != model.n_opv):
            L.orc_set_global_tau_sampling(self.h, int(gt["nt_seq_start"]), int(gt["nt_seq_end"]), int(gt["n_global_tau"]))
        gm = getattr(model, "global_move_tau_ising", None)
        if gm is not None:
            a = {k: np.ascontiguousarray(gm[k], dtype=np.int32) for k in ("move_start", "move_fields", "op_start", "term_start", "e_op", "e_dt")}
            w = np.ascontiguousarray(gm["w"], dtype=np.float64)
            L.orc_set_global_move_tau_ising(self.h, int(gm["n_sites"]), a["move_start"].ctypes.data_as(_ip), a["move_fields"].ctypes.data_as(_ip), int(gm["n_terms"]),
                                            a["op_start"].ctypes.data_as(_ip), a["term_start"].ctypes.data_as(_ip), a["e_op"].ctypes.data_as(_ip),
                                            a["e_dt"].ctypes.data_as(_ip), _d(w), int(gm["open_bc"]))

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def global_move_s0(self, site: int, nt: int) -> float:
        """S0_Matter of ham%Global_move_tau for the star move at `site` on slice nt, current configuration (1-based)."""
        return lib().orc_global_move_s0(self.h, int(site), int(nt))

    def get_gm_log(self):
        n = lib().orc_get_gm_log(self.h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint8)
        lib().orc_get_gm_log(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)), n)
        return out[:n]

    def langevin_forces(self):
        """Langevin_HMC_Forces (Prog/Langevin_HMC_mod.F90:107): fermionic forces [nt, n] (complex) of the current configuration."""
        F = np.zeros((self.m.Ltrot, self.m.n_opv), dtype=np.complex128)
        lib().orc_langevin_forces(self.h, _d(F))
        return F

    def langevin_update(self, delta_t, max_force):
        """One update of scheme "Langevin" (Prog/Langevin_HMC_mod.F90:355-392); returns Delta_t_running."""
        return lib().orc_langevin_update(self.h, C.c_double(delta_t), C.c_double(max_force))

    def hmc_update(self, delta_t, leapfrog_steps):
        """One update of scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571); returns (accepted, Weight)."""
        w = C.c_double(0.0)
        acc = lib().orc_hmc_update(self.h, C.c_double(delta_t), int(leapfrog_steps), C.byref(w))
        return bool(acc), w.value

    def compute_fermion_det(self):
        """Compute_Fermion_Det (Prog/Global_mod.F90:792), storage = "Empty": (Phase_det [nf], Det_Vec [nf, ndim])."""
        ph = np.zeros(self.m.N_FL, dtype=np.complex128); dv = np.zeros((self.m.N_FL, self.N))
        lib().orc_compute_fermion_det(self.h, _d(ph), _d(dv))
        return ph, dv

    def s0(self, n: int, nt: int) -> float:
        """ham%S0(n, nt, flipped value) on the current configuration (1-based n, nt)."""
        return lib().orc_s0(self.h, int(n), int(nt))

    # --- RNG / fields
    def ranset(self, seed: int):
        lib().orc_ranset(self.h, int(seed))

    def ranf(self) -> float:
        return lib().orc_ranf(self.h)

    def rng_state(self):
        s = np.zeros(4, dtype=np.uint64)
        lib().orc_get_rng_state(self.h, s.ctypes.data_as(C.c_void_p))
        return s

    def set_rng_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        lib().orc_set_rng_state(self.h, s.ctypes.data_as(C.c_void_p))

    def fields_set(self):
        lib().orc_fields_set(self.h)

    def get_fields(self):
        f = np.zeros((self.m.Ltrot, self.m.n_opv), dtype=np.complex128)
        lib().orc_get_fields(self.h, _d(f))
        return f                      # f[nt-1, n-1]

    def set_fields(self, f):
        f = np.ascontiguousarray(f, dtype=np.complex128)
        assert f.shape == (self.m.Ltrot, self.m.n_opv)
        lib().orc_set_fields(self.h, _d(f))

    # --- driver
    def init(self):
        lib().orc_init(self.h)

    def sweep(self, ltau: int = 0):
        lib().orc_sweep(self.h, int(ltau))

    def n_segments(self, ltau: int = 0) -> int:
        return lib().orc_n_segments(self.h, int(ltau))

    def sweep_segment(self, idx: int, ltau: int = 0):
        """One stabilisation interval of the sweep (sweep() == all segments in order); bench.py's bounded CPU sample."""
        lib().orc_sweep_segment(self.h, int(idx), int(ltau))

    def green(self, nf: int):
        g = np.zeros((self.N, self.N), dtype=np.complex128, order="F")
        lib().orc_get_green(self.h, nf, _d(g))
        return g

    def set_green(self, nf: int, g):
        g = _cplx(g)
        lib().orc_set_green(self.h, nf, _d(g))

    def phase(self) -> complex:
        p = np.zeros(2)
        lib().orc_get_phase(self.h, _d(p))
        return complex(p[0], p[1])

    def nstm(self):
        return lib().orc_nstm(self.h)

    def log(self, on=True):
        lib().orc_log(self.h, int(on))

    def get_log(self):
        n = lib().orc_get_log(self.h, None, None, 0)
        acc = np.zeros(n, dtype=np.uint8)
        w = np.zeros(n, dtype=np.float64)
        lib().orc_get_log(self.h, acc.ctypes.data_as(C.POINTER(C.c_uint8)), _d(w), n)
        return acc, w

    def control(self):
        out = np.zeros(13)
        lib().orc_get_control(self.h, _d(out))
        keys = ["XMEANG", "XMAXG", "NCG", "XMAXP", "XMEAN_tau", "XMAX_tau", "NCG_tau", "NC_up", "ACC_up", "NC_eff_up",
                "ACC_eff_up", "nan", "unstable"]
        return dict(zip(keys, out))

    def taum_capture(self, every=1):
        lib().orc_taum_capture(self.h, int(every))

    def taum_get(self):
        n = lib().orc_taum_get(self.h, None, 0)
        buf = np.zeros(n, dtype=np.complex128)
        lib().orc_taum_get(self.h, _d(buf), n)
        nn = self.N * self.N
        ntau = n // (4 * self.m.N_FL * nn)
        # [tau][which: GT0,G0T,G00,GTT][nf][col-major N*N]
        return buf.reshape(ntau, 4, self.m.N_FL, self.N, self.N).transpose(0, 1, 2, 4, 3)

    def taum_fresh_get(self):
        """[stabilisation][which: GT0,G0T,G00,GTT][nf][N,N]: the matrices right after every CGR2_2 of TAU_M."""
        n = lib().orc_taum_fresh_get(self.h, None, 0)
        buf = np.zeros(n, dtype=np.complex128)
        lib().orc_taum_fresh_get(self.h, _d(buf), n)
        ns = n // (4 * self.m.N_FL * self.N * self.N)
        return buf.reshape(ns, 4, self.m.N_FL, self.N, self.N).transpose(0, 1, 2, 4, 3)

    def obs(self):
        """[N_meas, sum sign, Re/Im sum Part ZP ZS] accumulated where main.F90 calls ham%Obser."""
        out = np.zeros(4)
        lib().orc_get_obs(self.h, _d(out))
        return out

    def set_obs_scal_tables(self, tab):
        """Kin / Pot tables of ham%Obser (see alf_b200_set_obs_scal_tables; `tab` = alf_b200.model.obs_scal_tables(model))."""
        a = {k: np.ascontiguousarray(tab[k], dtype=np.int32) for k in ("kin_i", "kin_j", "kin_nf", "pot_i1", "pot_nf1", "pot_i2", "pot_nf2")}
        kc = np.ascontiguousarray(tab["kin_coef"], dtype=np.complex128); pc = np.ascontiguousarray(tab["pot_coef"], dtype=np.complex128)
        p = lambda x: x.ctypes.data_as(_ip)
        lib().orc_set_obs_scal_tables(self.h, int(kc.size), p(a["kin_i"]), p(a["kin_j"]), p(a["kin_nf"]), _d(kc), int(pc.size), p(a["pot_i1"]), p(a["pot_nf1"]), p(a["pot_i2"]), p(a["pot_nf2"]), _d(pc))

    def obs_full(self):
        """[N_meas, sum sign, Part(re, im), Kin(re, im), Pot(re, im), Ener(re, im)]."""
        out = np.zeros(10)
        lib().orc_get_obs_full(self.h, _d(out))
        return out

    def obs_tau_enable(self):
        n_unit, norb, cell, orb, imj = self.m.lattice_tables()
        imj_f = np.ascontiguousarray(imj.T)
        lib().orc_obs_tau_enable(self.h, int(n_unit), int(norb), cell.ctypes.data_as(_ip), orb.ctypes.data_as(_ip), imj_f.ctypes.data_as(_ip))
        self._obst_dims = (4, lib().orc_obs_tau_ntau(self.h), int(norb), int(norb), int(n_unit))

    def obs_eq_enable(self):
        if not hasattr(self, "_obst_dims"):
            self.obs_tau_enable()
        lib().orc_obs_eq_enable(self.h)

    def obs_eq(self):
        d = self._obst_dims
        acc = np.zeros((4, 1, d[2], d[3], d[4]), dtype=np.complex128); bg = np.zeros((2, 1, d[2]), dtype=np.complex128); cnt = np.zeros(2)
        lib().orc_get_obs_eq(self.h, _d(acc), _d(bg), _d(cnt))
        return acc, bg, cnt[0], cnt[1]

    def obs_tau(self):
        acc = np.zeros(self._obst_dims, dtype=np.complex128); bg = np.zeros((2, self._obst_dims[1], self._obst_dims[2]), dtype=np.complex128); cnt = np.zeros(2)
        lib().orc_get_obs_tau(self.h, _d(acc), _d(bg), _d(cnt))
        return acc, bg, cnt[0], cnt[1]

    def eq_capture(self, on=True):
        lib().orc_eq_capture(self.h, int(on))

    def eq_get(self):
        n = lib().orc_eq_get(self.h, None, 0)
        buf = np.zeros(n, dtype=np.complex128)
        lib().orc_eq_get(self.h, _d(buf), n)
        nn = self.N * self.N
        nv = n // (self.m.N_FL * nn)
        return buf.reshape(nv, self.m.N_FL, self.N, self.N).transpose(0, 1, 3, 2)

    def op_phase_total(self) -> complex:
        """(prod_nf Op_phase(nf))^N_SUN on the current field configuration, starting from Phase = 1 (Prog/Operator_mod.F90:160-181)."""
        p = np.zeros(2); lib().orc_op_phase_total(self.h, _d(p)); return complex(p[0], p[1])

    # --- building blocks
    def hop_apply(self, which: int, nf: int, A):
        A = _cplx(A).copy(order="F")
        lib().orc_hop_apply(self.h, which, nf, _d(A), A.shape[0], A.shape[1])
        return A

    def op_mmultR(self, n, nf, A, field, cop="n"):
        A = _cplx(A).copy(order="F"); field = complex(field)
        lib().orc_op_mmultR(self.h, n, nf, _d(A), A.shape[0], A.shape[1], C.c_double(field.real), C.c_double(field.imag), C.c_char(cop.encode()))
        return A

    def op_mmultL(self, n, nf, A, field, cop="n", sign=1):
        A = _cplx(A).copy(order="F"); field = complex(field)
        lib().orc_op_mmultL(self.h, n, nf, _d(A), A.shape[0], A.shape[1], C.c_double(field.real), C.c_double(field.imag), C.c_char(cop.encode()), int(sign))
        return A

    def op_wrapup(self, n, nf, A, field, ntype):
        A = _cplx(A).copy(order="F"); field = complex(field)
        lib().orc_op_wrapup(self.h, n, nf, _d(A), C.c_double(field.real), C.c_double(field.imag), int(ntype))
        return A

    def op_wrapdo(self, n, nf, A, field, ntype):
        A = _cplx(A).copy(order="F"); field = complex(field)
        lib().orc_op_wrapdo(self.h, n, nf, _d(A), C.c_double(field.real), C.c_double(field.imag), int(ntype))
        return A

    def wrapgr_placegr(self, m, m1, ntau):
        """Wrapgr_PlaceGR (Prog/Wrapgr_mod.F90:247): move GR from operator position m to m1 inside slice ntau."""
        lib().orc_wrapgr_placegr(self.h, int(m), int(m1), int(ntau))

    def wrapgr_random_update(self, m, ntau, t0_ratio, s0_ratio, flip_list, flip_value):
        """One proposal of Wrapgr_Random_update (:317); flip_list 1-based.  Returns (accepted, new m)."""
        fl = np.ascontiguousarray(flip_list, dtype=np.int32); fv = np.ascontiguousarray(flip_value, dtype=np.complex128)
        mm = C.c_int(int(m))
        acc = lib().orc_wrapgr_random_update(self.h, C.byref(mm), int(ntau), C.c_double(t0_ratio), C.c_double(s0_ratio), int(fl.size),
                                             fl.ctypes.data_as(_ip), _d(fv))
        return bool(acc), mm.value

    def wrapgrup(self, ntau):
        lib().orc_wrapgrup(self.h, int(ntau))

    def wrapgrdo(self, ntau):
        lib().orc_wrapgrdo(self.h, int(ntau))

    def propr(self, nf, A, nt):
        A = _cplx(A).copy(order="F")
        lib().orc_propr(self.h, nf, _d(A), int(nt))
        return A

    def get_udv(self, which, nst, nf):
        """which: 0 udvl, 1 udvr, 2 udvst(nst)."""
        U = np.zeros((self.N, self.N), dtype=np.complex128, order="F"); V = U.copy(order="F"); D = np.zeros(self.N, dtype=np.complex128)
        lib().orc_get_udv(self.h, which, nst, nf, _d(U), _d(D), _d(V))
        return U, D, V


def qdrp(A):
    A = _cplx(A).copy(order="F"); n, m = A.shape
    D = np.zeros(m, dtype=np.complex128); ipvt = np.zeros(m, dtype=np.int32); tau = np.zeros(m, dtype=np.complex128)
    lib().orc_qdrp(n, m, _d(A), _d(D), ipvt.ctypes.data_as(_ip), _d(tau))
    return A, D, ipvt, tau


def udv_wrap_pivot(A):
    """UDV_Wrap_Pivot (Prog/UDV_WRAP_mod.F90:125-208, default variant): A (N1 x N2) -> U (N1 x N2), D (N2), V (N2 x N2) with A = U D V."""
    A = _cplx(A).copy(order="F"); n1, n2 = A.shape
    U = np.zeros((n1, n2), dtype=np.complex128, order="F"); V = np.zeros((n2, n2), dtype=np.complex128, order="F"); D = np.zeros(n2, dtype=np.complex128)
    lib().orc_udv_wrap_pivot(n1, n2, _d(A), _d(U), _d(D), _d(V))
    return U, D, V


def udv_decompose(U, D, V, side="r"):
    U = _cplx(U).copy(order="F"); V = _cplx(V).copy(order="F"); D = np.ascontiguousarray(D, dtype=np.complex128).copy()
    lib().orc_udv_decompose(U.shape[0], C.c_char(side.encode()), _d(U), _d(D), _d(V))
    return U, D, V


def cgr(UR, DR, VR, UL, DL, VL, nvar=1, stab3=False):
    n = UR.shape[0]
    G = np.zeros((n, n), dtype=np.complex128, order="F"); ph = np.zeros(2)
    a = [_cplx(x) for x in (UR, VR, UL, VL)]
    dr = np.ascontiguousarray(DR, dtype=np.complex128); dl = np.ascontiguousarray(DL, dtype=np.complex128)
    lib().orc_cgr(n, int(nvar), int(stab3), _d(a[0]), _d(dr), _d(a[1]), _d(a[2]), _d(dl), _d(a[3]), _d(G), _d(ph))
    return G, complex(ph[0], ph[1])


def cgr2_2(U2, D2, V2, U1, D1, V1, stab3=False):
    n = U2.shape[0]
    outs = [np.zeros((n, n), dtype=np.complex128, order="F") for _ in range(4)]
    a = [_cplx(x) for x in (U2, V2, U1, V1)]
    d2 = np.ascontiguousarray(D2, dtype=np.complex128); d1 = np.ascontiguousarray(D1, dtype=np.complex128)
    lib().orc_cgr2_2(n, int(stab3), _d(a[0]), _d(d2), _d(a[1]), _d(a[2]), _d(d1), _d(a[3]), *[_d(o) for o in outs])
    return dict(GRT0=outs[0], GR00=outs[1], GRTT=outs[2], GR0T=outs[3])


def get_blocks(V):
    """get_blocks (Prog/cgr2_2_mod.F90:55-72): (GR00, GR0T, GRT0, GRTT) = the four LQ x LQ blocks of the 2LQ x 2LQ matrix V."""
    V = _cplx(V); lq = V.shape[0] // 2
    out = [np.zeros((lq, lq), dtype=np.complex128, order="F") for _ in range(4)]
    lib().orc_get_blocks(int(lq), _d(V), *[_d(o) for o in out])
    return out


def solve_extended_system(UCT, VINV, inp):
    """QDRP_decompose(inp) followed by solve_extended_System (Prog/cgr2_2_mod.F90:155-191) as in testsuite/Prog.tests/17-solve-extended-system.F90."""
    lq = UCT.shape[0]; inp = _cplx(inp).copy(order="F")
    hlp = np.zeros((2 * lq, 2 * lq), dtype=np.complex128, order="F"); D = np.zeros(2 * lq, dtype=np.complex128)
    lib().orc_solve_extended_system(int(lq), _d(hlp), _d(_cplx(UCT)), _d(_cplx(VINV)), _d(inp), _d(D))
    return hlp


def cgrp(UR, UL):
    """CGRP (Prog/cgr1_mod.F90:464-515) on explicit N x N_part matrices; returns (G, phase)."""
    UR = _cplx(UR); UL = _cplx(UL); n, npart = UR.shape
    G = np.zeros((n, n), dtype=np.complex128, order="F"); ph = np.zeros(2)
    lib().orc_cgrp(int(n), int(npart), _d(UR), _d(UL), _d(G), _d(ph))
    return G, complex(ph[0], ph[1])
