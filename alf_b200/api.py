"""Python host binding of the C-ABI (include/alf_b200.h) via ctypes.

This is the Python twin of the Fortran ISO_C_BINDING shim (alf_b200/fortran/alf_b200_shim.F90): it flattens the
model's Operator tables exactly as the shim does and forwards every call to libalf_b200.so.  The CUDA library is
the only implementation: if it is missing or no sm_100 device is present, construction fails loudly (no fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .model import Model, flatten_ops

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalf_b200.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class AlfError(RuntimeError):
    """Non-zero return code of the C-ABI (the Fortran shim maps these to Terminate_on_error)."""

    def __init__(self, code, msg=""):
        super().__init__(f"alf_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AlfError(100, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.alf_b200_last_error.restype = C.c_char_p
        _lib.alf_b200_last_error.argtypes = [C.c_void_p]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class AlfB200:
    """All chains of one GPU: the batched replacement of the sequential-sweep branch of Prog/main.F90."""

    def __init__(self, model: Model, n_chains: int, nwrap: int = 10, stab: int = 0, device: int = 0):
        L = lib()
        self.m, self.C, self.N, self.nwrap = model, n_chains, model.Ndim, nwrap
        self.h = C.c_void_p()
        rc = L.alf_b200_create(C.byref(self.h), model.Ndim, model.N_FL, model.N_SUN, model.Ltrot, nwrap, model.n_opv, model.n_opt,
                               int(model.Symm), int(stab), int(n_chains), int(device))
        if rc != 0:
            raise AlfError(rc, "alf_b200_create failed (no sm_100 CUDA device?)")
        ov, ot = flatten_ops(model)
        for o in ov:
            self._ck(L.alf_b200_set_op_v(self.h, o["n"], o["nf"], o["N"], o["nnz"], o["diag"], o["type"], o["P"].ctypes.data_as(_ip),
                                         _d(o["U"]), _d(o["E"]), C.c_double(o["g"].real), C.c_double(o["g"].imag),
                                         C.c_double(o["alpha"].real), C.c_double(o["alpha"].imag)))
            if o.get("g_t") is not None:
                if o["g_t"].size != model.Ltrot:
                    raise ValueError("g_t needs Ltrot entries")
                self._ck(L.alf_b200_set_op_v_gt(self.h, o["n"], o["nf"], _d(o["g_t"])))
        for o in ot:
            self._ck(L.alf_b200_set_op_t(self.h, o["nc"], o["nf"], o["N"], o["diag"], o["P"].ctypes.data_as(_ip), _d(o["U"]), _d(o["E"]),
                                         C.c_double(o["g"].real), C.c_double(o["g"].imag)))
        if getattr(model, "Projector", False):
            self._ck(L.alf_b200_set_projector(self.h, int(model.Thtrot), int(model.N_part)))
            for nf in range(model.N_FL):
                pl = np.asfortranarray(model.WF_L[nf], dtype=np.complex128); pr = np.asfortranarray(model.WF_R[nf], dtype=np.complex128)
                self._ck(L.alf_b200_set_trial_wf(self.h, nf + 1, _d(pl), _d(pr)))
        s0 = getattr(model, "s0_ising", None)
        if s0 is not None or getattr(model, "propose_s0", False):
            t = s0 if s0 is not None else dict(n_terms=0, op_start=np.zeros(model.n_opv + 1, np.int32), term_start=np.zeros(1, np.int32),
                                               e_op=np.zeros(1, np.int32), e_dt=np.zeros(1, np.int32), w=np.zeros(2), open_bc=0)
            arr = {k: np.ascontiguousarray(t[k], dtype=np.int32) for k in ("op_start", "term_start", "e_op", "e_dt")}
            w = np.ascontiguousarray(t["w"], dtype=np.float64)
            self._ck(L.alf_b200_set_s0_ising(self.h, int(t["n_terms"]), arr["op_start"].ctypes.data_as(_ip), arr["term_start"].ctypes.data_as(_ip),
                                             arr["e_op"].ctypes.data_as(_ip), arr["e_dt"].ctypes.data_as(_ip), _d(w), int(t["open_bc"]),
                                             int(bool(getattr(model, "propose_s0", False)))))
        if getattr(model, "s0_gaussian", False):
            self._ck(L.alf_b200_set_s0_gaussian(self.h, 1))
        gt = getattr(model, "global_tau", None)
        if gt is not None and (gt["n_global_tau"] > 0 or gt["nt_seq_end"] != model.n_opv):
            self._ck(L.alf_b200_set_global_tau_sampling(self.h, int(gt["nt_seq_start"]), int(gt["nt_seq_end"]), int(gt["n_global_tau"])))
        gm = getattr(model, "global_move_tau_ising", None)
        if gm is not None:
            a = {k: np.ascontiguousarray(gm[k], dtype=np.int32) for k in ("move_start", "move_fields", "op_start", "term_start", "e_op", "e_dt")}
            w = np.ascontiguousarray(gm["w"], dtype=np.float64)
            self._ck(L.alf_b200_set_global_move_tau_ising(self.h, int(gm["n_sites"]), a["move_start"].ctypes.data_as(_ip), a["move_fields"].ctypes.data_as(_ip),
                                                          int(gm["n_terms"]), a["op_start"].ctypes.data_as(_ip), a["term_start"].ctypes.data_as(_ip),
                                                          a["e_op"].ctypes.data_as(_ip), a["e_dt"].ctypes.data_as(_ip), _d(w), int(gm["open_bc"])))
        self._ck(L.alf_b200_finalize_model(self.h))

    def _ck(self, rc):
        if rc != 0:
            msg = lib().alf_b200_last_error(self.h)
            raise AlfError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            lib().alf_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def is_complex(self):
        return bool(lib().alf_b200_is_complex(self.h))

    # ---- RNG / fields
    def set_seeds(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.int32)
        assert s.size == self.C
        self._ck(lib().alf_b200_set_seeds(self.h, s.ctypes.data_as(C.POINTER(C.c_int32))))

    def rng_state(self):
        s = np.zeros((self.C, 4), dtype=np.uint64)
        self._ck(lib().alf_b200_get_rng_state(self.h, s.ctypes.data_as(C.c_void_p)))
        return s

    def set_rng_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        self._ck(lib().alf_b200_set_rng_state(self.h, s.ctypes.data_as(C.c_void_p)))

    def fields_set(self):
        self._ck(lib().alf_b200_fields_set(self.h))

    def get_fields(self):
        f = np.zeros((self.C, self.m.Ltrot, self.m.n_opv), dtype=np.complex128)
        self._ck(lib().alf_b200_get_fields(self.h, _d(f)))
        return f

    def set_fields(self, f):
        f = np.ascontiguousarray(f, dtype=np.complex128)
        assert f.shape == (self.C, self.m.Ltrot, self.m.n_opv)
        self._ck(lib().alf_b200_set_fields(self.h, _d(f)))

    # ---- batched mode
    def init_sweep(self):
        self._ck(lib().alf_b200_init_sweep(self.h))

    def sweep(self, n_sweeps=1, ltau=0):
        self._ck(lib().alf_b200_sweep(self.h, int(n_sweeps), int(ltau)))

    def sweep_host(self, n_sweeps, ltau, fields_in, fields_out, obs_out, control_out):
        self._ck(lib().alf_b200_sweep_host(self.h, int(n_sweeps), int(ltau), _d(fields_in) if fields_in is not None else None,
                                           _d(fields_out) if fields_out is not None else None,
                                           _d(obs_out) if obs_out is not None else None, _d(control_out) if control_out is not None else None))

    # ---- compat mode
    def wrapgrup(self, ntau):
        self._ck(lib().alf_b200_wrapgrup(self.h, int(ntau)))

    def wrapgrdo(self, ntau):
        self._ck(lib().alf_b200_wrapgrdo(self.h, int(ntau)))

    def wrapur(self, ntau, ntau1):
        self._ck(lib().alf_b200_wrapur(self.h, int(ntau), int(ntau1)))

    def wrapul(self, ntau1, ntau):
        self._ck(lib().alf_b200_wrapul(self.h, int(ntau1), int(ntau)))

    def udv_reset(self, which, side):
        self._ck(lib().alf_b200_udv_reset(self.h, int(which), C.c_char(side.encode())))

    def cgr(self, nvar=1):
        self._ck(lib().alf_b200_cgr(self.h, int(nvar)))

    def tau_m(self):
        self._ck(lib().alf_b200_tau_m(self.h))

    def tau_p(self, nst_in):
        self._ck(lib().alf_b200_tau_p(self.h, int(nst_in)))

    # ---- device-side time-displaced lattice observables (ObserT of the shipped Hamiltonians)
    def obs_tau_enable(self, on=True):
        n_unit, norb, cell, orb, imj = self.m.lattice_tables()
        imj_f = np.ascontiguousarray(imj.T)                       # Fortran order: imj(I, J) at I-1 + (J-1) n_unit
        self._ck(lib().alf_b200_set_lattice(self.h, int(n_unit), int(norb), cell.ctypes.data_as(_ip), orb.ctypes.data_as(_ip), imj_f.ctypes.data_as(_ip)))
        self._ck(lib().alf_b200_obs_tau_enable(self.h, int(on)))
        self._obs_tau_on = bool(on) or getattr(self, "_obs_tau_on", False)

    def obs_eq_enable(self, on=True):
        n_unit, norb, cell, orb, imj = self.m.lattice_tables()
        imj_f = np.ascontiguousarray(imj.T)
        self._ck(lib().alf_b200_set_lattice(self.h, int(n_unit), int(norb), cell.ctypes.data_as(_ip), orb.ctypes.data_as(_ip), imj_f.ctypes.data_as(_ip)))
        self._ck(lib().alf_b200_obs_eq_enable(self.h, int(on)))
        self._obs_eq_on = bool(on) or getattr(self, "_obs_eq_on", False)

    def obs_eq(self):
        n_unit, norb = self.m.latt.N, self.m.n_orb
        acc = np.zeros((4, 1, norb, norb, n_unit), dtype=np.complex128); bg = np.zeros((2, 1, norb), dtype=np.complex128); cnt = np.zeros(2)
        self._ck(lib().alf_b200_get_obs_eq(self.h, _d(acc), _d(bg), _d(cnt)))
        return acc, bg, cnt[0], cnt[1]

    def obs_tau_reset(self):
        self._ck(lib().alf_b200_obs_tau_reset(self.h))

    def obs_tau(self):
        """(acc[ch, nt, no_J, no_I, imj] complex, bg[which, nt, no] complex, N, sum of signs), summed over the chains of the handle."""
        nch, ntau, norb, nu = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(lib().alf_b200_obs_tau_dims(self.h, C.byref(nch), C.byref(ntau), C.byref(norb), C.byref(nu)))
        acc = np.zeros((nch.value, ntau.value, norb.value, norb.value, nu.value), dtype=np.complex128)
        bg = np.zeros((2, ntau.value, norb.value), dtype=np.complex128); cnt = np.zeros(2)
        self._ck(lib().alf_b200_get_obs_tau(self.h, _d(acc), _d(bg), _d(cnt)))
        return acc, bg, cnt[0], cnt[1]

    # ---- global-in-slice moves (Wrapgr_PlaceGR / Wrapgr_Random_update)
    def wrapgr_set_position(self, m):
        self._ck(lib().alf_b200_wrapgr_set_position(self.h, int(m)))

    def wrapgr_get_position(self):
        m = np.zeros(self.C, dtype=np.int32)
        self._ck(lib().alf_b200_wrapgr_get_position(self.h, m.ctypes.data_as(_ip)))
        return m

    def langevin_forces(self):
        """Langevin_HMC_Forces (Prog/Langevin_HMC_mod.F90:107): fermionic forces [chain, nt, n] (complex)."""
        F = np.zeros((self.C, self.m.Ltrot, self.m.n_opv), dtype=np.complex128)
        self._ck(lib().alf_b200_langevin_forces(self.h, _d(F)))
        return F

    def langevin_update(self, delta_t, max_force):
        """One update of scheme "Langevin" (Prog/Langevin_HMC_mod.F90:355-392) for every chain; returns Delta_t_running [chain]."""
        dt = np.zeros(self.C)
        self._ck(lib().alf_b200_langevin_update(self.h, C.c_double(delta_t), C.c_double(max_force), _d(dt)))
        return dt

    def hmc_update(self, delta_t, leapfrog_steps):
        """One update of scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571) for every chain; returns (accepted [chain] bool, Weight [chain])."""
        w = np.zeros(self.C); acc = np.zeros(self.C, dtype=np.uint8)
        self._ck(lib().alf_b200_hmc_update(self.h, C.c_double(delta_t), int(leapfrog_steps), _d(w), acc.ctypes.data_as(C.POINTER(C.c_uint8))))
        return acc.astype(bool), w

    def compute_fermion_det(self):
        """Compute_Fermion_Det (Prog/Global_mod.F90:792) with storage = "Empty": (log|det| [chain, nf], phase [chain, nf]) of the current fields."""
        ld = np.zeros((self.C, self.m.N_FL)); ph = np.zeros((self.C, self.m.N_FL), dtype=np.complex128)
        self._ck(lib().alf_b200_compute_fermion_det(self.h, _d(ld), _d(ph)))
        return ld, ph

    def wrapgr_placegr(self, m1, ntau):
        self._ck(lib().alf_b200_wrapgr_placegr(self.h, int(m1), int(ntau)))

    def wrapgr_random_update(self, ntau, flip_length, flip_list, flip_value, t0_ratio, s0_ratio, place_to=-1):
        """flip_length, t0_ratio, s0_ratio: [chain, move]; flip_list (1-based), flip_value: [chain, move, maxlen]."""
        fl = np.ascontiguousarray(flip_length, dtype=np.int32); C_, nm = fl.shape
        li = np.ascontiguousarray(flip_list, dtype=np.int32); fv = np.ascontiguousarray(flip_value, dtype=np.complex128)
        t0 = np.ascontiguousarray(t0_ratio, dtype=np.float64); s0 = np.ascontiguousarray(s0_ratio, dtype=np.float64)
        assert C_ == self.C and li.shape == fv.shape == (C_, nm, li.shape[2])
        acc = np.zeros((C_, nm), dtype=np.uint8)
        self._ck(lib().alf_b200_wrapgr_random_update(self.h, int(ntau), int(nm), int(li.shape[2]), fl.ctypes.data_as(_ip), li.ctypes.data_as(_ip), _d(fv),
                                                     _d(t0), _d(s0), acc.ctypes.data_as(C.POINTER(C.c_uint8)), int(place_to)))
        return acc

    # ---- results
    def green(self, chain, nf, symmetrize=False):
        g = np.zeros((self.N, self.N), dtype=np.complex128, order="F")
        self._ck(lib().alf_b200_get_green(self.h, int(chain), int(nf), int(symmetrize), _d(g)))
        return g

    def set_green(self, chain, nf, g):
        g = np.asfortranarray(g, dtype=np.complex128)
        self._ck(lib().alf_b200_set_green(self.h, int(chain), int(nf), _d(g)))

    def phase(self):
        p = np.zeros(self.C, dtype=np.complex128)
        self._ck(lib().alf_b200_get_phase(self.h, _d(p)))
        return p

    def get_udv(self, which, nst, chain, nf):
        U = np.zeros((self.N, self.N), dtype=np.complex128, order="F"); V = U.copy(order="F"); D = np.zeros(self.N, dtype=np.complex128)
        self._ck(lib().alf_b200_get_udv(self.h, int(which), int(nst), int(chain), int(nf), _d(U), _d(D), _d(V)))
        return U, D, V

    def set_udv(self, which, nst, chain, nf, U, D, V=None):
        """Host UDV_State -> the handle's udvl (0) / udvr (1) / udvst(nst) (2) of one chain and flavor (compat mode)."""
        U = np.asfortranarray(U, dtype=np.complex128); D = np.ascontiguousarray(D, dtype=np.complex128)
        Vp = _d(np.asfortranarray(V, dtype=np.complex128)) if V is not None else None
        self._ck(lib().alf_b200_set_udv(self.h, int(which), int(nst), int(chain), int(nf), _d(U), _d(D), Vp))

    # ---- multi-GPU bin reduction behind the C-ABI (NCCL)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().alf_b200_comm_unique_id(buf)
        if rc != 0:
            raise AlfError(rc, "alf_b200_comm_unique_id (is libnccl.so.2 available?)")
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        self._ck(lib().alf_b200_comm_init(self.h, int(nranks), int(rank), C.create_string_buffer(uid, 128)))

    def reduce_bins(self, root=0):
        self._ck(lib().alf_b200_reduce_bins(self.h, int(root)))

    def reduce_control(self, root=0):
        out = np.zeros(16)
        self._ck(lib().alf_b200_reduce_control(self.h, int(root), _d(out)))
        keys = ["XMEANG", "XMAXG", "NCG", "XMAXP", "XMEAN_tau", "XMAX_tau", "NCG_tau", "NC_up", "ACC_up", "NC_eff_up", "ACC_eff_up", "nan", "unstable", "flushes"]
        return dict(zip(keys, out))

    def control(self):
        out = np.zeros(16)
        self._ck(lib().alf_b200_get_control(self.h, _d(out)))
        keys = ["XMEANG", "XMAXG", "NCG", "XMAXP", "XMEAN_tau", "XMAX_tau", "NCG_tau", "NC_up", "ACC_up", "NC_eff_up", "ACC_eff_up", "nan", "unstable", "flushes"]
        return dict(zip(keys, out))

    def accept_log(self, n_sweeps=1):
        self._log_sweeps = int(n_sweeps)
        self._ck(lib().alf_b200_accept_log(self.h, int(n_sweeps)))

    def get_accept_log(self):
        n = C.c_long(0)
        cap = 2 * self.m.Ltrot * self.m.n_opv * self.C * max(1, getattr(self, '_log_sweeps', 1))
        out = np.full(cap, 255, dtype=np.uint8)
        self._ck(lib().alf_b200_get_accept_log(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_long(cap), C.byref(n)))
        return out[: n.value * self.C].reshape(self.C, n.value)

    def taum_capture(self, every=1):
        self._ck(lib().alf_b200_taum_capture(self.h, int(every)))

    def get_taum(self, chain):
        n = C.c_long(0)
        self._ck(lib().alf_b200_get_taum(self.h, int(chain), None, C.c_long(0), C.byref(n)))
        buf = np.zeros(n.value, dtype=np.complex128)
        self._ck(lib().alf_b200_get_taum(self.h, int(chain), _d(buf), C.c_long(n.value), C.byref(n)))
        nn = self.N * self.N
        ntau = n.value // (4 * self.m.N_FL * nn)
        return buf.reshape(ntau, 4, self.m.N_FL, self.N, self.N).transpose(0, 1, 2, 4, 3)

    def get_taum_fresh(self, chain):
        n = C.c_long(0)
        self._ck(lib().alf_b200_get_taum_fresh(self.h, int(chain), None, C.c_long(0), C.byref(n)))
        buf = np.zeros(n.value, dtype=np.complex128)
        self._ck(lib().alf_b200_get_taum_fresh(self.h, int(chain), _d(buf), C.c_long(n.value), C.byref(n)))
        ns = n.value // (4 * self.m.N_FL * self.N * self.N)
        return buf.reshape(ns, 4, self.m.N_FL, self.N, self.N).transpose(0, 1, 2, 4, 3)

    def obs_size(self):
        return int(lib().alf_b200_obs_size(self.h))

    def obs(self):
        n = lib().alf_b200_obs_size(self.h)
        out = np.zeros(n)
        self._ck(lib().alf_b200_get_obs(self.h, _d(out)))
        return out

    def obs_reset(self):
        self._ck(lib().alf_b200_obs_reset(self.h))

    def set_measure_interval(self, lobs_st=0, lobs_en=0):
        """LOBS_ST / LOBS_EN of VAR_QMC (alf_b200_set_measure_interval): equal-time measurements on the slices lobs_st <= NTAU1 <= lobs_en; 0 = default."""
        self._ck(lib().alf_b200_set_measure_interval(self.h, int(lobs_st), int(lobs_en)))

    def set_obs_scal_tables(self, tab):
        """Kin / Pot / Ener of ham%Obser on the device from tables (alf_b200_set_obs_scal_tables); `tab` = alf_b200.model.obs_scal_tables(model)."""
        a = {k: np.ascontiguousarray(tab[k], dtype=np.int32) for k in ("kin_i", "kin_j", "kin_nf", "pot_i1", "pot_nf1", "pot_i2", "pot_nf2")}
        kc = np.ascontiguousarray(tab["kin_coef"], dtype=np.complex128); pc = np.ascontiguousarray(tab["pot_coef"], dtype=np.complex128)
        p = lambda x: x.ctypes.data_as(_ip)
        self._ck(lib().alf_b200_set_obs_scal_tables(self.h, int(kc.size), p(a["kin_i"]), p(a["kin_j"]), p(a["kin_nf"]), _d(kc),
                                                    int(pc.size), p(a["pot_i1"]), p(a["pot_nf1"]), p(a["pot_i2"]), p(a["pot_nf2"]), _d(pc)))

    def obs_device_ptr(self):
        p = _dp(); n = C.c_long(0)
        self._ck(lib().alf_b200_obs_device_ptr(self.h, C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, n.value

    # ---- measurement support
    KCATS = ["update", "ops_wrap", "qrp", "formq", "gemm", "trsm", "elementwise", "obs"]

    def stream_ptr(self):
        p = C.c_void_p()
        self._ck(lib().alf_b200_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def kernel_timing(self, mask=0):
        self._ck(lib().alf_b200_kernel_timing(self.h, C.c_uint(mask)))

    def kernel_stats(self):
        ms = np.zeros(8); n = np.zeros(8, dtype=np.int64)
        self._ck(lib().alf_b200_get_kernel_stats(self.h, _d(ms), n.ctypes.data_as(C.POINTER(C.c_long))))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KCATS)}

    def kernel_flops(self):
        fl = np.zeros(8)
        self._ck(lib().alf_b200_get_kernel_flops(self.h, _d(fl)))
        return {k: float(fl[i]) for i, k in enumerate(self.KCATS)}

    def hop_apply(self, which, nf, A):
        A = np.asfortranarray(A, dtype=np.complex128).copy(order="F")
        self._ck(lib().alf_b200_hop_apply(self.h, int(which), int(nf), _d(A)))
        return A


# ---- kernel-level test entry points -------------------------------------------------------------------------------
def _chk(rc, what):
    if rc != 0:
        raise AlfError(rc, what)


def test_gemm(A, B, ta=False, tb=False, is_complex=True, device=0):
    A = np.ascontiguousarray(A, dtype=np.complex128); B = np.ascontiguousarray(B, dtype=np.complex128)
    batch = A.shape[0]
    # inputs are [batch, rows, cols] in numpy order -> convert each to column-major storage
    Af = np.ascontiguousarray(A.transpose(0, 2, 1)); Bf = np.ascontiguousarray(B.transpose(0, 2, 1))
    m = A.shape[2] if ta else A.shape[1]; k = A.shape[1] if ta else A.shape[2]; n = B.shape[1] if tb else B.shape[2]
    Cc = np.zeros((batch, n, m), dtype=np.complex128)
    _chk(lib().alf_b200_test_gemm(device, int(is_complex), int(ta), int(tb), m, n, k, batch, _d(Af), _d(Bf), _d(Cc)), "test_gemm")
    return Cc.transpose(0, 2, 1)


def test_qdrp(A, is_complex=True, device=0):
    A = np.ascontiguousarray(A, dtype=np.complex128); batch, m, n = A.shape
    Af = np.ascontiguousarray(A.transpose(0, 2, 1)).copy()
    D = np.zeros((batch, n)); jp = np.zeros((batch, n), dtype=np.int32); tau = np.zeros((batch, n), dtype=np.complex128); ph = np.zeros((batch, 5))
    _chk(lib().alf_b200_test_qdrp(device, int(is_complex), m, n, batch, _d(Af), _d(D), jp.ctypes.data_as(_ip), _d(tau), _d(ph)), "test_qdrp")
    return Af.transpose(0, 2, 1), D, jp, tau, ph


def test_qdrp_blocked(A, is_complex=True, device=0):
    A = np.ascontiguousarray(A, dtype=np.complex128); batch, m, n = A.shape
    Af = np.ascontiguousarray(A.transpose(0, 2, 1)).copy(); Q = np.zeros((batch, m, m), dtype=np.complex128)
    D = np.zeros((batch, n)); jp = np.zeros((batch, n), dtype=np.int32); tau = np.zeros((batch, n), dtype=np.complex128); ph = np.zeros((batch, 5))
    _chk(lib().alf_b200_test_qdrp_blocked(device, int(is_complex), m, n, batch, _d(Af), _d(D), jp.ctypes.data_as(_ip), _d(tau), _d(ph), _d(Q)), "test_qdrp_blocked")
    return Af.transpose(0, 2, 1), D, jp, tau, ph, Q.transpose(0, 2, 1)


def udv_wrap_pivot(A, is_complex=True, device=0):
    """UDV_Wrap_Pivot(A, U, D, V, NCON, N1, N2) (Prog/UDV_WRAP_mod.F90:125-208) on a batch A[b] (N1 x N2): returns U, D, V with A = U D V."""
    A = np.ascontiguousarray(A, dtype=np.complex128); batch, n1, n2 = A.shape
    Af = np.ascontiguousarray(A.transpose(0, 2, 1)).copy()
    Uf = np.zeros((batch, n2, n1), dtype=np.complex128); Vf = np.zeros((batch, n2, n2), dtype=np.complex128); D = np.zeros((batch, n2), dtype=np.complex128)
    _chk(lib().alf_b200_udv_wrap_pivot(device, int(is_complex), n1, n2, batch, _d(Af), _d(Uf), _d(D), _d(Vf)), "udv_wrap_pivot")
    return Uf.transpose(0, 2, 1), D, Vf.transpose(0, 2, 1)


def test_udv_decompose(U, D, V, side="r", is_complex=True, device=0):
    U = np.ascontiguousarray(U, dtype=np.complex128); V = np.ascontiguousarray(V, dtype=np.complex128); batch, n, _ = U.shape
    Uf = np.ascontiguousarray(U.transpose(0, 2, 1)).copy(); Vf = np.ascontiguousarray(V.transpose(0, 2, 1)).copy()
    Dc = np.ascontiguousarray(D, dtype=np.complex128).copy()
    _chk(lib().alf_b200_test_udv_decompose(device, int(is_complex), n, batch, C.c_char(side.encode()), _d(Uf), _d(Dc), _d(Vf)), "test_udv")
    return Uf.transpose(0, 2, 1), Dc, Vf.transpose(0, 2, 1)


def test_cgr(UR, DR, VR, UL, DL, VL, detUR, detUL, nvar=1, stab=0, is_complex=True, device=0):
    batch, n, _ = UR.shape
    cm = lambda X: np.ascontiguousarray(np.asarray(X, dtype=np.complex128).transpose(0, 2, 1))
    a = [cm(x) for x in (UR, VR, UL, VL)]
    dr = np.ascontiguousarray(DR, dtype=np.complex128); dl = np.ascontiguousarray(DL, dtype=np.complex128)
    have_det = detUR is not None and detUL is not None      # None: det U_R, det U_L are computed on the device (the reference routine's own argument list)
    d1 = np.ascontiguousarray(detUR, dtype=np.complex128) if have_det else None; d2 = np.ascontiguousarray(detUL, dtype=np.complex128) if have_det else None
    G = np.zeros((batch, n, n), dtype=np.complex128); ph = np.zeros(batch, dtype=np.complex128)
    _chk(lib().alf_b200_test_cgr(device, int(is_complex), n, batch, int(nvar), int(stab), _d(a[0]), _d(dr), _d(a[1]), _d(a[2]), _d(dl), _d(a[3]),
                                 _d(d1) if have_det else None, _d(d2) if have_det else None, _d(G), _d(ph)), "test_cgr")
    return G.transpose(0, 2, 1), ph


def test_cgrp(UR, UL, is_complex=True, device=0):
    """UR, UL: [batch, n, n_part]; returns G [batch, n, n] and the phases."""
    UR = np.asarray(UR, dtype=np.complex128); UL = np.asarray(UL, dtype=np.complex128); batch, n, npart = UR.shape
    a = np.ascontiguousarray(UR.transpose(0, 2, 1)); b = np.ascontiguousarray(UL.transpose(0, 2, 1))
    G = np.zeros((batch, n, n), dtype=np.complex128); ph = np.zeros(batch, dtype=np.complex128)
    _chk(lib().alf_b200_test_cgrp(device, int(is_complex), n, npart, batch, _d(a), _d(b), _d(G), _d(ph)), "test_cgrp")
    return G.transpose(0, 2, 1), ph


def test_cgr2_2(U2, D2, V2, U1, D1, V1, stab=0, is_complex=True, device=0):
    batch, n, _ = U2.shape
    cm = lambda X: np.ascontiguousarray(np.asarray(X, dtype=np.complex128).transpose(0, 2, 1))
    a = [cm(x) for x in (U2, V2, U1, V1)]
    d2 = np.ascontiguousarray(D2, dtype=np.complex128); d1 = np.ascontiguousarray(D1, dtype=np.complex128)
    out = np.zeros((4, batch, n, n), dtype=np.complex128)
    _chk(lib().alf_b200_test_cgr2_2(device, int(is_complex), n, batch, int(stab), _d(a[0]), _d(d2), _d(a[1]), _d(a[2]), _d(d1), _d(a[3]), _d(out)), "test_cgr2_2")
    o = out.transpose(0, 1, 3, 2)
    return dict(GRT0=o[0], GR00=o[1], GRTT=o[2], GR0T=o[3])


def fp64_peak(device=0):
    a = C.c_double(0); b = C.c_double(0)
    _chk(lib().alf_b200_fp64_peak(device, C.byref(a), C.byref(b)), "fp64_peak")
    return a.value, b.value
