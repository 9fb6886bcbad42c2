"""Parallel tempering: the configuration exchange of Global_mod's Exchange_Step (Prog/Global_mod.F90:108-420) for handles that run the same lattice with
different parameters ("Temp_0", "Temp_1", ... in ALF: one MPI group per parameter set).

ALF pairs rank r of group g with rank r of a neighbouring group; here chain c of handle g is paired with chain c of the neighbouring handle, all chains at
once.  Per exchange step (Tempering_calc_det = .true.):
  1. Compute_Fermion_Det of every chain under its own Hamiltonian (:792, on the device: alf_b200_compute_fermion_det);
  2. a direction is drawn, the groups are paired on a ring (npbc_tempering), partners swap their configurations (MPI_Sendrecv, :232-237);
  3. Compute_Fermion_Det of the received configuration, Compute_Ratio_Global (:651-760) = weight of the new over the old configuration under the OWN
     Hamiltonian: Ratio(1) = prod gamma(new)/gamma(old) * prod_nf [Phase_Det_new/Phase_Det_old * exp(sum (phi_new - phi_old) g alpha)]^N_SUN,
     Ratio(2) = N_SUN sum_nf (log|det|_new - log|det|_old) + Delta_S0 (0 for the discrete Hubbard-Stratonovich actions handled here);
  4. the master of a pair accepts with Weight = |Ratio(1) Ratio_p(1) exp(Ratio(2) + Ratio_p(2))| > ranf (:277-281); rejected pairs take their old configurations back.
After the last step ALF moves GR, Phase and the UDV storage to wherever a configuration ended up (:380-412); here every handle rebuilds them from its fields
(alf_b200_init_sweep), which is the same state up to round-off.

Scope: finite-temperature or projective runs with discrete fields (types 1, 2) and S0 = 1.  The random numbers of the exchange come from the caller's
generator (ALF: the master rank's stream), which is the one documented difference."""
from __future__ import annotations

import numpy as np

_S6 = np.sqrt(6.0)
# Phi_st(s, type), Gama_st(s, type) for s = -2 .. 2 (index s + 2), Prog/Fields_mod.F90:270-285
PHI_ST = {1: np.array([-2.0, -1.0, 0.0, 1.0, 2.0]),
          2: np.array([-np.sqrt(2.0 * (3.0 + _S6)), -np.sqrt(2.0 * (3.0 - _S6)), 0.0, np.sqrt(2.0 * (3.0 - _S6)), np.sqrt(2.0 * (3.0 + _S6))])}
GAMA_ST = {1: np.ones(5), 2: np.array([1.0 - _S6 / 3.0, 1.0 + _S6 / 3.0, 1.0, 1.0 + _S6 / 3.0, 1.0 - _S6 / 3.0])}


def _phi_gama(model, fields):
    """phi, gama of a field array [chain, Ltrot, n_opv] (complex, as nsigma%f) for the model's field types."""
    s = np.rint(fields.real).astype(np.int64) + 2
    phi = np.zeros(s.shape); gam = np.ones(s.shape)
    for n, row in enumerate(model.Op_V):
        t = row[0].type
        if t not in (1, 2):
            raise ValueError("tempering exchange: discrete fields (types 1, 2) only")
        phi[..., n] = PHI_ST[t][s[..., n]]; gam[..., n] = GAMA_ST[t][s[..., n]]
    return phi, gam


def compute_ratio_global(model, ld_old, ph_old, ld_new, ph_new, f_old, f_new):
    """Compute_Ratio_Global (Prog/Global_mod.F90:651-760) per chain, with log_T0_Proposal_ratio = 0 and Delta_S0 = 0: returns (Ratio1 [chain] complex,
    Ratio2 [chain] real); ld_*, ph_*: [chain, N_FL] from compute_fermion_det; f_*: [chain, Ltrot, n_opv]."""
    n_sun = model.N_SUN
    r2 = n_sun * (ld_new - ld_old).sum(axis=1)
    phi_o, gam_o = _phi_gama(model, f_old); phi_n, gam_n = _phi_gama(model, f_new)
    r1 = np.prod(gam_n / gam_o, axis=(1, 2)).astype(np.complex128)
    dphi = phi_n - phi_o                                              # [chain, nt, n]
    for nf in range(model.N_FL):
        ga = np.array([row[nf].g * row[nf].alpha for row in model.Op_V])                     # g alpha per vertex
        gt = [row[nf].g_t for row in model.Op_V]
        if any(x is not None for x in gt):                                                   # g_t(nt) alpha (:723)
            gat = np.stack([(np.asarray(row[nf].g_t) if row[nf].g_t is not None else np.full(model.Ltrot, row[nf].g)) * row[nf].alpha for row in model.Op_V], axis=1)
            e = np.exp(n_sun * (dphi * gat[None, :, :]).sum(axis=(1, 2)))
        else:
            e = np.exp(n_sun * (dphi * ga[None, None, :]).sum(axis=(1, 2)))
        r1 = r1 * (ph_new[:, nf] / ph_old[:, nf]) ** n_sun * e
    return r1, r2


def exchange_step(handles, rng, n_exchange_steps=1, rebuild=True):
    """Exchange_Step for a ring of handles (one per parameter set, same lattice, same number of chains).  rng: numpy Generator.
    Returns (accepted [step, pair, chain] bool, weights [step, pair, chain], pairs [step] list of (master, partner)).  The handles' fields are updated in
    place; with rebuild=True every handle recomputes its storage, G and phase (alf_b200_init_sweep) at the end."""
    ng = len(handles)
    if ng < 2 or ng % 2:
        raise ValueError("tempering needs an even number (>= 2) of parameter sets")      # List_masters: every second group is a master
    C = handles[0].C
    acc_all, w_all, pairs_all = [], [], []
    fields = [h.get_fields() for h in handles]
    for _ in range(n_exchange_steps):
        old = [h.compute_fermion_det() for h in handles]
        n_step = -1 if rng.random() > 0.5 else 1                                          # :204-205
        partner = [0] * ng
        for m in range(0, ng, 2):
            p = (m + n_step) % ng
            partner[m] = p; partner[p] = m
        f_old = [f.copy() for f in fields]
        f_new = [f_old[partner[g]] for g in range(ng)]
        for g, h in enumerate(handles):
            h.set_fields(f_new[g])
        new = [h.compute_fermion_det() for h in handles]
        ratio = [compute_ratio_global(handles[g].m, old[g][0], old[g][1], new[g][0], new[g][1], f_old[g], f_new[g]) for g in range(ng)]
        acc_step, w_step, pairs = [], [], []
        for m in range(0, ng, 2):
            p = partner[m]
            weight = np.abs(ratio[m][0] * ratio[p][0] * np.exp(ratio[m][1] + ratio[p][1]))      # :277-279
            toggle = weight > rng.random(C)
            for g in (m, p):
                fields[g] = np.where(toggle[:, None, None], f_new[g], f_old[g])
            acc_step.append(toggle); w_step.append(weight); pairs.append((m, p))
        for g, h in enumerate(handles):
            h.set_fields(fields[g])
        acc_all.append(np.array(acc_step)); w_all.append(np.array(w_step)); pairs_all.append(pairs)
    if rebuild:
        for h in handles:
            h.init_sweep()
    return np.array(acc_all), np.array(w_all), pairs_all


def global_updates(handle, propose, rng, n_global=1, rebuild=True):
    """Global_Updates (Prog/Global_mod.F90:450-639) for all chains of a handle: n_global whole-configuration proposals per chain.
    propose(fields_old [chain, Ltrot, n_opv] complex, rng) -> (fields_new, log_T0_proposal_ratio [chain]) plays ham%Global_move_log_T0 (-inf = no proposal,
    LOG_T0_REJECTED).  Per proposal: Compute_Fermion_Det of the new configuration on the device, Ratiotot = Compute_Ratio_Global, Weight =
    |Re(Phase_old Ratiotot) / Re(Phase_old)| (:569), accepted if Weight > ranf; rejected chains take their fields back.  Returns (accepted [n_global, chain],
    weight [n_global, chain]); the phase a chain would carry is Phase_old Ratiotot / |Ratiotot| (what Control_PrecisionP_Glob compares, :571-572).
    With rebuild=True the handle recomputes storage, G and phase from its fields at the end (the reference rebuilds them in :601-637)."""
    m = handle.m; C = handle.C
    f_old = handle.get_fields(); phase_old = np.asarray(handle.phase(), dtype=np.complex128).copy()
    ld_old, ph_old = handle.compute_fermion_det()
    acc_all, w_all = [], []
    for _ in range(n_global):
        f_new, log_t0 = propose(f_old.copy(), rng)
        log_t0 = np.asarray(log_t0, dtype=np.float64); live = np.isfinite(log_t0)
        f_try = np.where(live[:, None, None], f_new, f_old)
        handle.set_fields(f_try)
        ld_new, ph_new = handle.compute_fermion_det()
        r1, r2 = compute_ratio_global(m, ld_old, ph_old, ld_new, ph_new, f_old, f_try)
        ratiotot = r1 * np.exp(r2 + np.where(live, log_t0, 0.0))
        weight = np.abs((phase_old * ratiotot).real / phase_old.real)
        toggle = live & (weight > rng.random(C))
        f_old = np.where(toggle[:, None, None], f_try, f_old)
        ld_old = np.where(toggle[:, None], ld_new, ld_old); ph_old = np.where(toggle[:, None], ph_new, ph_old)
        phase_old = np.where(toggle, phase_old * ratiotot / np.abs(ratiotot), phase_old)
        acc_all.append(toggle); w_all.append(np.where(live, weight, 0.0))
    handle.set_fields(f_old)
    if rebuild:
        handle.init_sweep()
    return np.array(acc_all), np.array(w_all)
