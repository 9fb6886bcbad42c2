/* alf_b200.h -- C-ABI of the B200-native auxiliary-field QMC sweep (drop-in for ALF's L2/L3 hot path).
 *
 * ALF has no FFI around this path: the seam is a set of Fortran module procedures that share module-global
 * state.  A Fortran shim (alf_b200/fortran/alf_b200_shim.F90, shown in INTEGRATION.md) keeps those
 * procedure names and forwards their bodies to the entry points below.  Plain pointers and sizes only; every
 * function returns 0 on success and a non-zero ALF error code otherwise (mapped by the shim to
 * Terminate_on_error, Libraries/Modules/runtime_error_mod.F90:56-68,81-130).  No global state: everything hangs
 * off the opaque handle; one CUDA stream per handle; not re-entrant per handle (as the reference: one thread/rank).
 *
 * Conventions at the boundary (all as on the Fortran side):
 *   - complex numbers are interleaved (re, im) doubles = complex(kind(0.d0));
 *   - matrices are column-major with leading dimension = number of rows;
 *   - index arrays (P, time slices nt, operator index n, flavor nf) are 1-BASED;
 *   - "chain" = one Markov chain = one MPI rank of the reference; chains are 0-based (they have no Fortran analogue).
 */
#ifndef ALF_B200_H
#define ALF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct alf_b200_handle alf_b200_handle;

/* error codes: Libraries/Modules/runtime_error_mod.F90:56-68 */
enum {
  ALF_OK = 0,
  ALF_ERROR_GENERIC = 1,
  ALF_ERROR_HAMILTONIAN = 2,
  ALF_ERROR_FIELDS = 3,
  ALF_ERROR_UNSTABLE_MATRIX = 4,   /* Control_PrecisionG threshold exceeded / NaN (control_mod.F90:219-283) */
  ALF_ERROR_CUDA = 100,            /* CUDA runtime failure or no sm_100 device: there is NO CPU fallback */
  ALF_ERROR_UNSUPPORTED = 101
};

/* ---- life cycle.  Replaces the allocations of Prog/main.F90:589-605 and Wrapgr_alloc (Wrapgr_mod.F90:70-78).
 * ndim, n_fl, n_sun, ltrot, symm: public state of Hamiltonian_main (Prog/Hamiltonian_main_mod.F90:181-197);
 * nwrap: VAR_QMC (Prog/QMC_runtime_var_mod.F90:73-79); n_opv = size(Op_V,1); n_opt = size(Op_T,1);
 * n_chains: number of independent Markov chains advanced by every call; device: CUDA ordinal.
 * stab: 0 = default CGR branch, 3 = scale-separated branch (-DSTAB3, configure.sh:332-357). */
int alf_b200_create(alf_b200_handle** h, int ndim, int n_fl, int n_sun, int ltrot, int nwrap, int n_opv, int n_opt,
                    int symm, int stab, int n_chains, int device);
int alf_b200_destroy(alf_b200_handle* h);
const char* alf_b200_last_error(const alf_b200_handle* h);

/* ---- model tables.  One call per Op_V(n,nf) / Op_T(nc,nf) with the PUBLIC fields of type Operator
 * (Prog/Operator_mod.F90:57-69).  M_exp/E_exp and Hop_mod's ExpOpT_vec are private in ALF, so finalize_model
 * recomputes them exactly as Op_set/Op_exp do (Operator_mod.F90:400-529, OpTTypes_mod.F90:92-133,203-234),
 * i.e. it replaces Hop_mod_init (Prog/Hop_mod.F90:118). */
int alf_b200_set_op_v(alf_b200_handle* h, int n, int nf, int N, int n_non_zero, int diag, int type, const int* P,
                      const double* U, const double* E, double g_re, double g_im, double alpha_re, double alpha_im);
int alf_b200_set_op_t(alf_b200_handle* h, int nc, int nf, int N, int diag, const int* P, const double* U, const double* E,
                      double g_re, double g_im);
/* Time-dependent coupling of an interaction vertex: Op_V(n,nf)%g_t(1:Ltrot) (Prog/Operator_mod.F90:66; used instead of g by Op_mmultL/R, Op_Wrapup/do,
 * Upgrade2 and Op_phase, :175,577,671,768-791,885-908).  g_t: Ltrot complex values (re, im).  After alf_b200_set_op_v of the same vertex, before
 * alf_b200_finalize_model; the vertex tables are then built once per time slice.  Not combined with Ising action tables / global-in-slice moves, nor with
 * the Langevin / HMC updates (ALF_ERROR_UNSUPPORTED). */
int alf_b200_set_op_v_gt(alf_b200_handle* h, int n, int nf, const double* g_t);
/* Projective algorithm (Projector = .T.): Thtrot and N_part, then WF_L(nf)%P / WF_R(nf)%P (Ndim x N_part, column-major
 * complex) per flavor -- the public module variables read by Prog/main.F90:366-376,596-599.  Call before finalize_model;
 * UDV_State then carries U(Ndim, N_part) without V (Prog/udv_state_mod.F90:131-150), CGR dispatches to CGRP
 * (Prog/cgr1_mod.F90:207-211, 464-515) and the sweep calls Tau_p instead of TAU_M (Prog/main.F90:829-831, 884-886). */
int alf_b200_set_projector(alf_b200_handle* h, int thtrot, int n_part);
int alf_b200_set_trial_wf(alf_b200_handle* h, int nf, const double* P_L, const double* P_R);
/* ham%S0 for Ising actions as tables (Prog/Hamiltonians/Hamiltonian_Z2_Matter_smod.F90:439-512, tables :841-857), and main.F90's
 * Propose_S0 (Prog/Wrapgr_mod.F90:127-132).  Field n (1-based) owns the terms op_start[n-1] .. op_start[n]-1; term t lists the entries
 * term_start[t] .. term_start[t+1]-1, each a (field entry_op, 1-based; time-slice offset entry_dt) pair; the product of the CURRENT +-1
 * values of those fields selects the flip ratio w[2t] (product -1) or w[2t+1] (product +1); S0(n, nt) is the product over the field's
 * terms (the field itself is listed where the coupling contains it, as in DW_Ising_tau(nsigma(n,nt)*nsigma(n,nt+1))).  Time offsets
 * wrap periodically; open_boundaries = 1 drops terms that leave 1..Ltrot (projective algorithm, :465-472).  n_terms = 0: S0 = 1.
 * Must be called before alf_b200_finalize_model.  With tables or Propose_S0 the slice is visited by the per-visit kernel. */
int alf_b200_set_s0_ising(alf_b200_handle* h, int n_terms, const int* op_start /* n_opv+1 */, const int* term_start /* n_terms+1 */,
                          const int* entry_op, const int* entry_dt, const double* w /* 2*n_terms */, int open_boundaries, int propose_s0);
/* Nt_sequential_start, Nt_sequential_end, N_Global_tau of Prog/main.F90 (ham%Overide_global_tau_sampling_parameters): WRAPGRUP / WRAPGRDO visit
 * the fields nt_sequential_start .. nt_sequential_end one by one and then (before, on the way down) make n_global_tau global-in-slice moves
 * (Prog/Wrapgr_mod.F90:112-153, 189-194).  Defaults: 1, size(Op_V,1), 0.  Before alf_b200_finalize_model. */
int alf_b200_set_global_tau_sampling(alf_b200_handle* h, int nt_sequential_start, int nt_sequential_end, int n_global_tau);
/* LOBS_ST, LOBS_EN of the VAR_QMC namelist (Prog/QMC_runtime_var_mod.F90:156-189, handed to ham%Obser's call sites Prog/main.F90:757-773,789-802 and
 * to TAU_M / Tau_p): equal-time measurements are taken on the slices lobs_st <= NTAU1 <= lobs_en.  0 = the reference's default (1 and Ltrot; Thtrot + 1
 * and Ltrot - Thtrot in the projective algorithm); a projective window outside [Thtrot + 1, Ltrot - Thtrot] is an error as in the reference. */
int alf_b200_set_measure_interval(alf_b200_handle* h, int lobs_st, int lobs_en);
/* The plugin callback ham%Global_move_tau for Ising "star" moves as tables (Prog/Hamiltonians/Hamiltonian_Z2_Matter_smod.F90:535-643), so that
 * the batched sweep needs no host in the loop: per move a site I = nranf(n_sites) is drawn from the chain's stream; Flip_list =
 * move_fields[move_start[I-1] .. move_start[I]-1] (1-based field indices, ascending = after Wrapgr_sort, at most 16, Ising fields only),
 * Flip_value = the flipped fields; S0_ratio = product of the site's coupling terms (site_term_start / term_start / entry_op / entry_dt / w
 * exactly as in alf_b200_set_s0_ising); T0_Proposal = 1 - 1/(1 + S0_ratio) is tested against one ranf() and T0_Proposal_ratio =
 * 1/S0_ratio or 0 (:633-641).  Wrapgr_Random_update then proceeds on the device as in alf_b200_wrapgr_random_update.
 * Before alf_b200_finalize_model. */
int alf_b200_set_global_move_tau_ising(alf_b200_handle* h, int n_sites, const int* move_start /* n_sites+1 */, const int* move_fields, int n_terms,
                                       const int* site_term_start /* n_sites+1 */, const int* term_start /* n_terms+1 */, const int* entry_op,
                                       const int* entry_dt, const double* w /* 2*n_terms */, int open_boundaries);
/* Continuous Hubbard-Stratonovich fields (Op_V%type = 3, e.g. Predefined_Int_U_MZ_continuous_HS / _U_SUN_continuous_HS, Prog/Predefined_Int_mod.F90:
 * 90-105,160-181; single-site vertices): nsigma%f is real, phi = f, gamma = 1 (Prog/Fields_mod.F90:112-170), the proposal is
 * f + Amplitude (ranf - 1/2) (:185-186), exp(g phi O) is evaluated on the fly (Op_exp, Prog/Operator_mod.F90:585-600).  The Gaussian weight of
 * the transformation is the plugin's ham%S0 = exp(-(f'^2 - f^2)/2) (Prog/Hamiltonians/Hamiltonian_Hubbard_smod.F90:880-882): switch it on here.
 * Amplitude is Fields_mod's module variable (default 1). */
int alf_b200_set_s0_gaussian(alf_b200_handle* h, int on);
int alf_b200_set_amplitude(alf_b200_handle* h, double amplitude);
int alf_b200_finalize_model(alf_b200_handle* h);
int alf_b200_is_complex(const alf_b200_handle* h);   /* 1 if the complex instantiation was selected */

/* ---- random numbers and fields.  Replaces Ranset per rank (Libraries/Modules/random_wrap_mod.F90:52-80,
 * Prog/Set_random_mod.F90:48-106) and Fields_set / nsigma%f access (Prog/Fields_mod.F90:588-610). */
int alf_b200_set_seeds(alf_b200_handle* h, const int32_t* seeds /* n_chains */);
int alf_b200_get_rng_state(alf_b200_handle* h, uint64_t* state /* n_chains*4 */);
int alf_b200_set_rng_state(alf_b200_handle* h, const uint64_t* state);
int alf_b200_fields_set(alf_b200_handle* h);                           /* random start, draws from each chain's stream */
int alf_b200_set_fields(alf_b200_handle* h, const double* f /* complex [chain][nt][n] */);
int alf_b200_get_fields(alf_b200_handle* h, double* f);

/* ---- batched mode: the body of the get_sequential() branch of Prog/main.F90 for all chains of the handle. */
int alf_b200_init_sweep(alf_b200_handle* h);                 /* main.F90:589-631: storage fill (WRAPUL), G(0) = CGR, Phase */
int alf_b200_sweep(alf_b200_handle* h, int n_sweeps, int ltau);  /* main.F90:714-887 (+ TAU_M when ltau = 1) */
/* end-to-end form with HOST buffers: upload fields -> n_sweeps sweeps -> download fields, observables, counters */
int alf_b200_sweep_host(alf_b200_handle* h, int n_sweeps, int ltau, const double* fields_in, double* fields_out,
                        double* obs_out /* alf_b200_obs_size() doubles */, double* control_out /* 16 doubles */);

/* ---- compat mode: one reference routine per call, on all chains (argument meaning as in the reference). */
int alf_b200_wrapgrup(alf_b200_handle* h, int ntau);          /* Prog/Wrapgr_mod.F90:81  (NTAU in 0..Ltrot-1) */
int alf_b200_wrapgrdo(alf_b200_handle* h, int ntau);          /* Prog/Wrapgr_mod.F90:160 (NTAU in Ltrot..1)   */
int alf_b200_wrapur(alf_b200_handle* h, int ntau, int ntau1); /* Prog/wrapur_mod.F90:37  on udvr */
int alf_b200_wrapul(alf_b200_handle* h, int ntau1, int ntau); /* Prog/wrapul_mod.F90:36  on udvl */
int alf_b200_udv_reset(alf_b200_handle* h, int which /*0 udvl,1 udvr*/, char side);   /* udv_state_mod.F90:224 */
int alf_b200_cgr(alf_b200_handle* h, int nvar);               /* Prog/cgr1_mod.F90:36: GR, Phase from udvr, udvl */
int alf_b200_tau_m(alf_b200_handle* h);                       /* Prog/tau_m_mod.F90:56 */
int alf_b200_tau_p(alf_b200_handle* h, int nst_in);           /* Prog/tau_p_mod.F90:74 (projector; udvr, udvst, GR as in main.F90:829) */

/* Device-side time-displaced lattice observables: what ham%ObserT of the shipped Hamiltonians accumulates through
 * Predefined_Obs_tau_Green / SpinMz / SpinSUN / Den_measure (Prog/Predefined_Obs_mod.F90:337-594) each time TAU_M / Tau_p reach a
 * time point (tau_m_mod.F90:115-177, tau_p_mod.F90:173-250), with Hop_mod_Symm applied first when Symm (as the reference does).
 * set_lattice: List(I1,1:2) (unit cell, orbital; 1-based; cell <= 0 = site not measured) and Latt%imj (n_unit x n_unit, column-major,
 * 1-based; Libraries/Modules/lattices_v3_mod.F90:316-330).  Accumulators are sums over the chains of the handle:
 * acc[ch][nt][no_J][no_I][imj] complex with ch = 0 Green, 1 SpinZ, 2 SpinXY (Mz only), 3 Den; bg[which][nt][no] complex (Obs_Latt0 of
 * SpinZ, Den); cnt = {N (counted at nt = 0), sum of signs}.  ham%ObserT of other models stays a host callback (taum_capture). */
int alf_b200_set_lattice(alf_b200_handle* h, int n_unit, int norb, const int* site_cell, const int* site_orb, const int* imj);
int alf_b200_obs_tau_enable(alf_b200_handle* h, int on);
int alf_b200_obs_tau_reset(alf_b200_handle* h);
/* equal-time variants (Predefined_Obs_eq_Green / SpinMz / SpinSUN / Den_measure, Predefined_Obs_mod.F90:77-325) accumulated at every
 * measured slice of the sweep where main.F90:757-773,789-802 call ham%Obser; same layout with a single time point; cnt[0] counts
 * chain-measurements */
int alf_b200_obs_eq_enable(alf_b200_handle* h, int on);
int alf_b200_get_obs_eq(alf_b200_handle* h, double* acc, double* bg, double* cnt);
int alf_b200_obs_tau_dims(const alf_b200_handle* h, int* n_channels, int* ntau, int* norb, int* n_unit);
int alf_b200_get_obs_tau(alf_b200_handle* h, double* acc, double* bg, double* cnt);

/* Langevin updates of continuous fields, Prog/Langevin_HMC_mod.F90 (scheme "Langevin"; single-site type-3 vertices, e.g. Hamiltonian_Hubbard with
 * Continuous = .true.; the state must be the one alf_b200_init_sweep or a finished sweep leaves):
 *  - alf_b200_langevin_forces: Langevin_HMC_Forces (:107-191, without its measurements) -> the fermionic forces dS_F/dphi of every chain,
 *    complex [chain][nt][n]; GR, udvr, udvl and Phase advance as in the reference (upward pass with stabilisation, storage udvst untouched).
 *  - alf_b200_langevin_update: Langevin_HMC_update (:355-392): forces, Forces_0 = dS_0/dphi = phi (Ham_Langevin_HMC_S0 of the Gaussian action,
 *    Hamiltonian_Hubbard_smod.F90:896-915), Delta_t_running = delta_t, reduced to max_force * delta_t / max|force| when a force exceeds
 *    max_force, phi -= (Forces_0 + Re(Phase F)/Re(Phase)) Delta_t_running - sqrt(2 Delta_t_running) rang() with rang_wrap's Box-Muller draws from
 *    the chain's stream (n outer, nt inner), then Langevin_HMC_Reset_storage (:228-285).  delta_t_running [chain] may be NULL. */
/*  - alf_b200_hmc_update: scheme "HMC" (:393-571) with L_Forces = .false. and the base Apply_B_HMC (identity): momenta from rang_wrap (nt outer, n inner),
 *    Leapfrog_Steps leapfrog steps (storage reset + force pass each), Compute_Fermion_Det before and after, Compute_Ratio_Global
 *    (Prog/Global_mod.F90:651-760) with the Gaussian Get_Delta_S0_global, Weight > ranf per chain, rejected chains restored, storage reset.
 *    All vertices must carry type-3 fields.  weight [chain], accepted [chain] may be NULL. */
int alf_b200_hmc_update(alf_b200_handle* h, double delta_t, int leapfrog_steps, double* weight, uint8_t* accepted);
int alf_b200_langevin_forces(alf_b200_handle* h, double* forces /* complex n_chains*ltrot*n_opv */);
int alf_b200_langevin_update(alf_b200_handle* h, double delta_t, double max_force, double* delta_t_running);

/* Compute_Fermion_Det(Phase_det, Det_Vec, udvl, udvst, Stab_nt, storage = "Empty"), Prog/Global_mod.F90:792-1000 (what Global_Updates :450 and
 * the tempering exchange :108 weigh configurations with): rebuilds the left propagation from the CURRENT fields (udvl and the storage udvst
 * are overwritten as after main.F90:589-627; GR is not touched) and returns, per chain and flavor ([chain][nf]), log|det| = sum_I Det_Vec(I, nf)
 * and Phase_det(nf) (complex).  Finite temperature: det(1 + B(beta, 0)); projector: det(P_L^H B(2 theta + beta, 0) P_R). */
int alf_b200_compute_fermion_det(alf_b200_handle* h, double* log_abs_det /* n_chains*n_fl */, double* phase_det /* complex n_chains*n_fl */);

/* Global-in-slice moves (N_Global_tau > 0), Prog/Wrapgr_mod.F90:247-433.  ham%Global_move_tau stays a host plugin callback:
 * its outputs (Flip_length, Flip_list (1-based), Flip_value, T0_Proposal_ratio, S0_ratio) are passed for every chain and every
 * one of the n_moves proposals ([chain][move], lists [chain][move][maxlen], maxlen <= 16); the device sorts the lists
 * (Wrapgr_sort), runs PlaceGR / Op_Wrapup / Upgrade2 ("Intermediate", "Final") with the chain's own random stream and rolls
 * back rejected multi-field moves (GR_st).  The operator position m of GR inside the slice is per-chain device state:
 * set it after WRAPGRUP (m = size(Op_V,1)) or before WRAPGRDO's sequential part; place_to >= 0 appends the final
 * Wrapgr_PlaceGR(GR, m, place_to, ntau) of WRAPGRUP/WRAPGRDO (:148-153, :190-195). */
int alf_b200_wrapgr_set_position(alf_b200_handle* h, int m);
int alf_b200_wrapgr_get_position(alf_b200_handle* h, int* m /* [chain] */);
int alf_b200_wrapgr_placegr(alf_b200_handle* h, int m1, int ntau);            /* Wrapgr_PlaceGR, Prog/Wrapgr_mod.F90:247 */
int alf_b200_wrapgr_random_update(alf_b200_handle* h, int ntau, int n_moves, int maxlen, const int* flip_length, const int* flip_list,
                                  const double* flip_value /* complex */, const double* t0_proposal_ratio, const double* s0_ratio,
                                  uint8_t* accepted /* [chain][move]: 1 accepted, 0 rejected, 2 not proposed (T0 ratio <= 1e-7); may be NULL */,
                                  int place_to);                             /* Wrapgr_Random_update, Prog/Wrapgr_mod.F90:317 */

/* ---- results */
int alf_b200_get_green(alf_b200_handle* h, int chain, int nf, int symmetrize, double* out /* complex N*N */);
int alf_b200_set_green(alf_b200_handle* h, int chain, int nf, const double* in);
int alf_b200_get_phase(alf_b200_handle* h, double* out /* complex [chain] */);
int alf_b200_get_udv(alf_b200_handle* h, int which /*0 udvl,1 udvr,2 udvst*/, int nst, int chain, int nf,
                     double* U, double* D, double* V /* complex */);
/* host UDV_State (type UDV_State, Prog/udv_state_mod.F90:85-110: U, D, V as complex arrays) -> the handle's udvl / udvr / udvst(nst) of one chain and flavor:
 * compat mode, so that WRAPUR(NTAU, NTAU1, UDVR) / WRAPUL / CGR(PHASE, NVAR, GRUP, udvr, udvl) act on the caller's states (set, call, get). */
int alf_b200_set_udv(alf_b200_handle* h, int which, int nst, int chain, int nf, const double* U, const double* D, const double* V);
/* control accumulators (Prog/control_mod.F90:53-71), reduced over chains:
 * [0] XMEANG sum [1] XMAXG [2] NCG [3] XMAXP [4] XMEAN_tau sum [5] XMAX_tau [6] NCG_tau [7] NC_up [8] ACC_up
 * [9] NC_eff_up [10] ACC_eff_up [11] NaN flag [12] unstable flag (XMAX > 10)
 * [13] (measurement support, no reference counterpart) number of rank-KD rewrites of G in global memory by the slice kernels */
int alf_b200_get_control(alf_b200_handle* h, double* out /* 16 */);
int alf_b200_accept_log(alf_b200_handle* h, int n_sweeps);   /* record accept/reject per field visit for the next n_sweeps sweeps (check 2); 0 = off */
int alf_b200_get_accept_log(alf_b200_handle* h, uint8_t* out, long cap, long* n_per_chain);
int alf_b200_taum_capture(alf_b200_handle* h, int every);    /* keep GT0,G0T,G00,GTT handed to ObserT every k-th slice */
int alf_b200_get_taum(alf_b200_handle* h, int chain, double* out, long cap_complex, long* n_complex);
/* the four matrices right after every CGR2_2 of TAU_M (freshly recomputed; not symmetrised), same layout */
int alf_b200_get_taum_fresh(alf_b200_handle* h, int chain, double* out, long cap_complex, long* n_complex);

/* ---- device-side scalar observables accumulated where main.F90 calls ham%Obser (:757-773,:789-802), summed over the chains of the handle.
 * Layout (alf_b200_obs_size() = 16 doubles): [0] N_meas (chain-measurements), [1] sum ZS, [2..3] Part, [4..5] Kin, [6..7] Pot, [8..9] Ener
 * (re, im each; every entry is sum Obs ZP ZS with ZP = Phase / Re Phase, ZS = sign Re Phase: Hamiltonian_Hubbard_smod.F90:711-772).
 * Part = N_SUN sum_nf sum_i GRC(i,i,nf) needs no model input.  Kin, Pot and Ener = Kin + Pot are model specific and are handed over as
 * tables, evaluated on the Green function ham%Obser receives (Hop_mod_Symm(GR) when Symm, main.F90:761-764), with
 * GRC(i,j,nf) = delta_ij - GR(j,i,nf):
 *   Kin = N_SUN sum_t kin_coef[t] GRC(kin_i[t], kin_j[t], kin_nf[t])        Predefined_Hoppings_Compute_Kin, Prog/Predefined_Hop_mod.F90:1850-1959:
 *                                                                            per bond Z GRC(I1,J1,nf) + conj(Z) GRC(J1,I1,nf), per site T_Loc GRC(I1,I1,nf)
 *   Pot = sum_t pot_coef[t] GRC(i1,i1,nf1) GRC(i2,i2,nf2)                   Hamiltonian_Hubbard_smod.F90:744-760 (ham_U n_up n_down), Kondo / tV likewise
 * Indices and flavors 1-based, coefficients complex.  n_kin = n_pot = 0 switches the tables off. */
int alf_b200_set_obs_scal_tables(alf_b200_handle* h, int n_kin, const int* kin_i, const int* kin_j, const int* kin_nf, const double* kin_coef,
                                 int n_pot, const int* pot_i1, const int* pot_nf1, const int* pot_i2, const int* pot_nf2, const double* pot_coef);
int alf_b200_obs_size(const alf_b200_handle* h);
int alf_b200_obs_reset(alf_b200_handle* h);
int alf_b200_obs_device_ptr(alf_b200_handle* h, double** dptr, long* n_doubles);  /* for the NCCL bin reduction */
int alf_b200_get_obs(alf_b200_handle* h, double* out);

/* ---- multi-GPU: the per-bin reduction that replaces MPI_REDUCE in Print_bin_Vec / Print_bin_Latt / Print_bin_Latt_Local and Control_Print
 * (Prog/observables_mod.F90:425-438, 648-653, 828-834; Prog/control_mod.F90:397-452).  One process per GPU; chains never talk during a sweep.
 *   alf_b200_comm_unique_id : rank 0 creates the 128-byte NCCL id; the host program hands it to the other ranks (ALF: MPI_BCAST; here any channel);
 *   alf_b200_comm_init      : every rank joins with its handle (one communicator per handle; NCCL is resolved with dlopen("libnccl.so.2"));
 *   alf_b200_reduce_bins    : SUM-reduces IN PLACE on the device, over NVLink, every observable accumulator of the handle -- the scalars of
 *                             alf_b200_get_obs, the equal-time and time-displaced lattice accumulators with their backgrounds and counters -- to
 *                             rank `root`, where alf_b200_get_obs / _get_obs_eq / _get_obs_tau then return the sum over all ranks' chains;
 *   alf_b200_reduce_control : the 16 control values of alf_b200_get_control reduced to `root` (SUM, and MAX for XMAXG, XMAXP, XMAX_tau and the flags).
 * Without a communicator (single rank) the two reductions are no-ops / plain reads. */
int alf_b200_comm_unique_id(char* id /* 128 bytes */);
int alf_b200_comm_init(alf_b200_handle* h, int nranks, int rank, const char* id /* 128 bytes */);
int alf_b200_comm_destroy(alf_b200_handle* h);
int alf_b200_reduce_bins(alf_b200_handle* h, int root);
int alf_b200_reduce_control(alf_b200_handle* h, int root, double* out /* 16 */);

/* ---- UDV_Wrap_Pivot(A, U, D, V, NCON, N1, N2), Prog/UDV_WRAP_mod.F90:125-208 (default variant; the stabilisation of the STAB1 / STAB2
 * builds: wrapur_mod.F90:96, wrapul_mod.F90:98-102, cgr1_mod.F90:115,137): norm-sorted, norm-scaled columns, unpivoted Householder QR
 * (UDV_C, mymats_mod.F90:933), det V = 1.  A, U: complex n1*n2*batch, D: complex n2*batch, V: complex n2*n2*batch, A = U D V per matrix.
 * NCON (a diagnostic print in the reference) has no counterpart.  UDV_Wrap (QR + SVD, :212) is only called from Global_mod.F90:926,
 * outside the sweep, and is not provided. */
int alf_b200_udv_wrap_pivot(int device, int is_complex, int n1, int n2, int batch, const double* A, double* U, double* D, double* V);

/* ---- kernel-level entry points used by the parity tests (host arrays in, host arrays out; batch of matrices) */
int alf_b200_test_qdrp(int device, int is_complex, int m, int n, int batch, double* A /* complex m*n*batch, in/out */,
                       double* D /* n*batch */, int* jpvt /* n*batch, 1-based */, double* tau /* complex n*batch */,
                       double* phases /* 5*batch: perm sign, diag phase re/im, detq re/im */);
/* the blocked, windowed-pivoting QR used by the sweep for matrices beyond one SM's shared memory (alf_qrblk.cuh), plus the explicit
 * Q (complex m*m*batch) formed through the compact-WY application (replaces ZUNGQR) */
int alf_b200_test_qdrp_blocked(int device, int is_complex, int m, int n, int batch, double* A, double* D, int* jpvt, double* tau,
                               double* phases, double* Q);
int alf_b200_test_udv_decompose(int device, int is_complex, int n, int batch, char side, double* U, double* D, double* V);
/* CGR(PHASE, NVAR, GRUP, udvr, udvl), Prog/cgr1_mod.F90:36, on a batch of host states.  detUR / detUL (complex per matrix: det U_R, det U_L, which the
 * sweep carries along) may be NULL: they are then computed on the device, so the call needs nothing beyond the reference routine's own arguments. */
int alf_b200_test_cgr(int device, int is_complex, int n, int batch, int nvar, int stab, const double* UR, const double* DR,
                      const double* VR, const double* UL, const double* DL, const double* VL, const double* detUR,
                      const double* detUL, double* G, double* phase);
/* CGR2_2 (Prog/cgr2_2_mod.F90:196): udv2 = right (side R), udv1 = left (side L) propagation; out4 = GRT0, GR00, GRTT, GR0T,
 * each complex n*n*batch */
/* CGRP (Prog/cgr1_mod.F90:464) on explicit U_R, U_L (n x n_part each): G (n x n) and the phase of det(U_L^H U_R) */
int alf_b200_test_cgrp(int device, int is_complex, int n, int n_part, int batch, const double* UR, const double* UL, double* G, double* phase);
int alf_b200_test_cgr2_2(int device, int is_complex, int n, int batch, int stab, const double* U2, const double* D2, const double* V2,
                         const double* U1, const double* D1, const double* V1, double* out4);
int alf_b200_test_gemm(int device, int is_complex, int ta, int tb, int m, int n, int k, int batch, const double* A,
                       const double* B, double* C);
int alf_b200_hop_apply(alf_b200_handle* h, int which /* 0 mmthr,1 mmthr_m1,2 mmthl,3 mmthl_m1,4 mmthlc,5 Symm */,
                       int nf, double* A /* complex N*N, in/out */);
/* ---- measurement support (bench.py): the handle's CUDA stream (for CUDA-event timing on the launching stream), and per-kernel
 * launch counts / device time.  Categories: 0 update (k_wrapgr) 1 op-list wrap (k_apply_ops) 2 pivoted QR 3 form-Q 4 GEMM 5 TRSM
 * 6 element-wise 7 observables.  mask bit c = 1 records CUDA events around every launch of category c. */
#define ALF_B200_NKCAT 8
int alf_b200_get_stream(alf_b200_handle* h, void** stream);
int alf_b200_kernel_timing(alf_b200_handle* h, unsigned mask);        /* resets the statistics */
int alf_b200_get_kernel_stats(alf_b200_handle* h, double* ms /* 8 */, long* launches /* 8 */);
int alf_b200_get_kernel_flops(alf_b200_handle* h, double* flops /* 8: algorithmic FP64 flops per kernel category since the last reset */);
int alf_b200_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops);   /* roofline denominator microbenchmark */

#ifdef __cplusplus
}
#endif
#endif
