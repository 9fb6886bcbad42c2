#pragma once
// alf_engine.cuh -- handle, model tables and the sweep schedule, templated on the arithmetic type (instantiated for double in
// alf_inst_real.cu and for cplx in alf_inst_cplx.cu; the C-ABI of include/alf_b200.h lives in alf_b200.cu).
// The schedule mirrors Prog/main.F90:446-457,589-631,714-887; every chain of the handle follows the same
// stabilisation schedule, so each reference routine becomes ONE batched launch sequence over (chain, flavor).
// There is no CPU fallback: without a CUDA device every entry point fails with ALF_ERROR_CUDA.
#include <complex>
#include <cstring>
#include <algorithm>
#include <memory>
#include "../../include/alf_b200.h"
#include "alf_la_host.cuh"
#include "alf_ops.cuh"
#include "alf_update.cuh"
#include "alf_update_fast.cuh"
#include "alf_global_move.cuh"
#include "alf_obs_tau.cuh"
#include "alf_udv_wrap.cuh"

typedef std::complex<double> cd;
static const double kEpsMachine = 2.220446049250313e-16;

// ------------------------------------------------------------------------------------------------ host model
struct HostOp {
  int N = 0, nnz = 0, diag = 0, type = 0; bool set = false;
  std::vector<int> P; std::vector<cd> U; std::vector<double> E; cd g = 0, alpha = 0;
  std::vector<cd> g_t;          // time-dependent coupling g_t(1..Ltrot) (Operator_mod.F90:66), empty if not allocated
};

static void host_op_exp(cd g, const HostOp& op, std::vector<cd>& Mat) {   // Prog/Operator_mod.F90:491-529 (Kahan summation)
  const int N = op.N; Mat.assign((size_t)N * N, cd(0, 0));
  if (op.diag) { for (int n = 0; n < N; ++n) Mat[n + (size_t)n * N] = std::exp(g * op.E[n]); return; }
  std::vector<cd> c((size_t)N * N, cd(0, 0));
  for (int n = 0; n < N; ++n) {
    cd Z = std::exp(g * op.E[n]);
    for (int J = 0; J < N; ++J) {
      cd Z1 = Z * std::conj(op.U[J + (size_t)n * N]);
      for (int I = 0; I < N; ++I) {
        cd y = Z1 * op.U[I + (size_t)n * N] - c[I + (size_t)J * N];
        cd t = Mat[I + (size_t)J * N] + y;
        c[I + (size_t)J * N] = (t - Mat[I + (size_t)J * N]) - y;
        Mat[I + (size_t)J * N] = t;
      }
    }
  }
}

struct HostExpT { int N; std::vector<int> P; std::vector<cd> mat, invmat, mat12, invmat12; bool active; };

static void host_expopt(const HostOp& op, HostExpT& e) {   // Prog/OpTTypes_mod.F90:92-133, 203-234
  const int N = op.N; e.N = N; e.P = op.P;
  host_op_exp(op.g, op, e.mat); host_op_exp(-op.g, op, e.invmat); host_op_exp(op.g / 2.0, op, e.mat12); host_op_exp(-op.g / 2.0, op, e.invmat12);
  bool real = std::abs(op.g.imag()) == 0.0;
  for (auto& u : op.U) if (u.imag() != 0.0) real = false;
  auto sym = [&](std::vector<cd>& M) {
    if (real) for (auto& z : M) z = cd(z.real(), 0.0);
    for (int i = 0; i < N; ++i) for (int j = i; j < N; ++j) M[i + (size_t)j * N] = (M[i + (size_t)j * N] + std::conj(M[j + (size_t)i * N])) / 2.0;
    for (int i = 0; i < N; ++i) for (int j = i + 1; j < N; ++j) M[j + (size_t)i * N] = std::conj(M[i + (size_t)j * N]);   // 'U' storage
  };
  sym(e.mat); sym(e.invmat); sym(e.mat12); sym(e.invmat12);
  double g2 = real ? op.g.real() * op.g.real() : std::norm(op.g);
  e.active = g2 > kEpsMachine;
}

static double phi_st(int type, int s) {   // Prog/Fields_mod.F90:270-279
  if (type == 1) return (double)s;
  switch (s) { case -2: return -std::sqrt(2.0 * (3.0 + std::sqrt(6.0))); case -1: return -std::sqrt(2.0 * (3.0 - std::sqrt(6.0)));
               case 1: return std::sqrt(2.0 * (3.0 - std::sqrt(6.0))); case 2: return std::sqrt(2.0 * (3.0 + std::sqrt(6.0))); }
  return 0.0;
}
static double gama_st(int type, int s) {   // Fields_mod.F90:281-287
  if (type != 2) return 1.0;
  if (s == -2 || s == 2) return 1.0 - std::sqrt(6.0) / 3.0;
  if (s == -1 || s == 1) return 1.0 + std::sqrt(6.0) / 3.0;
  return 1.0;
}

template <typename T> T to_T(cd z);
template <> inline double to_T<double>(cd z) { return z.real(); }
template <> inline cplx to_T<cplx>(cd z) { return cplx(z.real(), z.imag()); }
template <typename T> cd from_T(T v);
template <> inline cd from_T<double>(double v) { return cd(v, 0.0); }
template <> inline cd from_T<cplx>(cplx v) { return cd(v.x, v.y); }

// one operator list under construction
struct ListBuild {
  long nt_stride = 0;                   // > 0: mat holds one table per time slice (g_t), nt_stride entries each
  int nvar = 1; std::vector<int> k, P, fidx; std::vector<cd> mat;   // mat: per op nvar*KMAX*KMAX
  std::vector<unsigned char> cont;      // 1: continuous field, slot 0 of mat holds the coefficient c of exp(c phi)
  void add(int kk, const int* p, int fi, const std::vector<std::vector<cd>>& mats /* nvar matrices kk x kk col-major */, bool is_cont = false) {
    k.push_back(kk); fidx.push_back(fi); cont.push_back(is_cont ? 1 : 0);
    for (int a = 0; a < ALF_KMAX; ++a) P.push_back(a < kk ? p[a] : 0);
    for (int v = 0; v < nvar; ++v) for (int b = 0; b < ALF_KMAX; ++b) for (int a = 0; a < ALF_KMAX; ++a)
      mat.push_back((a < kk && b < kk) ? mats[v][a + (size_t)b * kk] : cd(0, 0));
  }
  // greedy levels: consecutive operators with pairwise disjoint support (keeps the sequential semantics exactly)
  // (a level is additionally cut after max_ops operators: the kernel's descriptor buffers hold one chunk)
  std::vector<int> levels(int ndim, int max_ops) const {
    std::vector<int> ls; ls.push_back(0); std::vector<char> used(ndim, 0);
    if (k.empty()) return ls;                       // no chunk at all
    int in_level = 0;
    for (size_t o = 0; o < k.size(); ++o) {
      bool clash = in_level >= max_ops;
      for (int a = 0; a < k[o]; ++a) if (used[P[o * ALF_KMAX + a]]) clash = true;
      if (clash) { ls.push_back((int)o); std::fill(used.begin(), used.end(), 0); in_level = 0; }
      for (int a = 0; a < k[o]; ++a) used[P[o * ALF_KMAX + a]] = 1;
      ++in_level;
    }
    ls.push_back((int)k.size());
    return ls;
  }
};

// ------------------------------------------------------------------------------------------------ engine
struct EngineBase {
  virtual ~EngineBase() {}
  virtual void init_sweep() = 0;
  virtual void sweep(int ltau) = 0;
  virtual void wrapgrup(int ntau) = 0;
  virtual void wrapgrdo(int ntau) = 0;
  virtual void wrapur(int ntau, int ntau1) = 0;
  virtual void wrapul(int ntau1, int ntau) = 0;
  virtual void udv_reset(int which, char side) = 0;
  virtual void cgr_call(int nvar) = 0;
  virtual void tau_m() = 0;
  virtual void tau_p(int nst_in) = 0;
  virtual void obs_tau_setup() = 0;
  virtual void gm_set_position(int m) = 0;
  virtual void gm_get_position(int* m) = 0;
  virtual void gm_random_update(int ntau, int n_moves, int maxlen, const int* len, const int* list0, const int8_t* val, const double* t0, const double* s0,
                                uint8_t* acc_out, int place_to) = 0;
  virtual void get_green(int chain, int nf, int symm, cd* out) = 0;
  virtual void set_green(int chain, int nf, const cd* in) = 0;
  virtual void get_udv(int which, int nst, int chain, int nf, cd* U, cd* D, cd* V) = 0;
  virtual void set_udv(int which, int nst, int chain, int nf, const cd* U, const cd* D, const cd* V) = 0;
  virtual void hop_apply(int which, int nf, cd* A) = 0;
  virtual void fermion_det(double* logdet, cd* phase) = 0;
  virtual void langevin_get_forces(cd* out) = 0;
  virtual void langevin_update(double delta_t, double max_force, double* dt_running) = 0;
  virtual void hmc_update(double delta_t, int leapfrog_steps, double* weight, unsigned char* acc) = 0;
  virtual void sync() = 0;
};

struct alf_b200_handle {
  int ndim, n_fl, n_sun, ltrot, nwrap, n_opv, n_opt, symm, stab, n_chains, device;
  std::vector<HostOp> opv, opt;
  bool finalized = false, is_complex = false;
  std::unique_ptr<EngineBase> eng;
  std::string err;
  cudaStream_t stream = 0;
  Prof prof;
  int8_t* pin_fields = nullptr;      // pinned staging buffer of sweep_host
  // untyped device state shared with the engine
  double* d_fields_c = nullptr;      // continuous fields (type 3 vertices), same [chain][nt][n] layout; nullptr without such vertices
  bool s0_gaussian = false; double amplitude = 1.0;   // ham%S0 = exp(-(phi'^2 - phi^2)/2) of the continuous HS transformation; Fields_mod Amplitude
  int8_t* d_fields = nullptr; uint64_t* d_rng = nullptr; cplx* d_phase = nullptr; unsigned long long* d_counters = nullptr;
  double* d_ctl = nullptr;            // per chain: 0 XMEANG 1 XMAXG 2 NCG 3 XMAXP 4 XMEAN_tau 5 XMAX_tau 6 NCG_tau 7 flags(nan=1, unstable=2)
  uint8_t* d_acclog = nullptr; long acclog_per_chain = 0; long acclog_pos = 0; bool acclog_on = false;
  double* d_obs = nullptr; int obs_size = 0;
  void* comm = nullptr; int comm_nranks = 1, comm_rank = 0; double* d_ctlred = nullptr; long n_reduce_calls = 0;   // NCCL communicator of the bin reduction (alf_b200_comm_init)
  // model-specific scalar observables as tables (alf_b200_set_obs_scal_tables): Kin terms (i, j, nf, coef), Pot terms (i1, nf1, i2, nf2, coef); 0-based
  int n_kin = 0, n_pot = 0; int *d_kin_idx = nullptr, *d_pot_idx = nullptr; cplx *d_kin_coef = nullptr, *d_pot_coef = nullptr;
  int taum_every = 0; std::vector<std::vector<cd>> taum_host, taum_fresh_host;   // per chain captured matrices
  std::vector<int> types;            // operator type per n
  // lattice tables for the device-side lattice observables (0-based; site -> unit cell / orbital, imj(I,J) column-major)
  int n_unit = 0, norb = 1; std::vector<int> site_cell, site_orb, imj;
  bool obs_tau_on = false; double *d_obst_acc = nullptr, *d_obst_bg = nullptr, *d_obst_cnt = nullptr; int obst_ntau = 0;
  bool obs_eq_on = false; double *d_obse_acc = nullptr, *d_obse_bg = nullptr, *d_obse_cnt = nullptr;      // equal-time lattice observables (one time point)
  // table-driven Ising action for ham%S0 (see S0TabDev) and main.F90's Propose_S0
  bool s0_on = false; int s0_open_bc = 0, propose_s0 = 0; std::vector<int> s0_op_start, s0_term_start, s0_e_op, s0_e_dt; std::vector<double> s0_w;
  // Nt_sequential_start / _end, N_Global_tau (Overide_global_tau_sampling_parameters) and ham%Global_move_tau as tables (see GmtDev)
  int nt_seq_start = 1, nt_seq_end = -1, n_global_tau = 0;
  bool has_gt = false;                   // some Op_V carries g_t: vertex tables are built per time slice
  int lobs_st = 0, lobs_en = 0;          // 0 = default window (alf_b200_set_measure_interval)
  bool gmt_on = false; int gmt_n_sites = 0, gmt_open_bc = 0; std::vector<int> gmt_move_start, gmt_move_fields, gmt_op_start, gmt_term_start, gmt_e_op, gmt_e_dt; std::vector<double> gmt_w;
  // projective algorithm (Prog/Hamiltonian_main_mod.F90:181-197: Projector, Thtrot, WF_L, WF_R)
  bool projector = false; int thtrot = 0, n_part = 0; std::vector<std::vector<cd>> wf_l, wf_r;   // per flavor, Ndim x N_part column-major
};

static __global__ void k_ranset(uint64_t* rng, const int32_t* seeds, int n) {   // Ranset: random_wrap_mod.F90:52-80 with K = 8, N = 1
  int c = blockIdx.x * blockDim.x + threadIdx.x; if (c >= n) return;
  int32_t iseed0 = seeds[c], iseed = iseed0; int nn = 1;
  if (iseed == 0) { iseed = 8752143; nn = 0; }
  uint32_t v[8];
  for (int i = 1; i <= 8; ++i) {
    if (i <= nn) v[i - 1] = (uint32_t)iseed0;
    else { long long res = (long long)iseed; res = 62089911LL * res + 4349LL; iseed = (int32_t)(uint32_t)(unsigned long long)res; v[i - 1] = (uint32_t)iseed; }
  }
  uint64_t s[4];
  for (int i = 0; i < 4; ++i) s[i] = ((uint64_t)v[2 * i + 1] << 32) | (uint64_t)v[2 * i];
  if ((s[0] | s[1] | s[2] | s[3]) == 0) s[0] = 0x9E3779B97F4A7C15ULL;
  for (int i = 0; i < 4; ++i) rng[c * 4 + i] = s[i];
}

// Fields_set (Prog/Fields_mod.F90:588-610) for types < 4: f = +1, set to -1 if ranf() > 0.5; n fastest, then nt.
static __global__ void k_fields_set(int8_t* fields, uint64_t* rng, int n_chains, long per_chain) {
  int c = blockIdx.x * blockDim.x + threadIdx.x; if (c >= n_chains) return;
  Xoshiro r; r.s0 = rng[c * 4]; r.s1 = rng[c * 4 + 1]; r.s2 = rng[c * 4 + 2]; r.s3 = rng[c * 4 + 3];
  int8_t* f = fields + (long)c * per_chain;
  for (long i = 0; i < per_chain; ++i) f[i] = (r.ranf() > 0.5) ? -1 : 1;
  rng[c * 4] = r.s0; rng[c * 4 + 1] = r.s1; rng[c * 4 + 2] = r.s2; rng[c * 4 + 3] = r.s3;
}

// sum over (n, nt) of Im(g alpha phi(s)) per (chain, flavor)  -- Op_phase, Prog/Operator_mod.F90:160-181
static __global__ void k_op_phase(const int8_t* __restrict__ fields, const double* __restrict__ angle_tab, int F, int n_opv, int Ltrot, double* __restrict__ out,
                                  const double* __restrict__ fields_c, const unsigned char* __restrict__ is_cont, long nt_stride) {
  __shared__ double red[8];
  const int b = blockIdx.x, chain = b / F, f = b % F;
  const int8_t* fl = fields + (long)chain * Ltrot * n_opv;
  double s = 0.0;
  const double* fc = fields_c ? fields_c + (long)chain * Ltrot * n_opv : nullptr;
  for (long e = threadIdx.x; e < (long)Ltrot * n_opv; e += blockDim.x) { int n = (int)(e % n_opv); const double* at = angle_tab + (e / n_opv) * nt_stride;      // g_t: one table per time slice
    if (fc && is_cont[n]) s += at[((long)n * F + f) * ALF_NVAR] * fc[e]; else s += at[((long)n * F + f) * ALF_NVAR + fl[e] + 2]; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s += red[w]; out[b] = s; }
}

// Phase = (prod_f z_f e^{i angle_f})^N_SUN ; Control_PrecisionP (control_mod.F90:314-320)
static __global__ void k_phase_update(const cplx* __restrict__ z, const double* __restrict__ angle, int F, int n_sun, cplx* __restrict__ phase,
                               double* __restrict__ ctl, int compare, int n_chains) {
  int c = blockIdx.x * blockDim.x + threadIdx.x; if (c >= n_chains) return;
  cplx Z = cplx(1.0, 0.0);
  for (int f = 0; f < F; ++f) { cplx zz = z[c * F + f]; if (angle) { double a = angle[c * F + f]; zz = zz * cplx(cos(a), sin(a)); } Z = Z * zz; }
  cplx Zn = Z; for (int q = 1; q < n_sun; ++q) Zn = Zn * Z;
  if (compare) { double x = abs_(Zn - phase[c]); if (x > ctl[c * 8 + 3]) ctl[c * 8 + 3] = x; }
  phase[c] = Zn;
}

// fold k_compare output into the per-chain accumulators; which = 0: Control_PrecisionG, 1: Control_Precision_tau
static __global__ void k_ctl_accum(const double* __restrict__ cmp, int F, double* __restrict__ ctl, int which, int n_chains) {
  int c = blockIdx.x * blockDim.x + threadIdx.x; if (c >= n_chains) return;
  for (int f = 0; f < F; ++f) {
    double o[3] = {0.0, 0.0, 0.0};                                   // fold the CMP_SPLIT parts of the matrix
    for (int p = 0; p < CMP_SPLIT; ++p) { const double* q = cmp + ((long)(c * F + f) * CMP_SPLIT + p) * 3; o[0] = fmax(o[0], q[0]); o[1] += q[1]; if (q[2] != 0.0) o[2] = 1.0; }
    if (which == 0) {
      ctl[c * 8 + 0] += o[1]; if (o[0] > ctl[c * 8 + 1]) ctl[c * 8 + 1] = o[0]; ctl[c * 8 + 2] += 1.0;
      int flags = (int)ctl[c * 8 + 7]; if (o[2] != 0.0) flags |= 1; if (o[0] > 10.0) flags |= 2; ctl[c * 8 + 7] = (double)flags;
    } else { ctl[c * 8 + 4] += o[1]; if (o[0] > ctl[c * 8 + 5]) ctl[c * 8 + 5] = o[0]; ctl[c * 8 + 6] += 1.0; }
  }
}

// Model-independent scalar measurement at one time slice, where main.F90:757-773,789-802 call ham%Obser: per chain
// ZP = Phase/Re(Phase), ZS = sign(Re Phase) (Hamiltonian_Hubbard_smod.F90:577-580) and the particle number
// Part = N_SUN sum_nf sum_i (1 - G(i,i,nf)) -- a trace, hence identical for G and the symmetrised G~ = Hop_mod_Symm(G).
// obs: [0] N_meas (chain-slices), [1] sum ZS, [2..3] sum Part ZP ZS (re, im)
template <typename T>
__global__ void __launch_bounds__(128) k_obs_scalar(const T* __restrict__ G, long sM, int N, int F, int n_sun, const cplx* __restrict__ phase, double* __restrict__ obs) {
  __shared__ double red[2][4];
  const int c = blockIdx.x;
  cplx tr = cplx(0.0, 0.0);
  for (int e = threadIdx.x; e < F * N; e += blockDim.x) { const int f = e / N, i = e % N; const T g = G[((long)c * F + f) * sM + i + (long)i * N]; tr = tr + cplx(1.0 - real_(g), -imag_(g)); }
  double a = warp_sum(tr.x), b = warp_sum(tr.y);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = red[0][0] + red[0][1] + red[0][2] + red[0][3]; b = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const cplx ph = phase[c]; const double zs = (ph.x >= 0.0) ? 1.0 : -1.0; const cplx zp = cplx(1.0, ph.y / ph.x);
    const cplx v = (cplx(a * n_sun, b * n_sun) * zp) * zs;
    atomicAdd(obs + 0, 1.0); atomicAdd(obs + 1, zs); atomicAdd(obs + 2, v.x); atomicAdd(obs + 3, v.y);
  }
}

// Kin, Pot, Ener of ham%Obser from tables (see alf_b200_set_obs_scal_tables) on the Green function handed to ham%Obser; one CTA per chain.
// obs: [4..5] Kin, [6..7] Pot, [8..9] Ener, each sum over chains of value ZP ZS.
template <typename T>
__global__ void __launch_bounds__(256) k_obs_scal_tables(const T* __restrict__ G, long sM, int N, int F, int n_sun, const cplx* __restrict__ phase, int n_kin, const int* __restrict__ kin_idx,
                                                         const cplx* __restrict__ kin_coef, int n_pot, const int* __restrict__ pot_idx, const cplx* __restrict__ pot_coef, double* __restrict__ obs) {
  __shared__ double red[4][8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const T* Gc = G + (long)c * F * sM;
  auto grc = [&](int i, int j, int f) { const T g = Gc[(long)f * sM + j + (long)i * N]; return cplx(((i == j) ? 1.0 : 0.0) - real_(g), -imag_(g)); };      // delta_ij - GR(j,i)
  cplx kin = cplx(0.0, 0.0), pot = cplx(0.0, 0.0);
  for (int t = tid; t < n_kin; t += blockDim.x) kin = kin + kin_coef[t] * grc(kin_idx[3 * t], kin_idx[3 * t + 1], kin_idx[3 * t + 2]);
  for (int t = tid; t < n_pot; t += blockDim.x) { const int i1 = pot_idx[4 * t], f1 = pot_idx[4 * t + 1], i2 = pot_idx[4 * t + 2], f2 = pot_idx[4 * t + 3]; pot = pot + pot_coef[t] * (grc(i1, i1, f1) * grc(i2, i2, f2)); }
  double v[4] = {kin.x * n_sun, kin.y * n_sun, pot.x, pot.y};
#pragma unroll
  for (int q = 0; q < 4; ++q) { v[q] = warp_sum(v[q]); if ((tid & 31) == 0) red[q][tid >> 5] = v[q]; }
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < 4; ++q) { v[q] = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v[q] += red[q][w]; }
    const cplx ph = phase[c]; const double zs = (ph.x >= 0.0) ? 1.0 : -1.0; const cplx zpzs = cplx(zs, zs * ph.y / ph.x);
    const cplx k = cplx(v[0], v[1]) * zpzs, p = cplx(v[2], v[3]) * zpzs;
    atomicAdd(obs + 4, k.x); atomicAdd(obs + 5, k.y); atomicAdd(obs + 6, p.x); atomicAdd(obs + 7, p.y); atomicAdd(obs + 8, k.x + p.x); atomicAdd(obs + 9, k.y + p.y);
  }
}

static __global__ void k_i8_to_f64(const int8_t* __restrict__ a, double* __restrict__ b, long n) { for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = (double)a[i]; }
static __global__ void k_fill_int(int* __restrict__ p, int n, int v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
// Wrapgrup_Forces (Prog/Langevin_HMC_mod.F90:194-226) for single-site continuous vertices: with G already propagated through the slice,
// Forces(n, nt) = - sum_nf g_nf N_SUN ( O(1,1) (1 - G_nf(P, P)) + alpha_nf ).  (The conjugation with a diagonal vertex leaves the diagonal of G
// unchanged, so every force of the slice is read off one G.)  coef: per (n, f) [g O(1,1), g alpha] as T.
template <typename T>
__global__ void k_langevin_forces(const T* __restrict__ G, long sM, int N, int F, int n_sun, int M, const int* __restrict__ site /* [n][f] */, const T* __restrict__ coef /* [n][f][2] */,
                                  const unsigned char* __restrict__ is_cont, cplx* __restrict__ forces /* [chain][nt][n] */, int Ltrot, int nt) {
  const int chain = blockIdx.y;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
    cplx acc = cplx(0.0, 0.0);
    if (is_cont[n]) for (int f = 0; f < F; ++f) {
      const int p = site[n * F + f]; const T g = G[((long)chain * F + f) * sM + p + (long)p * N];
      const T z = coef[(n * F + f) * 2] * (one_<T>() - g) + coef[(n * F + f) * 2 + 1];
      acc = acc - cplx(real_(z), imag_(z)) * (double)n_sun;
    }
    forces[((long)chain * Ltrot + (nt - 1)) * M + n] = acc;
  }
}
// Scheme "Langevin" (Prog/Langevin_HMC_mod.F90:362-390): per chain the adaptive step Delta_t_running from the largest |Re force| (fermionic and
// Forces_0 = phi of the Gaussian action), then for n (outer), nt (inner):  phi -= (phi + Re(Phase F)/Re(Phase)) dt - sqrt(2 dt) rang()
static __global__ void k_langevin_update(double* __restrict__ fc, const cplx* __restrict__ forces, const unsigned char* __restrict__ is_cont, int M, int Ltrot,
                                         uint64_t* __restrict__ rng, const cplx* __restrict__ phase, double delta_t, double max_force, double* __restrict__ dt_out, int n_chains) {
  __shared__ double red[32];
  const int chain = blockIdx.x; if (chain >= n_chains) return;
  double* f = fc + (long)chain * Ltrot * M; const cplx* F = forces + (long)chain * Ltrot * M;
  double xm = 0.0;
  for (long e = threadIdx.x; e < (long)Ltrot * M; e += blockDim.x) { const int n = (int)(e % M); xm = fmax(xm, fabs(F[e].x)); if (is_cont[n]) xm = fmax(xm, fabs(f[e])); }
  for (int o = 16; o > 0; o >>= 1) xm = fmax(xm, __shfl_xor_sync(0xffffffffu, xm, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = xm;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) xm = fmax(xm, red[w]);
    double dt = delta_t; if (xm > max_force) dt = max_force * delta_t / xm;
    dt_out[chain] = dt;
    Xoshiro r; r.s0 = rng[chain * 4]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3];
    const cplx ph = phase[chain]; const double sq = sqrt(2.0 * dt);
    for (int n = 0; n < M; ++n) if (is_cont[n]) for (int nt = 0; nt < Ltrot; ++nt) {
      const long e = (long)nt * M + n;
      const double f0 = f[e]; const cplx pf = ph * F[e]; const double fr = pf.x / ph.x;
      const double ranmod = sqrt(-2.0 * log(r.ranf())); const double theta = 6.283185307179586476925286766559 * r.ranf();      // rang_wrap, random_wrap_mod.F90:144-157
      f[e] = f0 - (f0 + fr) * dt + sq * (ranmod * cos(theta));
    }
    rng[chain * 4] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
  }
}
// Scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571): momenta p(n, nt) = rang_wrap() drawn nt outer / n inner (:429-434), kinetic energy
static __global__ void k_hmc_momenta(double* __restrict__ p, uint64_t* __restrict__ rng, int M, int Ltrot, double* __restrict__ ekin, int n_chains) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x; if (chain >= n_chains) return;
  Xoshiro r; r.s0 = rng[chain * 4]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3];
  double* pc = p + (long)chain * Ltrot * M; double e = 0.0;
  for (long q = 0; q < (long)Ltrot * M; ++q) { const double ranmod = sqrt(-2.0 * log(r.ranf())); const double theta = 6.283185307179586476925286766559 * r.ranf();
    const double x = ranmod * cos(theta); pc[q] = x; e += 0.5 * x * x; }
  ekin[chain] = e;
  rng[chain * 4] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
}
// p -= x dt (Forces_0 + Re(Phase F)/Re(Phase)), Forces_0 = phi (Gaussian action)      (:436-440, :476-484)
static __global__ void k_hmc_kick(double* __restrict__ p, const double* __restrict__ fc, const cplx* __restrict__ forces, const cplx* __restrict__ phase, double xdt, long per_chain, int n_chains) {
  const int chain = blockIdx.y; const cplx ph = phase[chain];
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < per_chain; e += (long)gridDim.x * blockDim.x) {
    const long q = (long)chain * per_chain + e; const cplx pf = ph * forces[q];
    p[q] -= xdt * (fc[q] + pf.x / ph.x);
  }
}
static __global__ void k_hmc_drift(double* __restrict__ fc, const double* __restrict__ p, double dt, long n) {      // nsigma%f += dt p (:447-449)
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) fc[e] += dt * p[e];
}
// Compute_Ratio_Global (Prog/Global_mod.F90:651-760) with the Gaussian Get_Delta_S0_global, Weight, Metropolis test, restore on rejection (:516-563).
// ga[n][f] = N_SUN g alpha (complex).  One CTA per chain.
static __global__ void k_hmc_decide(double* __restrict__ fc, const double* __restrict__ fc_old, const double* __restrict__ p, const double* __restrict__ ekin_old,
                                    const double* __restrict__ ld_old, const cplx* __restrict__ pd_old, const double* __restrict__ ld_new, const cplx* __restrict__ pd_new,
                                    const cplx* __restrict__ ga, const cplx* __restrict__ phase_old, int F, int n_sun, int M, int Ltrot, uint64_t* __restrict__ rng,
                                    double* __restrict__ weight_out, unsigned char* __restrict__ acc_out) {
  __shared__ double red[4][8]; __shared__ int s_acc;
  const int chain = blockIdx.x, tid = threadIdx.x; const long per = (long)Ltrot * M;
  double* f = fc + chain * per; const double* fo = fc_old + chain * per; const double* pc = p + chain * per;
  double ek = 0.0, ds = 0.0, zr = 0.0, zi = 0.0;      // E_kin_new, sum(new^2 - old^2), sum_{n,nt,f} dphi N_SUN g alpha
  for (long e = tid; e < per; e += blockDim.x) {
    const int n = (int)(e % M); const double a = f[e], b = fo[e], d = a - b;
    ek += 0.5 * pc[e] * pc[e]; ds += a * a - b * b;
    for (int ff = 0; ff < F; ++ff) { const cplx g = ga[n * F + ff]; zr += d * g.x; zi += d * g.y; }
  }
  ek = warp_sum(ek); ds = warp_sum(ds); zr = warp_sum(zr); zi = warp_sum(zi);
  if ((tid & 31) == 0) { red[0][tid >> 5] = ek; red[1][tid >> 5] = ds; red[2][tid >> 5] = zr; red[3][tid >> 5] = zi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { ek += red[0][w]; ds += red[1][w]; zr += red[2][w]; zi += red[3][w]; }
    double r2 = 0.0; cplx r1 = cplx(1.0, 0.0);
    for (int ff = 0; ff < F; ++ff) {
      r2 += (double)n_sun * (ld_new[chain * F + ff] - ld_old[chain * F + ff]);
      const cplx q = pd_new[chain * F + ff] * conj_(pd_old[chain * F + ff]);      // unit-modulus phases: ratio = new * conj(old)
      cplx qn = q; for (int k = 1; k < n_sun; ++k) qn = qn * q;
      r1 = r1 * qn;
    }
    const double em = exp(zr); r1 = r1 * cplx(em * cos(zi), em * sin(zi));
    r2 += -0.5 * ds + (-ek + ekin_old[chain]);
    const cplx rt = r1 * exp(r2); const cplx ph = phase_old[chain]; const cplx pr = ph * rt;
    const double weight = fabs(pr.x / ph.x);
    Xoshiro r; r.s0 = rng[chain * 4]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3];
    const int acc = (weight > r.ranf()) ? 1 : 0;
    rng[chain * 4] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
    s_acc = acc; if (weight_out) weight_out[chain] = weight; if (acc_out) acc_out[chain] = (unsigned char)acc;
  }
  __syncthreads();
  if (!s_acc) for (long e = tid; e < per; e += blockDim.x) f[e] = fo[e];
}
// Compute_Fermion_Det (Prog/Global_mod.F90:792-1000), finite temperature: TP = U + V diag(D) with the scale separation of the STAB3 branch
// (columns with D > 1 are divided by D and log D is put aside, :912-918); extra[b] = sum_{D_J > 1} log D_J
template <typename T>
__global__ void k_fdet_build(T* __restrict__ TP, const T* __restrict__ U, const T* __restrict__ V, const double* __restrict__ D, long sM, int n, double* __restrict__ extra) {
  const int b = blockIdx.y; TP += (long)b * sM; U += (long)b * sM; V += (long)b * sM; D += (long)b * n;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e / n); const double d = D[j];
    TP[e] = (d <= 1.0) ? U[e] + V[e] * d : U[e] * (1.0 / d) + V[e];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { double s = 0.0; for (int j = 0; j < n; ++j) if (D[j] > 1.0) s += log(D[j]); extra[b] = s; }
}
// det U / |det U| of a matrix from the outputs of its pivoted QR (det R > 0 after the D scaling up to the unit-modulus diagonal)
static __global__ void k_det_phase(const QrOut* __restrict__ q, cplx* __restrict__ det, int nm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= nm) return;
  const cplx p = (q[b].detq * q[b].diag_phase) * q[b].perm_sign; const double a = abs_(p); det[b] = cplx(p.x / a, p.y / a);
}
// log|det| = sum log Dq + extra (+ sum of the log D of the given states, projector);  phase = detq * diag_phase * perm_sign [* conj(det U)]
static __global__ void k_fdet_finish(const double* __restrict__ Dq, int n, const QrOut* __restrict__ q, const double* __restrict__ extra, const cplx* __restrict__ detU,
                                     double* __restrict__ logdet, cplx* __restrict__ phase, int nm) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= nm) return;
  double s = extra ? extra[b] : 0.0;
  for (int i = 0; i < n; ++i) s += log(Dq[(long)b * n + i]);
  cplx p = (q[b].detq * q[b].diag_phase) * q[b].perm_sign;
  if (detU) p = p * conj_(detU[b]);
  const double a = abs_(p);
  logdet[b] = s; phase[b] = cplx(p.x / a, p.y / a);
}
static __global__ void k_fdet_addlog(const double* __restrict__ D, int ldD, int np, double* __restrict__ acc, int nm) {      // acc[b] += sum_{n < np} log D(n)
  const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= nm) return;
  double s = 0.0; for (int i = 0; i < np; ++i) s += log(D[(long)b * ldD + i]);
  acc[b] += s;
}
// G0T = -(1 - G)  (tau_m_mod.F90:96-104)
template <typename T>
__global__ void k_g0t_init(T* __restrict__ G0T, const T* __restrict__ G, long sM, int n) {
  const int b = blockIdx.y; G0T += (long)b * sM; G += (long)b * sM;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)n), i = (int)((unsigned)e - (unsigned)j * (unsigned)n);
    G0T[e] = G[e] - ((i == j) ? one_<T>() : zero_<T>());
  }
}

// reset_UDV_state with a trial wave function (udv_state_mod.F90:320-345): U(:, 1:N_part) = P of the matrix's flavor
template <typename T>
__global__ void k_set_wf(T* __restrict__ U, long sM, const T* __restrict__ wf, int F, int N, int NP) {
  const int b = blockIdx.y, f = b % F; U += (long)b * sM; wf += (long)f * N * NP;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)N * NP; e += (long)gridDim.x * blockDim.x) U[e] = wf[e];
}
// dst = alpha * src + beta * 1
template <typename T>
__global__ void k_axpb_identity(T* __restrict__ dst, const T* __restrict__ src, long sM, int n, double alpha, double beta) {
  const int b = blockIdx.y; dst += (long)b * sM; src += (long)b * sM;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)((unsigned)e / (unsigned)n), i = (int)((unsigned)e - (unsigned)j * (unsigned)n);
    dst[e] = alpha * src[e] + ((i == j) ? make_<T>(beta, 0.0) : zero_<T>());
  }
}

template <typename T>
struct Engine : EngineBase {
  alf_b200_handle* h;
  int C, F, N, L, M, S, NM; long n2; cudaStream_t st;
  std::vector<int> stab_nt;
  T *G = nullptr, *G2 = nullptr;
  UdvDev<T> udvl, udvr; std::vector<UdvDev<T>> udvst;
  LaWork<T> w;
  cplx* d_z = nullptr; double* d_angle = nullptr; double* d_angle_tab = nullptr; double* d_cmp = nullptr;
  VopDev<T>* d_vops = nullptr; T* d_place_tab = nullptr; int* d_place_pk = nullptr; unsigned char* d_is_cont = nullptr; ModelDev md; FieldTabDev ft;
  ModelFixDev mf; bool fix_any = false; int fix_max_blob = 0, fix_nmax = 8, fix_nbuf = 1, fix_nt = 256, fix_grid = 148; size_t fix_smem = 0;      // resident-descriptor form of the bond-operator lists (k_apply_ops_fixed)
  std::vector<void*> owned;     // device allocations of the op lists
  bool dense_t = false; T* d_dense[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // fwd, inv, c, half, halfinv (N*N*F each)
  int KD = 16; size_t upd_smem = 0; int ops_lk = 0; size_t ops_smem = 0;   // ops_lk = log2 of the largest small-operator dimension
  bool fast_upd = false; int KDf = 0, ldxf = 0, iptf = 1; size_t fast_smem = 0;   // k_wrapgr_fast (groups of diagonal / diagonalisable vertices)
  // The vertices of a slice are visited in GROUPS of consecutive vertices with pairwise disjoint supports:
  //   kind 1: diagonal single-site vertices (k = 1)                        -> windowed fast kernel
  //   kind 2: k = 2 vertices, rotated as a whole into their eigenbasis      -> fast kernel in pair mode between two op-list rotations
  //   kind 0: anything else (k > 2, repeated sites)                         -> per-visit kernel k_wrapgr
  struct VGroup { int n0 = 0, cnt = 0, kind = 0; OpListDev rot[4][ALF_FMAX]; };   // rot: in-left U^H, in-right U, out-left U, out-right U^H
  std::vector<VGroup> groups;
  S0TabDev s0dev = {0, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
  GmtDev gmtdev = {0, 0, nullptr, nullptr, {0, 0, nullptr, nullptr, nullptr, nullptr, nullptr}};
  int stage_g = 0;                 // k_wrapgr keeps G in shared memory
  int seq_lo = 0, seq_hi = 0;      // sequential visits: fields seq_lo .. seq_hi - 1 (0-based)
  // tau_m work
  T *GT0 = nullptr, *G0T = nullptr, *G00 = nullptr, *GTT = nullptr, *TMPG = nullptr; UdvDev<T> udvr2;
  // projective algorithm
  bool proj = false; int NP = 0, thtrot = 0; T *d_wfl = nullptr, *d_wfr = nullptr; LaWork<T> wp;

  template <typename X> X* dalloc(size_t n) { X* p = nullptr; CK(cudaMalloc(&p, sizeof(X) * (n ? n : 1))); owned.push_back(p); return p; }
  template <typename X> X* dupload(const std::vector<X>& v) { X* p = dalloc<X>(v.size()); if (!v.empty()) CK(cudaMemcpy(p, v.data(), sizeof(X) * v.size(), cudaMemcpyHostToDevice)); return p; }
  UdvDev<T> alloc_udv() { UdvDev<T> u; u.U = dalloc<T>(n2 * NM); u.V = dalloc<T>(n2 * NM); u.D = dalloc<double>((size_t)N * NM); u.det = dalloc<cplx>(NM); return u; }

  OpListDev upload_list(const ListBuild& lb) {
    OpListDev d; d.n_ops = (int)lb.k.size(); d.nvar = lb.nvar; d.mat_nt_stride = lb.nt_stride;
    std::vector<int> ls = lb.levels(N, OPS_CH); d.n_levels = (int)ls.size() - 1;
    for (int kk : lb.k) { while ((1 << ops_lk) < kk) ++ops_lk; }
    d.level_start = dupload(ls); d.k = dupload(lb.k); d.P = dupload(lb.P); d.fidx = dupload(lb.fidx);
    std::vector<T> m(lb.mat.size()); for (size_t i = 0; i < m.size(); ++i) m[i] = to_T<T>(lb.mat[i]);
    d.mat = dupload(m);
    std::vector<unsigned char> uni(std::max(d.n_levels, 1), 0);
    if (lb.nvar == 1) for (int c = 0; c < d.n_levels; ++c) {       // chunks whose operators all carry the same 2 x 2 matrix
      bool u = true;
      for (int o = ls[c]; o < ls[c + 1] && u; ++o) {
        if (lb.k[o] != 2) u = false;
        for (int e = 0; e < ALF_KMAX * ALF_KMAX && u; ++e) if (lb.mat[(size_t)o * ALF_KMAX * ALF_KMAX + e] != lb.mat[(size_t)ls[c] * ALF_KMAX * ALF_KMAX + e]) u = false;
      }
      uni[c] = u ? 1 : 0;
    }
    d.uniform = dupload(uni); d.cont = dupload(lb.cont);
    return d;
  }

  // Resident-descriptor form of a fixed list of k = 2 bond operators (k_apply_ops_fixed): families = maximal runs of consecutive operators with
  // pairwise disjoint supports, one 32-bit word per operator.
  FixListDev upload_fixed(const ListBuild& lb) {
    FixListDev d; std::memset(&d, 0, sizeof(d));
    if (lb.nvar != 1 || lb.k.empty() || N * OPS_PW > 65535 + OPS_PW || getenv("ALF_B200_NO_FIXED_OPS")) return d;
    for (int kk : lb.k) if (kk != 2) return d;
    const int n = (int)lb.k.size();
    std::vector<int> fs = lb.levels(N, 1 << 30);
    const int nfam = (int)fs.size() - 1;
    std::vector<unsigned> offs(n); std::vector<T> m((size_t)n * 4); std::vector<unsigned char> uni(nfam, 1);
    auto P0 = [&](int o) { return lb.P[(size_t)o * ALF_KMAX]; };
    auto P1 = [&](int o) { return lb.P[(size_t)o * ALF_KMAX + 1]; };
    auto A = [&](int o, int a, int b) { return lb.mat[(size_t)o * ALF_KMAX * ALF_KMAX + a + (size_t)b * ALF_KMAX]; };
    for (int o = 0; o < n; ++o) {
      offs[o] = (unsigned)(P0(o) * OPS_PW) | ((unsigned)(P1(o) * OPS_PW) << 16);
      m[4 * o] = to_T<T>(A(o, 0, 0)); m[4 * o + 1] = to_T<T>(A(o, 1, 0)); m[4 * o + 2] = to_T<T>(A(o, 0, 1)); m[4 * o + 3] = to_T<T>(A(o, 1, 1));
    }
    for (int c = 0; c < nfam; ++c) for (int o = fs[c]; o < fs[c + 1]; ++o) for (int e = 0; e < 4; ++e)
      if (A(o, e & 1, e >> 1) != A(fs[c], e & 1, e >> 1)) uni[c] = 0;
    // ---- ring groups (alf_ops.cuh): consecutive families over two perfect matchings A, B of the same sites whose union consists of rings of one length <= 32
    std::vector<FixGroupDev> grp; std::vector<FixFamDev> rfam(nfam, FixFamDev{0, 0, 0, 0}); std::vector<unsigned> ring; std::vector<T> rmat;
    const bool rings_on = !getenv("ALF_B200_NO_RING_OPS");
    auto partner_of = [&](int c, std::vector<int>& part, std::vector<int>& opi) {      // part[site] = partner in family c (or -1), opi[site] = its operator
      part.assign(N, -1); opi.assign(N, -1);
      for (int o = fs[c]; o < fs[c + 1]; ++o) { part[P0(o)] = P1(o); part[P1(o)] = P0(o); opi[P0(o)] = o; opi[P1(o)] = o; }
    };
    int c = 0;
    while (c < nfam) {
      FixGroupDev g = {0, c, 1, 0, 0, 0, 0, 0};
      std::vector<int> pa, oa, pb, ob, pc, oc;
      bool ok = rings_on && c + 1 < nfam;
      if (ok) {
        partner_of(c, pa, oa); partner_of(c + 1, pb, ob);
        for (int i = 0; i < N && ok; ++i) if ((pa[i] < 0) != (pb[i] < 0)) ok = false;                 // same support
        if (ok && pa == pb) ok = false;
      }
      std::vector<std::vector<int>> rings;
      if (ok) {
        std::vector<char> seen(N, 0);
        for (int s0 = 0; s0 < N && ok; ++s0) {
          if (pa[s0] < 0 || seen[s0]) continue;
          std::vector<int> r; int cur = s0;
          while (true) {
            r.push_back(cur); seen[cur] = 1; const int nx = pa[cur]; r.push_back(nx); seen[nx] = 1;
            cur = pb[nx]; if (cur == s0) break;
            if (seen[cur] || (int)r.size() > 32) { ok = false; break; }
          }
          if (ok && (int)r.size() > 32) ok = false;
          if (ok && !rings.empty() && r.size() != rings[0].size()) ok = false;
          if (ok) rings.push_back(r);
        }
      }
      if (ok) {
        const int rn = (int)rings[0].size(), nmax = rn <= 8 ? 8 : rn <= 16 ? 16 : 32;
        if (sizeof(T) > 8 && nmax > 16 && !getenv("ALF_B200_CPLX_RING32")) ok = false;       // complex: at most 16 ring values in registers by default
        if (ok) {
          // which of the following families reuse the two matchings
          int c2 = c;
          std::vector<int> types;
          while (c2 < nfam) { partner_of(c2, pc, oc); if (pc == pa) types.push_back(0); else if (pc == pb) types.push_back(1); else break; ++c2; }
          g.kind = 1; g.nfam = c2 - c; g.n = rn; g.nrings = (int)rings.size(); g.nmax = nmax; g.ring_off = (int)ring.size(); g.cover = (rn * g.nrings == N) ? 1 : 0;
          for (auto& r : rings) for (int i = 0; i < OPS_RSTR; ++i) ring.push_back((unsigned)(i < rn ? r[i] * OPS_PW + (r[i] & 31) : 0));
          for (int fi = 0; fi < g.nfam; ++fi) {
            const int cf = c + fi; partner_of(cf, pc, oc);
            std::vector<T> mm((size_t)g.nrings * (OPS_RSTR / 2) * 4, zero_<T>());
            bool u = true; cd first[4] = {cd(0, 0), cd(0, 0), cd(0, 0), cd(0, 0)}; bool have = false;
            for (int rg = 0; rg < g.nrings; ++rg) for (int i = 0; i < rn / 2; ++i) {
              const int sa = types[fi] == 0 ? rings[rg][2 * i] : rings[rg][2 * i + 1], sb = types[fi] == 0 ? rings[rg][2 * i + 1] : rings[rg][(2 * i + 2) % rn];
              const int o = oc[sa];
              cd e4[4];                                              // a00, a10, a01, a11 in the orientation (sa, sb)
              if (P0(o) == sa && P1(o) == sb) { e4[0] = A(o, 0, 0); e4[1] = A(o, 1, 0); e4[2] = A(o, 0, 1); e4[3] = A(o, 1, 1); }
              else { e4[0] = A(o, 1, 1); e4[1] = A(o, 0, 1); e4[2] = A(o, 1, 0); e4[3] = A(o, 0, 0); }
              for (int e = 0; e < 4; ++e) { mm[((size_t)rg * (OPS_RSTR / 2) + i) * 4 + e] = to_T<T>(e4[e]); if (have && e4[e] != first[e]) u = false; }
              if (!have) { for (int e = 0; e < 4; ++e) first[e] = e4[e]; have = true; }
            }
            rfam[cf].type = types[fi]; rfam[cf].uniform = u ? 1 : 0; rfam[cf].mat_off = (int)rmat.size();
            if (u) for (int e = 0; e < 4; ++e) rmat.push_back(to_T<T>(first[e])); else rmat.insert(rmat.end(), mm.begin(), mm.end());
          }
          c = c2;
        }
      }
      if (!ok) { g.kind = 0; g.nfam = 1; c += 1; }
      grp.push_back(g);
    }
    // ---- one blob per list (copied to shared memory by every CTA)
    std::vector<unsigned char> blob;
    auto put = [&](const void* src, size_t bytes) { while (blob.size() % 16) blob.push_back(0); const int off = (int)blob.size(); const unsigned char* q = (const unsigned char*)src; blob.insert(blob.end(), q, q + bytes); return off; };
    bool any_plain = false; for (auto& gg : grp) if (gg.kind == 0) any_plain = true;
    std::vector<T> m4((size_t)nfam * 4); for (int c2 = 0; c2 < nfam; ++c2) for (int e = 0; e < 4; ++e) m4[(size_t)c2 * 4 + e] = m[(size_t)fs[c2] * 4 + e];
    if (rmat.empty()) rmat.push_back(zero_<T>());
    if (ring.empty()) ring.resize(8, 0);
    d.o_grp = put(grp.data(), grp.size() * sizeof(FixGroupDev)); d.o_rfam = put(rfam.data(), rfam.size() * sizeof(FixFamDev));
    d.o_rmat = put(rmat.data(), rmat.size() * sizeof(T)); d.o_ring = put(ring.data(), ring.size() * sizeof(unsigned));
    d.o_offs = put(offs.data(), any_plain ? offs.size() * sizeof(unsigned) : 4); d.o_fs = put(fs.data(), fs.size() * sizeof(int));
    d.o_uni = put(uni.data(), uni.size()); d.o_m4 = put(m4.data(), m4.size() * sizeof(T));
    while (blob.size() % 16) blob.push_back(0);
    d.n_fam = nfam; d.n_ops = n; d.n_grp = (int)grp.size(); d.blob_bytes = (int)blob.size(); d.blob = dupload(blob); d.mat = dupload(m);
    fix_any = true; fix_max_blob = std::max(fix_max_blob, (int)blob.size()); for (auto& gg : grp) if (gg.kind == 1) fix_nmax = std::max(fix_nmax, gg.nmax);
    return d;
  }
  // vertex list usable as a row scaling: only k = 1 (diagonal) factors, every site at most once
  unsigned char diag_list_ok(const ListBuild& lb) const {
    std::vector<char> seen(N, 0);
    for (size_t o = 0; o < lb.k.size(); ++o) { if (lb.k[o] != 1) return 0; const int p = lb.P[o * ALF_KMAX]; if (seen[p]) return 0; seen[p] = 1; }
    return 1;
  }
  VDiagDev upload_vdiag(const ListBuild& lb) {             // diagonal vertex list by site (k_apply_ops_fixed)
    const int nsl = lb.nt_stride > 0 ? L : 1;
    std::vector<T> tab((size_t)nsl * N * ALF_NVAR, one_<T>()); std::vector<int> fi(N, -1); std::vector<unsigned char> ct(N, 0);
    if (diag_list_ok(lb) && lb.nvar == ALF_NVAR) for (int sl = 0; sl < nsl; ++sl) for (size_t o = 0; o < lb.k.size(); ++o) {
      const int p = lb.P[o * ALF_KMAX]; fi[p] = lb.fidx[o]; ct[p] = lb.cont[o];
      for (int var = 0; var < ALF_NVAR; ++var) tab[((size_t)sl * N + p) * ALF_NVAR + var] = to_T<T>(lb.mat[(size_t)sl * lb.nt_stride + (o * ALF_NVAR + var) * ALF_KMAX * ALF_KMAX]);
    }
    VDiagDev v; v.tab = dupload(tab); v.fidx = dupload(fi); v.cont = dupload(ct); v.tab_nt_stride = nsl > 1 ? (long)N * ALF_NVAR : 0;
    return v;
  }
  bool fixed_mode_ok(int mode) const {
    if (!fix_any) return false;
    int t = -1, v = -1;
    switch (mode) {
      case MODE_WRAPUR: t = L_TL_FWD; v = L_VL_N; break;
      case MODE_WRAPUL: t = L_TL_C; v = L_VL_C; break;
      case MODE_TL_FWD: t = L_TL_FWD; break;
      case MODE_TL_INV: t = L_TL_INV; break;
      case MODE_TL_C: t = L_TL_C; break;
      case MODE_TL_HALF: t = L_TL_HALF; break;
      case MODE_TR_FWD: t = L_TR_FWD; break;
      case MODE_TR_INV: t = L_TR_INV; break;
      case MODE_TR_HALFINV: t = L_TR_HALFINV; break;
      case MODE_PROPRM1: t = L_TR_INV; v = L_VR_INV; break;
      default: return false;
    }
    for (int f = 0; f < F; ++f) { if (mf.fix[t][f].n_fam <= 0) return false; if (v >= 0 && !mf.diag_ok[v][f]) return false; }
    return true;
  }

  Engine(alf_b200_handle* hh) : h(hh) {
    C = h->n_chains; F = h->n_fl; N = h->ndim; L = h->ltrot; M = h->n_opv; NM = C * F; n2 = (long)N * N; st = h->stream;
    if (F > ALF_FMAX) throw CudaError("more than ALF_FMAX flavors");
    if ((long)C * F > 65535) throw CudaError("n_chains * N_FL > 65535: the batched kernels index the matrices of a handle with gridDim.y; use several handles");
    if (N > 576) throw CudaError("Ndim > 576 is not supported in this build (QR / TRSM kernels hold <= 18 rows per lane; TAU_M needs 2 Ndim <= 576)");
    S = (L % h->nwrap == 0) ? L / h->nwrap : L / h->nwrap + 1;                 // main.F90:446-457
    std::memset(&mf, 0, sizeof(mf));
    stab_nt.assign(S + 1, 0); for (int n = 1; n < S; ++n) stab_nt[n] = h->nwrap * n; stab_nt[S] = L;
    build_model();
    G = dalloc<T>(n2 * NM); G2 = dalloc<T>(n2 * NM);
    udvl = alloc_udv(); udvr = alloc_udv(); udvst.resize(S); for (int s = 0; s < S; ++s) udvst[s] = alloc_udv();
    w.alloc(N, NM, st);
    proj = h->projector; NP = proj ? h->n_part : N; thtrot = h->thtrot;
    if (proj) {
      if (NP < 1 || NP > N) throw CudaError("projector: illegal number of particles (0 < N_part <= Ndim)");
      wp.alloc(NP, NM, st);
      std::vector<T> a((size_t)F * N * NP), b2((size_t)F * N * NP);
      for (int f = 0; f < F; ++f) for (long i = 0; i < (long)N * NP; ++i) { a[(size_t)f * N * NP + i] = to_T<T>(h->wf_l[f][i]); b2[(size_t)f * N * NP + i] = to_T<T>(h->wf_r[f][i]); }
      d_wfl = dupload(a); d_wfr = dupload(b2);
    }
    d_z = dalloc<cplx>(NM); d_angle = dalloc<double>(NM); d_cmp = dalloc<double>((size_t)NM * 3 * CMP_SPLIT);
    // update kernel configuration: as many delayed columns as fit in ~200 KB of shared memory
    size_t per_kd = (size_t)F * 2 * (N + 2) * sizeof(T), fixed = (size_t)F * 3 * N * sizeof(T) + 256;
    {   // small lattices: G itself in shared memory next to at least 8 delayed columns (per-visit kernel only)
      const size_t gsz = sizeof(T) * (size_t)F * (N | 1) * N;
      if (fixed + gsz + 8 * per_kd <= 224 * 1024 && !getenv("ALF_B200_NO_STAGE_G")) { stage_g = 1; fixed += gsz; }
    }
    KD = (int)(((stage_g ? 224 : 200) * 1024 - fixed) / per_kd); if (KD > 32) KD = 32; if (KD < 4) throw CudaError("Ndim too large for the update kernel's shared-memory factors");
    KD = (KD / 4) * 4;
    upd_smem = per_kd * KD + fixed;
    CK(alf_raise_smem(k_wrapgr<T, 1>));
    CK(alf_raise_smem(k_wrapgr<T, 0>));
    if (h->s0_on) {
      s0dev.on = 1; s0dev.open_bc = h->s0_open_bc; s0dev.op_start = dupload(h->s0_op_start); s0dev.term_start = dupload(h->s0_term_start);
      s0dev.e_op = dupload(h->s0_e_op); s0dev.e_dt = dupload(h->s0_e_dt); s0dev.w = dupload(h->s0_w);
    }
    seq_lo = h->nt_seq_start - 1; seq_hi = (h->nt_seq_end < 0) ? M : h->nt_seq_end;
    if (seq_lo < 0 || seq_hi > M || seq_lo > seq_hi) throw CudaError("illegal Nt_sequential_start / Nt_sequential_end");
    if ((seq_lo > 0 || seq_hi < M) && h->n_global_tau <= 0) throw CudaError("a restricted sequential range needs N_Global_tau > 0 (Wrapgr_mod.F90:148-153)");
    if (h->gmt_on) {
      gmtdev.on = 1; gmtdev.n_sites = h->gmt_n_sites; gmtdev.move_start = dupload(h->gmt_move_start); gmtdev.move_fields = dupload(h->gmt_move_fields);
      gmtdev.terms.on = 1; gmtdev.terms.open_bc = h->gmt_open_bc; gmtdev.terms.op_start = dupload(h->gmt_op_start); gmtdev.terms.term_start = dupload(h->gmt_term_start);
      gmtdev.terms.e_op = dupload(h->gmt_e_op); gmtdev.terms.e_dt = dupload(h->gmt_e_dt); gmtdev.terms.w = dupload(h->gmt_w);
    }
    // vertex groups (see VGroup): greedy over n = 1 .. M
    {
      auto kind_of = [&](int n) {
        int kd = -1;
        for (int f = 0; f < F; ++f) {
          const HostOp& o = h->opv[n + (size_t)M * f]; int kf = 0;
          if (o.type == 1 || o.type == 2) {
            if (o.N == 1 && o.nnz == 1 && o.diag) kf = 1;
            else if (o.N == 2 && o.nnz >= 1 && o.nnz <= 2) kf = 2;
          }
          if (kd < 0) kd = kf; else if (kd != kf) kd = 0;
        }
        if (getenv("ALF_B200_GENERIC_UPDATE") || h->s0_on || h->propose_s0) kd = 0;      // S0 tables / Propose_S0: per-visit kernel (draws depend on the data)
        return kd;
      };
      std::vector<std::vector<char>> seen(F, std::vector<char>(N, 0));
      for (int n = 0; n < M; ++n) {
        const int kd = kind_of(n); bool clash = false;
        for (int f = 0; f < F; ++f) { const HostOp& o = h->opv[n + (size_t)M * f]; for (int a = 0; a < o.N; ++a) if (seen[f][o.P[a]]) clash = true; }
        if (groups.empty() || groups.back().kind != kd || (kd != 0 && clash) || n == seq_lo || n == seq_hi) {
          VGroup g; g.n0 = n; g.kind = kd; groups.push_back(g);
          for (int f = 0; f < F; ++f) std::fill(seen[f].begin(), seen[f].end(), 0);
        }
        groups.back().cnt++;
        for (int f = 0; f < F; ++f) { const HostOp& o = h->opv[n + (size_t)M * f]; for (int a = 0; a < o.N; ++a) seen[f][o.P[a]] = 1; }
      }
      int mv_max = 0;                               // pseudo-visits of the largest fast group (sizes the kernel's per-visit tables)
      for (auto& g : groups) if (g.kind) mv_max = std::max(mv_max, g.cnt * g.kind);
      fast_upd = mv_max > 0;
      if (fast_upd) {
        ldxf = N; while (ldxf % 16 != 4) ++ldxf;
        const size_t Mv = mv_max;
        const size_t fixedf = (size_t)(2 * F * N + 6 * F * Mv + F * ALF_WIN * (ALF_WIN + 1) + 2 * ALF_WIN * F * ALF_WIN + ALF_WIN * F) * sizeof(T) + (size_t)2 * Mv * 8 + (size_t)F * Mv * 4 + (size_t)3 * Mv + 64;
        const size_t perkd = (size_t)2 * F * ldxf * sizeof(T);
        const size_t avail = 227 * 1024 - 1024;
        if (fixedf + 4 * perkd > avail || F * N > 4 * 512) fast_upd = false;
        else {
          KDf = (int)((avail - fixedf) / perkd); KDf = (KDf / 4) * 4; if (KDf > 64) KDf = 64;
          if (const char* e = getenv("ALF_B200_KD")) { int v = atoi(e); if (v >= 4 && v <= KDf) KDf = (v / 4) * 4; }
          fast_smem = fixedf + perkd * KDf;
          iptf = (F * N + 511) / 512; if (iptf > 1) iptf = 4;
#define FAST_ATTR(IPT, PR) do { CK(alf_raise_smem(k_wrapgr_fast<T, 1, IPT, PR>)); \
                                CK(alf_raise_smem(k_wrapgr_fast<T, 0, IPT, PR>)); } while (0)
          if (iptf == 1) { FAST_ATTR(1, 0); FAST_ATTR(1, 1); } else { FAST_ATTR(4, 0); FAST_ATTR(4, 1); }
#undef FAST_ATTR
        }
      }
      if (!fast_upd) for (auto& g : groups) g.kind = 0;
      // basis rotations of the pair groups as op lists: left matrices act on rows P, "right" lists act on the transpose (alf_ops.cuh)
      for (auto& g : groups) if (g.kind == 2) {
        for (int f = 0; f < F; ++f) {
          ListBuild il, ir, ol, orr;
          for (int n = g.n0; n < g.n0 + g.cnt; ++n) {
            const HostOp& o = h->opv[n + (size_t)M * f];
            std::vector<cd> U(4), UH(4), UT(4), UC(4);
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) { const cd u = o.U[a + (size_t)b * 2]; U[a + b * 2] = u; UH[b + a * 2] = std::conj(u); UT[b + a * 2] = u; UC[a + b * 2] = std::conj(u); }
            il.add(2, o.P.data(), -1, {UH}); ir.add(2, o.P.data(), -1, {UT});      // G <- U^H G U
            ol.add(2, o.P.data(), -1, {U}); orr.add(2, o.P.data(), -1, {UC});      // G <- U G U^H
          }
          g.rot[0][f] = upload_list(il); g.rot[1][f] = upload_list(ir); g.rot[2][f] = upload_list(ol); g.rot[3][f] = upload_list(orr);
        }
      }
    }
    // op-list kernel: panel of 32 columns (rows) + double-buffered operator descriptors
    ops_smem = (((size_t)N * OPS_PW + 1) & ~(size_t)1) * sizeof(T) + 2 * OPS_CH * ((sizeof(T) << (2 * ops_lk)) + 4 * sizeof(int));
    if (ops_smem > 227 * 1024) throw CudaError("Ndim too large for the op-list kernel's shared-memory panel");
#define OPS_ATTR(LKV) do { CK(alf_raise_smem(k_apply_ops<T, 0, LKV>)); \
                            CK(alf_raise_smem(k_apply_ops<T, 1, LKV>)); } while (0)
    if (ops_lk == 0) OPS_ATTR(0); else if (ops_lk == 1) OPS_ATTR(1); else OPS_ATTR(2);
#undef OPS_ATTR
    if (fix_any) {      // persistent resident-descriptor kernel: two panel buffers if they fit, CTAs per SM from the occupancy calculator
      fix_nbuf = ops_fixed_smem2(sizeof(T), N, fix_max_blob, F, 2) <= 227 * 1024 ? 2 : 1;
      fix_smem = ops_fixed_smem2(sizeof(T), N, fix_max_blob, F, fix_nbuf);
      fix_nt = (N >= 192 && fix_nmax <= 16 && sizeof(T) == 8) ? OPSF_NT : 256;      // rings of 32 values (and complex ones) need more than 128 registers per thread
      if (fix_smem > 227 * 1024) fix_any = false;
    }
    if (fix_any) {
      int dev = 0, nsm = 148, occ = 1; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
#define OPSF_ATTR(NMX) do { CK(alf_raise_smem(k_apply_ops_fixed<T, 0, NMX>)); CK(alf_raise_smem(k_apply_ops_fixed<T, 1, NMX>)); \
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_apply_ops_fixed<T, 0, NMX>, fix_nt, fix_smem)); } while (0)
      if (fix_nmax <= 8) OPSF_ATTR(8); else if (fix_nmax <= 16) OPSF_ATTR(16); else OPSF_ATTR(32);
#undef OPSF_ATTR
      fix_grid = nsm * std::max(occ, 1);
    }
  }
  ~Engine() { for (void* p : owned) cudaFree(p); w.release(); if (proj) wp.release(); if (w2_ready) w2.release(); }
  // Optional (ALF_B200_L2_PERSIST=1): persisting-L2 access window over the batch of Green functions.  Measured on B200 with
  // 148 chains x 1 MB (slightly more than L2): the slice kernel got SLOWER (1.28 ms vs 1.05 ms), so it is off by default.
  double l2_persist_frac = -1.0; size_t l2_persist_bytes = 0;
  void l2_window(const void* base, size_t bytes) {
    if (l2_persist_frac < 0.0) {
      l2_persist_frac = 0.0;
      if (getenv("ALF_B200_L2_PERSIST")) {
        int dev = h->device, maxp = 0, maxw = 0;
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev); cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (maxp > 0 && maxw > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) == cudaSuccess) { l2_persist_bytes = (size_t)maxp; l2_persist_frac = 1.0; l2_maxw = (size_t)maxw; }
        cudaGetLastError();
      }
    }
    if (l2_persist_frac <= 0.0) return;
    cudaStreamAttrValue a; memset(&a, 0, sizeof(a));
    const size_t nb = std::min(bytes, l2_maxw);
    a.accessPolicyWindow.base_ptr = const_cast<void*>(base); a.accessPolicyWindow.num_bytes = nb;
    a.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)l2_persist_bytes / (double)nb);
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) { cudaGetLastError(); l2_persist_frac = 0.0; }
  }
  size_t l2_maxw = 0;
  void sync() override { CK(cudaStreamSynchronize(st)); }

  // ---------------------------------------------------------------- model tables (Hop_mod_init + Op_set tables)
  void build_model() {
    // hopping
    std::vector<HostExpT> ex((size_t)h->n_opt * F);
    int maxk = 0;
    for (int nc = 0; nc < h->n_opt; ++nc) for (int f = 0; f < F; ++f) { host_expopt(h->opt[nc + (size_t)h->n_opt * f], ex[nc + (size_t)h->n_opt * f]); maxk = std::max(maxk, h->opt[nc + (size_t)h->n_opt * f].N); }
    dense_t = maxk > ALF_KMAX;
    auto tr = [](const std::vector<cd>& A, int k) { std::vector<cd> B(A.size()); for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) B[i + (size_t)j * k] = A[j + (size_t)i * k]; return B; };
    auto one = [](const std::vector<cd>& m) { return std::vector<std::vector<cd>>(1, m); };
    for (int f = 0; f < F; ++f) {
      ListBuild fwd, inv, cc, half, rfwd, rinv, rhalfinv;
      if (!dense_t) {
        for (int nc = h->n_opt - 1; nc >= 0; --nc) { const HostExpT& e = ex[nc + (size_t)h->n_opt * f]; if (!e.active) continue;
          fwd.add(e.N, e.P.data(), -1, one(e.mat)); half.add(e.N, e.P.data(), -1, one(e.mat12)); rinv.add(e.N, e.P.data(), -1, one(tr(e.invmat, e.N))); rhalfinv.add(e.N, e.P.data(), -1, one(tr(e.invmat12, e.N))); }
        for (int nc = 0; nc < h->n_opt; ++nc) { const HostExpT& e = ex[nc + (size_t)h->n_opt * f]; if (!e.active) continue;
          inv.add(e.N, e.P.data(), -1, one(e.invmat)); cc.add(e.N, e.P.data(), -1, one(e.mat)); rfwd.add(e.N, e.P.data(), -1, one(tr(e.mat, e.N))); }
      }
      md.lists[L_TL_FWD][f] = upload_list(fwd); md.lists[L_TL_INV][f] = upload_list(inv); md.lists[L_TL_C][f] = upload_list(cc);
      md.lists[L_TL_HALF][f] = upload_list(half); md.lists[L_TR_FWD][f] = upload_list(rfwd); md.lists[L_TR_INV][f] = upload_list(rinv);
      md.lists[L_TR_HALFINV][f] = upload_list(rhalfinv);
      if (!dense_t) {
        mf.fix[L_TL_FWD][f] = upload_fixed(fwd); mf.fix[L_TL_INV][f] = upload_fixed(inv); mf.fix[L_TL_C][f] = upload_fixed(cc); mf.fix[L_TL_HALF][f] = upload_fixed(half);
        mf.fix[L_TR_FWD][f] = upload_fixed(rfwd); mf.fix[L_TR_INV][f] = upload_fixed(rinv); mf.fix[L_TR_HALFINV][f] = upload_fixed(rhalfinv);
      }
    }
    if (dense_t) {
      // products of the (few) dense factors in the orders of Hop_mod.F90:143-250
      for (int q = 0; q < 5; ++q) d_dense[q] = dalloc<T>(n2 * F);
      for (int f = 0; f < F; ++f) {
        auto eye = [&]() { std::vector<cd> I(n2, cd(0, 0)); for (int i = 0; i < N; ++i) I[i + (size_t)i * N] = 1; return I; };
        auto lmul = [&](const std::vector<cd>& A, std::vector<cd>& X) { std::vector<cd> Y(n2, cd(0, 0)); for (int j = 0; j < N; ++j) for (int k = 0; k < N; ++k) { cd x = X[k + (size_t)j * N]; if (x == cd(0, 0)) continue; for (int i = 0; i < N; ++i) Y[i + (size_t)j * N] += A[i + (size_t)k * N] * x; } X.swap(Y); };
        auto embed = [&](const HostExpT& e, const std::vector<cd>& m) { std::vector<cd> A = eye(); for (int a = 0; a < e.N; ++a) for (int b = 0; b < e.N; ++b) A[e.P[a] + (size_t)e.P[b] * N] = m[a + (size_t)b * e.N]; return A; };
        std::vector<cd> Ef = eye(), Ei = eye(), Ec = eye(), Eh = eye(), Ehi = eye();
        for (int nc = h->n_opt - 1; nc >= 0; --nc) { const HostExpT& e = ex[nc + (size_t)h->n_opt * f]; if (!e.active) continue; lmul(embed(e, e.mat), Ef); lmul(embed(e, e.mat12), Eh); }
        for (int nc = 0; nc < h->n_opt; ++nc) { const HostExpT& e = ex[nc + (size_t)h->n_opt * f]; if (!e.active) continue; lmul(embed(e, e.invmat), Ei); lmul(embed(e, e.mat), Ec); }
        // right products: In * inv_N ... inv_1  (mmthl_m1) equals Ei as a matrix; In * mat_1 ... mat_N equals Ef.  Half inverse:
        { std::vector<cd> X = eye(); for (int nc = 0; nc < h->n_opt; ++nc) { const HostExpT& e = ex[nc + (size_t)h->n_opt * f]; if (!e.active) continue; lmul(embed(e, e.invmat12), X); } Ehi = X; }
        std::vector<cd>* src[5] = {&Ef, &Ei, &Ec, &Eh, &Ehi};
        for (int q = 0; q < 5; ++q) { std::vector<T> t(n2); for (long i = 0; i < n2; ++i) t[i] = to_T<T>((*src[q])[i]); CK(cudaMemcpy(d_dense[q] + n2 * f, t.data(), sizeof(T) * n2, cudaMemcpyHostToDevice)); }
      }
    }
    // vertices.  With time-dependent couplings g_t (Operator_mod.F90:66: the reference then evaluates Op_exp on the fly, :583-604, 677-699, 768-791, 885-908)
    // every table below exists once per time slice: vops[sl][n][f], angle_tab[sl][n][f][var], the matrices of the three vertex lists [sl][op][var].
    const int nsl = h->has_gt ? L : 1;
    std::vector<VopDev<T>> vops((size_t)nsl * M * F);
    std::vector<double> angle_tab((size_t)nsl * M * F * ALF_NVAR, 0.0);
    for (int f = 0; f < F; ++f) {
      ListBuild vn0, vc0, vri0; std::vector<cd> vn_mat, vc_mat, vri_mat;
      for (int sl = 0; sl < nsl; ++sl) {
      ListBuild vn, vc, vri; vn.nvar = vc.nvar = vri.nvar = ALF_NVAR;
      auto gof = [&](const HostOp& o) { return o.g_t.empty() ? o.g : o.g_t[sl]; };
      std::vector<std::vector<std::vector<cd>>> mexp(M);   // [n][var] k x k
      for (int n = 0; n < M; ++n) {
        const HostOp& op = h->opv[n + (size_t)M * f]; const int k = op.N; const cd og = gof(op);
        if (k > ALF_KMAX) throw CudaError("interaction vertex with N > ALF_KMAX is not supported in this build");
        VopDev<T>& v = vops[((size_t)sl * M + n) * F + f];
        std::memset(&v, 0, sizeof(v));
        v.k = k; v.nnz = op.nnz; v.diag = op.diag; v.type = op.type;
        for (int a = 0; a < ALF_KMAX; ++a) v.gE[a] = to_T<T>(a < op.nnz ? og * op.E[a] : cd(0, 0));
        v.galpha = to_T<T>(og * op.alpha);
        for (int a = 0; a < k; ++a) v.P[a] = op.P[a];
        for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) v.U[a + b * ALF_KMAX] = to_T<T>(op.U[a + (size_t)b * k]);
        mexp[n].resize(ALF_NVAR);
        for (int var = 0; var < ALF_NVAR; ++var) {
          const int s = var - 2; const bool valid = (s != 0) && (std::abs(s) <= op.type);
          const double ph = valid ? phi_st(op.type, s) : 0.0;
          for (int a = 0; a < ALF_KMAX; ++a) v.E_exp[a][var] = to_T<T>((a < op.nnz && valid) ? std::exp(og * op.E[a] * ph) : cd(1, 0));
          host_op_exp(og * ph, op, mexp[n][var]);
          angle_tab[(((size_t)sl * M + n) * F + f) * ALF_NVAR + var] = (op.type == 3) ? (og * op.alpha).imag() : (valid ? (og * op.alpha * ph).imag() : 0.0);   // type 3: coefficient of phi
          for (int var2 = 0; var2 < ALF_NVAR; ++var2) {
            const int s2 = var2 - 2; const bool valid2 = (s2 != 0) && (std::abs(s2) <= op.type);
            const double dphi = (valid && valid2) ? (phi_st(op.type, s2) - ph) : 0.0;
            for (int a = 0; a < ALF_KMAX; ++a) v.delta[a][var][var2] = to_T<T>((a < op.nnz) ? std::exp(og * dphi * op.E[a]) - 1.0 : cd(0, 0));
            v.expalpha[var][var2] = to_T<T>(std::exp(og * dphi * op.alpha));
          }
        }
      }
      auto ct = [](const std::vector<cd>& A, int k) { std::vector<cd> B(A.size()); for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) B[i + (size_t)j * k] = std::conj(A[j + (size_t)i * k]); return B; };
      for (int n = 0; n < M; ++n) { const HostOp& op = h->opv[n + (size_t)M * f]; const cd og = gof(op); if (!h->has_gt && std::abs(op.g) < kEpsMachine) continue;   // quick return, Operator_mod.F90:576,668 (with g_t the factor of a slice with g_t = 0 is the identity)
        if (op.type == 3) {      // continuous field, k = 1: the kernels evaluate exp(c phi) with c = +g E (B), -g E (B^-1)
          std::vector<std::vector<cd>> c1(ALF_NVAR, std::vector<cd>(1, og * op.E[0])), c2(ALF_NVAR, std::vector<cd>(1, -og * op.E[0]));
          vn.add(1, op.P.data(), n, c1, true); vri.add(1, op.P.data(), n, c2, true); continue; }
        vn.add(op.N, op.P.data(), n, mexp[n]);
        std::vector<std::vector<cd>> mi(ALF_NVAR); for (int var = 0; var < ALF_NVAR; ++var) mi[var] = tr(mexp[n][ALF_NVAR - 1 - var], op.N);   // exp(-phi g O)^T
        vri.add(op.N, op.P.data(), n, mi); }
      for (int n = M - 1; n >= 0; --n) { const HostOp& op = h->opv[n + (size_t)M * f]; const cd og = gof(op); if (!h->has_gt && std::abs(op.g) < kEpsMachine) continue;
        if (op.type == 3) { std::vector<std::vector<cd>> c3(ALF_NVAR, std::vector<cd>(1, std::conj(og * op.E[0]))); vc.add(1, op.P.data(), n, c3, true); continue; }   // (B)^dagger
        std::vector<std::vector<cd>> mc(ALF_NVAR); for (int var = 0; var < ALF_NVAR; ++var) mc[var] = ct(mexp[n][var], op.N);
        vc.add(op.N, op.P.data(), n, mc); }
      if (sl == 0) { vn0 = vn; vc0 = vc; vri0 = vri; }
      vn_mat.insert(vn_mat.end(), vn.mat.begin(), vn.mat.end()); vc_mat.insert(vc_mat.end(), vc.mat.begin(), vc.mat.end()); vri_mat.insert(vri_mat.end(), vri.mat.begin(), vri.mat.end());
      }   // time slices
      ListBuild& vn = vn0; ListBuild& vc = vc0; ListBuild& vri = vri0;
      if (nsl > 1) { vn.nt_stride = (long)vn.mat.size(); vc.nt_stride = (long)vc.mat.size(); vri.nt_stride = (long)vri.mat.size(); vn.mat = vn_mat; vc.mat = vc_mat; vri.mat = vri_mat; }
      md.lists[L_VL_N][f] = upload_list(vn); md.lists[L_VL_C][f] = upload_list(vc); md.lists[L_VR_INV][f] = upload_list(vri);
      mf.diag_ok[L_VL_N][f] = diag_list_ok(vn); mf.diag_ok[L_VL_C][f] = diag_list_ok(vc); mf.diag_ok[L_VR_INV][f] = diag_list_ok(vri);
      mf.vd[L_VL_N][f] = upload_vdiag(vn); mf.vd[L_VL_C][f] = upload_vdiag(vc); mf.vd[L_VR_INV][f] = upload_vdiag(vri);
    }
    d_vops = dupload(vops); d_angle_tab = dupload(angle_tab); md.fields_c = h->d_fields_c;
    { std::vector<unsigned char> tc(M); for (int n = 0; n < M; ++n) tc[n] = h->opv[n].type == 3 ? 1 : 0; d_is_cont = dupload(tc); }
    {   // e^{+V_n(s)} and e^{-V_n(s)} per vertex and field value for Wrapgr_PlaceGR (gm_place_step)
      std::vector<T> pt((size_t)M * F * ALF_NVAR * 2 * ALF_KMAX * ALF_KMAX, zero_<T>());
      for (int f = 0; f < F; ++f) for (int n = 0; n < M; ++n) {
        const HostOp& op = h->opv[n + (size_t)M * f]; const int k = op.N;
        for (int var = 0; var < ALF_NVAR; ++var) {
          const int sv = var - 2; const bool valid = (sv != 0) && (std::abs(sv) <= op.type);
          const double ph = valid ? phi_st(op.type, sv) : 0.0;
          std::vector<cd> Ep, Em; host_op_exp(op.g * ph, op, Ep); host_op_exp(-op.g * ph, op, Em);
          T* dst = pt.data() + (((size_t)n * F + f) * ALF_NVAR + var) * 2 * ALF_KMAX * ALF_KMAX;
          for (int a = 0; a < k; ++a) for (int b2 = 0; b2 < k; ++b2) { dst[a + b2 * ALF_KMAX] = to_T<T>(Ep[a + (size_t)b2 * k]); dst[ALF_KMAX * ALF_KMAX + a + b2 * ALF_KMAX] = to_T<T>(Em[a + (size_t)b2 * k]); }
        }
      }
      d_place_tab = dupload(pt);
      std::vector<int> pk((size_t)M * F * 8, 0);
      for (int f = 0; f < F; ++f) for (int n = 0; n < M; ++n) { const HostOp& op = h->opv[n + (size_t)M * f];
        for (int a = 0; a < op.N && a < ALF_KMAX; ++a) pk[((size_t)n * F + f) * 8 + a] = op.P[a];
        pk[((size_t)n * F + f) * 8 + 4] = op.N; }
      d_place_pk = dupload(pk);
    }
    for (int t = 0; t < 3; ++t) for (int var = 0; var < ALF_NVAR; ++var) ft.gama[t][var] = gama_st(t, var - 2);
    const int fl[5][4] = {{0, -1, 1, 2}, {0, 1, 2, -2}, {0, 0, 0, 0}, {0, 2, -2, -1}, {0, -2, -1, 1}};   // Fields_mod.F90:289-301
    for (int a = 0; a < 5; ++a) for (int b = 0; b < 4; ++b) ft.flip[a][b] = fl[a][b];
  }

  // ---------------------------------------------------------------- op-list launches
  void apply_ops(T* Mx, int side, int mode, int nt_a, int nt_b, int nvec = -1, const ModelDev* mdo = nullptr, T* Mout = nullptr) {   // nvec: columns (side 0) / rows (side 1)
    if (nvec < 0) nvec = N;
    if (!Mout) Mout = Mx;                          // in place unless a result buffer is given (full matrices only)
    const ModelDev& md = mdo ? *mdo : this->md;
    dim3 grid((nvec + OPS_PW - 1) / OPS_PW, NM);
    if (!mdo && fixed_mode_ok(mode)) {             // bond-operator hopping list (+ diagonal vertices): resident-descriptor kernel
      const int npan = (nvec + OPS_PW - 1) / OPS_PW, total = npan * NM, gridp = std::min(total, fix_grid), bstride = (fix_max_blob + 15) & ~15;
#define OPSF_LAUNCH(SD, NMX) KL(KC_OPS, st, k_apply_ops_fixed<T, SD, NMX><<<gridp, fix_nt, fix_smem, st>>>(Mx, n2, N, nvec, md, mf, F, mode, nt_a, nt_b, h->d_fields, L, M, Mout, npan, total, fix_nbuf, bstride))
      if (side == 0) { if (fix_nmax <= 8) OPSF_LAUNCH(0, 8); else if (fix_nmax <= 16) OPSF_LAUNCH(0, 16); else OPSF_LAUNCH(0, 32); }
      else { if (fix_nmax <= 8) OPSF_LAUNCH(1, 8); else if (fix_nmax <= 16) OPSF_LAUNCH(1, 16); else OPSF_LAUNCH(1, 32); }
#undef OPSF_LAUNCH
      CKL(); return;
    }
#define OPS_LAUNCH(SD, LKV) KL(KC_OPS, st, k_apply_ops<T, SD, LKV><<<grid, OPS_NT, ops_smem, st>>>(Mx, n2, N, nvec, md, F, mode, nt_a, nt_b, h->d_fields, L, M, Mout))
    if (side == 0) { if (ops_lk == 0) OPS_LAUNCH(0, 0); else if (ops_lk == 1) OPS_LAUNCH(0, 1); else OPS_LAUNCH(0, 2); }
    else { if (ops_lk == 0) OPS_LAUNCH(1, 0); else if (ops_lk == 1) OPS_LAUNCH(1, 1); else OPS_LAUNCH(1, 2); }
#undef OPS_LAUNCH
    CKL();
  }
  // dense hopping: Mx <- E * Mx (left) or Mx * E (right), E per flavor (batch stride 0 inside a flavor is emulated per flavor)
  void dense_mult(T* Mx, int which, bool left) {
    for (int f = 0; f < F; ++f) {
      // matrices of flavor f are at b = c*F + f : stride F*n2 over chains
      if (left) gemm<T, 0, 0, 0>(st, N, N, N, d_dense[which] + n2 * f, N, 0, Mx + n2 * f, N, n2 * F, w.W[3] + n2 * f, N, n2 * F, C);
      else gemm<T, 0, 0, 0>(st, N, N, N, Mx + n2 * f, N, n2 * F, d_dense[which] + n2 * f, N, 0, w.W[3] + n2 * f, N, n2 * F, C);
    }
    CK(cudaMemcpyAsync(Mx, w.W[3], sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st));
  }
  // Hop_mod entry points on a batch of N x N matrices
  void mmthr(T* Mx) { if (dense_t) dense_mult(Mx, 0, true); else apply_ops(Mx, 0, MODE_TL_FWD, 0, 0); }
  void mmthr_m1(T* Mx) { if (dense_t) dense_mult(Mx, 1, true); else apply_ops(Mx, 0, MODE_TL_INV, 0, 0); }
  void mmthl(T* Mx) { if (dense_t) dense_mult(Mx, 0, false); else apply_ops(Mx, 1, MODE_TR_FWD, 0, 0); }
  void mmthl_m1(T* Mx) { if (dense_t) dense_mult(Mx, 1, false); else apply_ops(Mx, 1, MODE_TR_INV, 0, 0); }
  void mmthlc(T* Mx) { if (dense_t) dense_mult(Mx, 2, true); else apply_ops(Mx, 0, MODE_TL_C, 0, 0); }
  void hop_symm(T* Mx) { if (dense_t) { dense_mult(Mx, 3, true); dense_mult(Mx, 4, false); } else { apply_ops(Mx, 0, MODE_TL_HALF, 0, 0); apply_ops(Mx, 1, MODE_TR_HALFINV, 0, 0); } }

  void set_udv_identity(UdvDev<T>& u) {   // reset_UDV_state, udv_state_mod.F90:224-249
    dim3 eg(ew_blocks(n2), NM);
    KL(KC_EW, st, k_set_identity<T><<<eg, 256, 0, st>>>(u.U, N, n2, N, N)); KL(KC_EW, st, k_set_identity<T><<<eg, 256, 0, st>>>(u.V, N, n2, N, N));
    KL(KC_EW, st, k_fill_double<<<ew_blocks((long)N * NM), 256, 0, st>>>(u.D, (long)N * NM, 1.0));
    KL(KC_EW, st, k_fill_cplx<<<ew_blocks(NM), 256, 0, st>>>(u.det, NM, cplx(1.0, 0.0)));
  }
  void copy_udv(UdvDev<T>& dst, const UdvDev<T>& src) {   // assign_UDV_state, udv_state_mod.F90:300-330
    CK(cudaMemcpyAsync(dst.U, src.U, sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dst.V, src.V, sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dst.D, src.D, sizeof(double) * N * NM, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dst.det, src.det, sizeof(cplx) * NM, cudaMemcpyDeviceToDevice, st));
  }
  // reset_UDV_state: identity (finite temperature) or the trial wave function of the given side (projector)
  void reset_udv(UdvDev<T>& u, char side) {
    if (!proj) { set_udv_identity(u); return; }
    KL(KC_EW, st, k_set_wf<T><<<dim3(ew_blocks((long)N * NP), NM), 256, 0, st>>>(u.U, n2, (side == 'l' || side == 'L') ? d_wfl : d_wfr, F, N, NP));
    KL(KC_EW, st, k_fill_double<<<ew_blocks((long)N * NM), 256, 0, st>>>(u.D, (long)N * NM, 1.0));
    KL(KC_EW, st, k_fill_cplx<<<ew_blocks(NM), 256, 0, st>>>(u.det, NM, cplx(1.0, 0.0)));
  }
  void decompose(UdvDev<T>& u, char side) { if (proj) la_decompose_proj<T>(w, u, side, NP); else la_decompose<T>(w, u, side); }
  void udv_reset(int which, char side) override { reset_udv(which == 0 ? udvl : udvr, side); }

  // WRAPUR (Prog/wrapur_mod.F90:102-123) / WRAPUL (Prog/wrapul_mod.F90:108-129) on a batch
  void wrapur_on(UdvDev<T>& u, int ntau, int ntau1) {
    if (!dense_t) apply_ops(u.U, 0, MODE_WRAPUR, ntau + 1, ntau1, NP);
    else for (int nt = ntau + 1; nt <= ntau1; ++nt) { dense_mult(u.U, 0, true); apply_ops(u.U, 0, MODE_WRAPUR, nt, nt, NP); }
    decompose(u, 'r');
  }
  void wrapul_on(UdvDev<T>& u, int ntau1, int ntau) {
    if (!dense_t) apply_ops(u.U, 0, MODE_WRAPUL, ntau + 1, ntau1, NP);
    else for (int nt = ntau1; nt >= ntau + 1; --nt) { apply_ops(u.U, 0, MODE_WRAPUL, nt, nt, NP); dense_mult(u.U, 2, true); }
    decompose(u, 'l');
  }
  void wrapur(int ntau, int ntau1) override { wrapur_on(udvr, ntau, ntau1); }
  void wrapul(int ntau1, int ntau) override { wrapul_on(udvl, ntau1, ntau); }

  // CGR + Op_phase + Control_PrecisionG/P  (main.F90:742-753)
  void cgr_and_phase(int nvar, bool compare) {
    if (proj) la_cgrp<T>(w, wp, udvr, udvl, G2, d_z);      // cgr1_mod.F90:207-211
    else la_cgr<T>(w, nvar, h->stab, udvr, udvl, G2, d_z);
    if (compare) {
      KL(KC_EW, st, k_compare<T><<<dim3(NM, CMP_SPLIT), 256, 0, st>>>(G2, G, n2, n2, d_cmp));
      KL(KC_EW, st, k_ctl_accum<<<(C + 127) / 128, 128, 0, st>>>(d_cmp, F, h->d_ctl, 0, C));
    }
    std::swap(G, G2);
    l2_window(G, sizeof(T) * n2 * NM);
    double* ang = nullptr;
    if (h->is_complex) { KL(KC_EW, st, k_op_phase<<<NM, 256, 0, st>>>(h->d_fields, d_angle_tab, F, M, L, d_angle, h->d_fields_c, d_is_cont, h->has_gt ? (long)M * F * ALF_NVAR : 0)); ang = d_angle; }
    KL(KC_EW, st, k_phase_update<<<(C + 127) / 128, 128, 0, st>>>(d_z, ang, F, h->n_sun, h->d_phase, h->d_ctl, compare ? 1 : 0, C));
  }
  void cgr_call(int nvar) override { cgr_and_phase(nvar, false); }

  // WRAPGRUP / WRAPGRDO (Prog/Wrapgr_mod.F90:81-157, 160-243)
  void wrapgrup(int ntau) override {
    mmthr(G); mmthl_m1(G);
    launch_update(1, ntau + 1);
    if (h->n_global_tau > 0 && h->gmt_on) { gm_set_position(seq_hi); gm_device_moves(ntau + 1, M); }           // Wrapgr_mod.F90:148-153
  }
  void wrapgrdo(int ntau) override {
    if (h->n_global_tau > 0 && h->gmt_on) { gm_set_position(M); gm_device_moves(ntau, seq_hi); }               // :189-194
    launch_update(0, ntau);
    mmthl(G); mmthr_m1(G);
  }
  // Wrapgr_Random_update with ham%Global_move_tau evaluated on the device from the tables, then Wrapgr_PlaceGR(GR, m, place_to, ntau)
  void gm_device_moves(int ntau, int place_to) {
    gm_alloc();
    const size_t smem = gm_smem();
    CK(alf_raise_smem(k_random_update<T>));
    KL(KC_UPDATE, st, k_random_update<T><<<C, 512, smem, st>>>(G, G2, N, F, h->n_sun, M, d_vops + (h->has_gt ? (size_t)(ntau - 1) * M * F : 0), ft, h->d_fields, L, ntau, h->d_rng, h->d_phase, h->d_counters, d_mpos,
                                                               h->n_global_tau, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, place_to, gmtdev, gm_stage, d_place_tab, d_place_pk, gm_stage_f));
  }
  void rotate_group(const VGroup& g, bool in) {     // G <- U^H G U (in) / U G U^H (out) for all vertices of a pair group
    ModelDev m2 = md;
    for (int f = 0; f < F; ++f) { m2.lists[L_TL_FWD][f] = g.rot[in ? 0 : 2][f]; m2.lists[L_TR_FWD][f] = g.rot[in ? 1 : 3][f]; }
    apply_ops(G, 0, MODE_TL_FWD, 0, 0, -1, &m2); apply_ops(G, 1, MODE_TR_FWD, 0, 0, -1, &m2);
  }
  void launch_update(int up, int nt) {
    uint8_t* lg = nullptr;
    if (h->acclog_on && h->d_acclog && h->acclog_pos + M <= h->acclog_per_chain) { lg = h->d_acclog + (long)h->acclog_pos * C; h->acclog_pos += M; }
    int off = 0;
    for (int gi = 0; gi < (int)groups.size(); ++gi) {
      const VGroup& g = groups[up ? gi : (int)groups.size() - 1 - gi];
      if (g.n0 < seq_lo || g.n0 >= seq_hi) continue;                  // not visited sequentially (Nt_sequential_start .. Nt_sequential_end)
      if (g.kind == 0) {
        if (up) KL(KC_UPDATE, st, k_wrapgr<T, 1><<<C, 512, upd_smem, st>>>(G, N, F, h->n_sun, g.n0, g.cnt, M, off, d_vops + (h->has_gt ? (size_t)(nt - 1) * M * F : 0), ft, h->d_fields, L, nt, h->d_rng, h->d_phase, h->d_counters, KD, lg, h->propose_s0, s0dev, stage_g, h->d_fields_c, h->s0_gaussian ? 1 : 0, h->amplitude));
        else KL(KC_UPDATE, st, k_wrapgr<T, 0><<<C, 512, upd_smem, st>>>(G, N, F, h->n_sun, g.n0, g.cnt, M, off, d_vops + (h->has_gt ? (size_t)(nt - 1) * M * F : 0), ft, h->d_fields, L, nt, h->d_rng, h->d_phase, h->d_counters, KD, lg, h->propose_s0, s0dev, stage_g, h->d_fields_c, h->s0_gaussian ? 1 : 0, h->amplitude));
      } else {
        if (g.kind == 2) rotate_group(g, true);
#define FAST_LAUNCH(UPV, IPT, PR) KL(KC_UPDATE, st, k_wrapgr_fast<T, UPV, IPT, PR><<<C, 512, fast_smem, st>>>(G, N, F, h->n_sun, g.n0, g.cnt, M, off, d_vops + (h->has_gt ? (size_t)(nt - 1) * M * F : 0), ft, h->d_fields, L, nt, h->d_rng, h->d_phase, h->d_counters, KDf, ldxf, lg))
        if (g.kind == 1) {
          if (up) { if (iptf == 1) FAST_LAUNCH(1, 1, 0); else FAST_LAUNCH(1, 4, 0); }
          else { if (iptf == 1) FAST_LAUNCH(0, 1, 0); else FAST_LAUNCH(0, 4, 0); }
        } else {
          if (up) { if (iptf == 1) FAST_LAUNCH(1, 1, 1); else FAST_LAUNCH(1, 4, 1); }
          else { if (iptf == 1) FAST_LAUNCH(0, 1, 1); else FAST_LAUNCH(0, 4, 1); }
        }
#undef FAST_LAUNCH
        if (g.kind == 2) rotate_group(g, false);
      }
      off += g.cnt;
    }
  }

  // main.F90:589-631
  void init_sweep() override {
    reset_udv(udvl, 'l'); reset_udv(udvr, 'r'); reset_udv(udvst[S - 1], 'l');
    for (int NST = S - 1; NST >= 1; --NST) { wrapul_on(udvl, stab_nt[NST + 1], stab_nt[NST]); copy_udv(udvst[NST - 1], udvl); }
    wrapul_on(udvl, stab_nt[1], 0);
    cgr_and_phase(1, false);
  }

  // Langevin_HMC_Forces (Prog/Langevin_HMC_mod.F90:107-191, without its measurements): an upward pass without updates,
  // G <- B(nt) G B(nt)^-1 slice by slice with the forces read off G, and the stabilisation of main.F90:731-755 against the stored udvst
  // (which stays untouched here).  Needs the state init_sweep / a finished sweep leaves behind.  Forces on the device: [chain][nt][n] complex.
  cplx* d_forces = nullptr; int* d_lf_site = nullptr; T* d_lf_coef = nullptr; double* d_lf_dt = nullptr;
  void langevin_alloc() {
    if (d_forces) return;
    if (!h->d_fields_c) throw CudaError("Langevin updates need continuous fields (Op_V type 3)");
    d_forces = dalloc<cplx>((size_t)C * L * M); d_lf_dt = dalloc<double>(C);
    std::vector<int> site((size_t)M * F, 0); std::vector<T> coef((size_t)M * F * 2, zero_<T>());
    for (int n = 0; n < M; ++n) for (int f = 0; f < F; ++f) { const HostOp& op = h->opv[n + (size_t)M * f];
      site[(size_t)n * F + f] = op.P[0];
      if (op.type == 3) { coef[((size_t)n * F + f) * 2] = to_T<T>(op.g * op.E[0]); coef[((size_t)n * F + f) * 2 + 1] = to_T<T>(op.g * op.alpha); } }      // O(1,1) = E(1) for N = 1
    d_lf_site = dupload(site); d_lf_coef = dupload(coef);
  }
  void langevin_forces() {
    langevin_alloc();
    reset_udv(udvr, 'r');
    int NST = 1;
    for (int NTAU1 = 1; NTAU1 <= L; ++NTAU1) {
      propr(G, NTAU1); proprm1(G, NTAU1);
      KL(KC_OBS, st, k_langevin_forces<T><<<dim3((M + 127) / 128, C), 128, 0, st>>>(G, n2, N, F, h->n_sun, M, d_lf_site, d_lf_coef, d_is_cont, d_forces, L, NTAU1));
      if (NTAU1 == stab_nt[NST]) {
        wrapur_on(udvr, stab_nt[NST - 1], NTAU1);
        copy_udv(udvl, udvst[NST - 1]);             // udvl = udvst(NST); the storage is intent(in) here
        cgr_and_phase(NTAU1 > L / 2 ? 2 : 1, true);
        NST++;
      }
    }
  }
  // Scheme "HMC" (Prog/Langevin_HMC_mod.F90:393-571), L_Forces = .false., Apply_B_HMC = identity: leapfrog trajectory of all chains, then per chain the
  // Metropolis test on Compute_Ratio_Global and the restore of rejected configurations; ends with Langevin_HMC_Reset_storage.
  double *d_hmc_p = nullptr, *d_hmc_fold = nullptr, *d_hmc_ek = nullptr, *d_hmc_ld[2] = {nullptr, nullptr}, *d_hmc_w = nullptr; cplx *d_hmc_pd[2] = {nullptr, nullptr}, *d_hmc_ga = nullptr, *d_hmc_ph = nullptr;
  unsigned char* d_hmc_acc = nullptr;
  void hmc_update(double delta_t, int leapfrog_steps, double* weight_host, unsigned char* acc_host) override {
    langevin_alloc();
    for (int n = 0; n < M; ++n) if (h->opv[n].type != 3) throw CudaError("HMC moves every field: all vertices must carry continuous (type 3) fields");
    const long per = (long)L * M, tot = per * C;
    if (!d_hmc_p) {
      d_hmc_p = dalloc<double>(tot); d_hmc_fold = dalloc<double>(tot); d_hmc_ek = dalloc<double>(C); d_hmc_w = dalloc<double>(C); d_hmc_acc = dalloc<unsigned char>(C); d_hmc_ph = dalloc<cplx>(C);
      for (int q = 0; q < 2; ++q) { d_hmc_ld[q] = dalloc<double>(NM); d_hmc_pd[q] = dalloc<cplx>(NM); }
      std::vector<cplx> ga((size_t)M * F); for (int n = 0; n < M; ++n) for (int f = 0; f < F; ++f) { const HostOp& op = h->opv[n + (size_t)M * f]; const cd z = (double)h->n_sun * op.g * op.alpha; ga[(size_t)n * F + f] = cplx(z.real(), z.imag()); }
      d_hmc_ga = dupload(ga);
    }
    fermion_det_dev(d_hmc_ld[0], d_hmc_pd[0]);
    CK(cudaMemcpyAsync(d_hmc_fold, h->d_fields_c, sizeof(double) * tot, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(d_hmc_ph, h->d_phase, sizeof(cplx) * C, cudaMemcpyDeviceToDevice, st));
    langevin_forces();
    KL(KC_UPDATE, st, k_hmc_momenta<<<(C + 31) / 32, 32, 0, st>>>(d_hmc_p, h->d_rng, M, L, d_hmc_ek, C));
    dim3 eg(ew_blocks(per), C);
    KL(KC_EW, st, k_hmc_kick<<<eg, 256, 0, st>>>(d_hmc_p, h->d_fields_c, d_forces, h->d_phase, 0.5 * delta_t, per, C));
    for (int t = 1; t <= leapfrog_steps; ++t) {
      KL(KC_EW, st, k_hmc_drift<<<ew_blocks(tot), 256, 0, st>>>(h->d_fields_c, d_hmc_p, delta_t, tot));
      init_sweep();
      const bool last = t == leapfrog_steps;
      if (last) fermion_det_dev(d_hmc_ld[1], d_hmc_pd[1]);
      langevin_forces();
      KL(KC_EW, st, k_hmc_kick<<<eg, 256, 0, st>>>(d_hmc_p, h->d_fields_c, d_forces, h->d_phase, (last ? 0.5 : 1.0) * delta_t, per, C));
    }
    KL(KC_UPDATE, st, k_hmc_decide<<<C, 256, 0, st>>>(h->d_fields_c, d_hmc_fold, d_hmc_p, d_hmc_ek, d_hmc_ld[0], d_hmc_pd[0], d_hmc_ld[1], d_hmc_pd[1], d_hmc_ga, d_hmc_ph,
                                                      F, h->n_sun, M, L, h->d_rng, d_hmc_w, d_hmc_acc));
    init_sweep();
    if (weight_host) CK(cudaMemcpyAsync(weight_host, d_hmc_w, sizeof(double) * C, cudaMemcpyDeviceToHost, st));
    if (acc_host) CK(cudaMemcpyAsync(acc_host, d_hmc_acc, C, cudaMemcpyDeviceToHost, st));
    sync();
  }
  void langevin_get_forces(cd* out) override {
    langevin_forces(); std::vector<cplx> b((size_t)C * L * M);
    CK(cudaMemcpyAsync(b.data(), d_forces, sizeof(cplx) * b.size(), cudaMemcpyDeviceToHost, st)); sync();
    for (size_t i = 0; i < b.size(); ++i) out[i] = cd(b[i].x, b[i].y);
  }
  void langevin_update(double delta_t, double max_force, double* dt_running_host) override {
    langevin_forces();
    KL(KC_UPDATE, st, k_langevin_update<<<C, 256, 0, st>>>(h->d_fields_c, d_forces, d_is_cont, M, L, h->d_rng, h->d_phase, delta_t, max_force, d_lf_dt, C));
    init_sweep();                                   // Langevin_HMC_Reset_storage (:228-285)
    if (dt_running_host) { CK(cudaMemcpyAsync(dt_running_host, d_lf_dt, sizeof(double) * C, cudaMemcpyDeviceToHost, st)); }
    sync();
  }

  // Compute_Fermion_Det with storage = "Empty" (Prog/Global_mod.F90:792-1000): the left propagation is rebuilt from the current fields
  // (udvl, udvst as after main.F90:589-627), then per chain and flavor log|det| and phase of the fermion determinant.  Finite temperature:
  // det(1 + B(beta,0)) through TP = U_L + V_L D_L; the reference factorises TP by QR + SVD (UDV_WRAP) and keeps the individual log singular values,
  // of which only the sum enters Compute_Ratio_Global (:700-760) -- here the sum comes from the pivoted QR.  Projector: sum of the log D of all
  // stored decompositions + log|det(U_L^H P_R)| (:846-884).
  void fermion_det(double* logdet_host, cd* phase_host) override {
    double* d_ld = dalloc_tmp<double>(NM); cplx* d_ph = dalloc_tmp<cplx>(NM);
    fermion_det_dev(d_ld, d_ph);
    std::vector<cplx> ph(NM);
    CK(cudaMemcpyAsync(logdet_host, d_ld, sizeof(double) * NM, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(ph.data(), d_ph, sizeof(cplx) * NM, cudaMemcpyDeviceToHost, st)); sync();
    for (int b = 0; b < NM; ++b) phase_host[b] = cd(ph[b].x, ph[b].y);
    cudaFree(d_ld); cudaFree(d_ph);
  }
  void fermion_det_dev(double* d_ld, cplx* d_ph) {
    reset_udv(udvl, 'l'); reset_udv(udvst[S - 1], 'l');
    double* d_extra = dalloc_tmp<double>(NM);
    CK(cudaMemsetAsync(d_extra, 0, sizeof(double) * NM, st));
    for (int NST = S - 1; NST >= 1; --NST) { wrapul_on(udvl, stab_nt[NST + 1], stab_nt[NST]); copy_udv(udvst[NST - 1], udvl);
      if (proj) KL(KC_EW, st, k_fdet_addlog<<<(NM + 127) / 128, 128, 0, st>>>(udvl.D, N, NP, d_extra, NM)); }
    wrapul_on(udvl, stab_nt[1], 0);
    if (proj) {
      KL(KC_EW, st, k_fdet_addlog<<<(NM + 127) / 128, 128, 0, st>>>(udvl.D, N, NP, d_extra, NM));
      // S = U_L^H P_R (N_part x N_part), P_R per flavor
      reset_udv(udvr, 'r');
      gemm<T, 1, 0, 0>(st, NP, NP, N, udvl.U, N, n2, udvr.U, N, n2, wp.W[2], NP, wp.n2(), NM);
      la_qrp<T>(wp, wp.W[2], NP, NP, wp.Dq);
      KL(KC_EW, st, k_fdet_finish<<<(NM + 127) / 128, 128, 0, st>>>(wp.Dq, NP, wp.qrout, d_extra, nullptr, d_ld, d_ph, NM));
    } else {
      KL(KC_EW, st, k_fdet_build<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(w.W[0], udvl.U, udvl.V, udvl.D, n2, N, d_extra));
      la_qrp<T>(w, w.W[0], N, N, w.Dq);
      KL(KC_EW, st, k_fdet_finish<<<(NM + 127) / 128, 128, 0, st>>>(w.Dq, N, w.qrout, d_extra, udvl.det, d_ld, d_ph, NM));
    }
    sync(); cudaFree(d_extra);
  }
  template <typename X> X* dalloc_tmp(size_t n) { X* p = nullptr; CK(cudaMalloc(&p, sizeof(X) * (n ? n : 1))); return p; }

  // where main.F90:757-773 / 789-802 call ham%Obser: NTAU1 in [LOBS_ST, LOBS_EN] (defaults of QMC_runtime_var_mod.F90:156-189)
  void measure_hook(int ntau1) {
    const int lobs_st = h->lobs_st > 0 ? h->lobs_st : (proj ? thtrot + 1 : 1), lobs_en = h->lobs_en > 0 ? h->lobs_en : (proj ? L - thtrot : L);
    if (ntau1 < lobs_st || ntau1 > lobs_en) return;
    KL(KC_OBS, st, k_obs_scalar<T><<<C, 128, 0, st>>>(G, n2, N, F, h->n_sun, h->d_phase, h->d_obs));
    const bool scal_tab = h->n_kin > 0 || h->n_pot > 0;
    if (h->obs_eq_on || scal_tab) {      // observables on the symmetrised G (main.F90:761-764 hands GR_Tilde to ham%Obser)
      if (h->obs_eq_on) obs_tau_setup(); else if (!obsS[0]) { for (int q = 0; q < 2; ++q) obsS[q] = dalloc<T>(n2 * NM); }
      const T* gs = G;
      if (h->symm) {
        if (dense_t) { CK(cudaMemcpyAsync(obsS[0], G, sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st)); hop_symm(obsS[0]); }
        else { apply_ops(G, 0, MODE_TL_HALF, 0, 0, -1, nullptr, obsS[0]); apply_ops(obsS[0], 1, MODE_TR_HALFINV, 0, 0); }
        gs = obsS[0];
      }
      if (scal_tab) KL(KC_OBS, st, k_obs_scal_tables<T><<<C, 256, 0, st>>>(gs, n2, N, F, h->n_sun, h->d_phase, h->n_kin, h->d_kin_idx, h->d_kin_coef, h->n_pot, h->d_pot_idx, h->d_pot_coef, h->d_obs));
      if (h->obs_eq_on) {
        KL(KC_EW, st, k_g0t_init<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(obsS[1], gs, n2, N));      // G - 1
        if (lt.shift32) KL(KC_OBS, st, k_obs_tau_diag<T, 1><<<dim3(N / 32, C), 256, obst_smem, st>>>(gs, obsS[1], gs, gs, n2, N, F, h->n_sun, h->d_phase, lt, 0, 1, h->d_obse_acc, h->d_obse_bg, h->d_obse_cnt));
        else KL(KC_OBS, st, k_obs_tau<T, 1><<<C, 256, obst_smem, st>>>(gs, obsS[1], gs, gs, n2, N, F, h->n_sun, h->d_phase, lt, 0, 1, h->d_obse_acc, h->d_obse_bg, h->d_obse_cnt));
      }
    }
  }

  // main.F90:714-887
  void sweep(int ltau) override {
    reset_udv(udvr, 'r');
    int NST = 1;
    for (int NTAU = 0; NTAU <= L - 1; ++NTAU) {
      const int NTAU1 = NTAU + 1;
      wrapgrup(NTAU);
      if (NTAU1 == stab_nt[NST]) {
        wrapur_on(udvr, stab_nt[NST - 1], NTAU1);
        std::swap(udvl, udvst[NST - 1]);            // udvl = udvst(NST) (old udvl content is dead)
        copy_udv(udvst[NST - 1], udvr);             // udvst(NST) = udvr
        cgr_and_phase(NTAU1 > L / 2 ? 2 : 1, true);
        NST++;
      }
      measure_hook(NTAU1);
    }
    reset_udv(udvl, 'l');
    NST = S - 1;
    for (int NTAU = L; NTAU >= 1; --NTAU) {
      const int NTAU1 = NTAU - 1;
      wrapgrdo(NTAU);
      measure_hook(NTAU1);
      if (NST >= 0 && stab_nt[NST] == NTAU1 && NTAU1 != 0) {
        wrapul_on(udvl, stab_nt[NST + 1], NTAU1);
        std::swap(udvr, udvst[NST - 1]);            // udvr = udvst(NST)
        copy_udv(udvst[NST - 1], udvl);             // udvst(NST) = udvl
        cgr_and_phase(NTAU1 > L / 2 ? 2 : 1, true);
        if (ltau == 1 && proj && stab_nt[NST] <= thtrot + 1 && thtrot + 1 < stab_nt[NST + 1]) tau_p(NST);     // main.F90:829-831
        NST--;
      }
    }
    wrapul_on(udvl, stab_nt[1], stab_nt[0]);
    reset_udv(udvr, 'r');
    cgr_and_phase(1, true);
    reset_udv(udvst[S - 1], 'l');
    if (ltau == 1 && !proj) tau_m();
    if (ltau == 1 && proj && stab_nt[1] > thtrot + 1) tau_p(0);                                              // main.F90:884-886
  }

  // ---------------------------------------------------------------- TAU_M (Prog/tau_m_mod.F90:56-211) for all chains
  LaWork<T> w2; bool w2_ready = false; T* tmN[4] = {nullptr, nullptr, nullptr, nullptr}; int* d_first = nullptr; T* capbuf = nullptr;
  bool taum_ready = false;
  void taum_alloc() {
    if (taum_ready) return;
    if (!proj) { w2.alloc(2 * N, NM, st); w2_ready = true; }
    GT0 = dalloc<T>(n2 * NM); G0T = dalloc<T>(n2 * NM); G00 = dalloc<T>(n2 * NM); GTT = dalloc<T>(n2 * NM);
    for (int q = 0; q < 4; ++q) tmN[q] = dalloc<T>(n2 * NM);
    udvr2 = alloc_udv(); d_first = dalloc<int>(NM); capbuf = dalloc<T>(n2 * NM);
    taum_ready = true;
  }
  void taum_capture(int nt, bool fresh = false) {     // what ham%ObserT receives (tau_m_mod.F90:115-124,151-177); test support only
    if (!h->taum_every) return;
    if (!fresh && h->taum_every > 1 && (nt % h->taum_every) != 0) return;
    std::vector<std::vector<cd>>& dst = fresh ? h->taum_fresh_host : h->taum_host;
    if (dst.size() != (size_t)C) dst.assign(C, std::vector<cd>());
    T* arr[4] = {GT0, G0T, G00, GTT};
    std::vector<T> host(n2 * NM);
    std::vector<std::vector<cd>> part(C, std::vector<cd>((size_t)4 * F * n2));
    for (int q = 0; q < 4; ++q) {
      CK(cudaMemcpyAsync(capbuf, arr[q], sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st));
      if (h->symm && !fresh) hop_symm(capbuf);
      CK(cudaMemcpyAsync(host.data(), capbuf, sizeof(T) * n2 * NM, cudaMemcpyDeviceToHost, st)); sync();
      for (int c = 0; c < C; ++c) for (int f = 0; f < F; ++f) for (long i = 0; i < n2; ++i)
        part[c][((size_t)q * F + f) * n2 + i] = from_T<T>(host[n2 * ((long)c * F + f) + i]);
    }
    for (int c = 0; c < C; ++c) dst[c].insert(dst[c].end(), part[c].begin(), part[c].end());
  }
  // ---- device-side ObserT (alf_obs_tau.cuh): symmetrised copies of the four matrices, then one binning kernel per time point
  LattDev lt; T* obsS[4] = {nullptr, nullptr, nullptr, nullptr}; size_t obst_smem = 0;
  const T* g00_sym_of = nullptr;     // which buffer obsS[2] is the symmetrised copy of (reset whenever G00's content is rewritten)
  void obs_tau_setup() override {
    if (lt.cell) return;
    if (h->n_unit <= 0 || (int)h->site_cell.size() != N) throw CudaError("obs_tau: alf_b200_set_lattice has not been called");
    if (F > 2) throw CudaError("obs_tau: more than two flavors are not supported");
    lt.n_unit = h->n_unit; lt.norb = h->norb;
    lt.cell = dupload(h->site_cell); lt.orb = dupload(h->site_orb); lt.imj = dupload(h->imj);
    {   // is the site numbering invariant under a shift by 32 sites (a lattice translation that keeps orbitals and lattice displacements)?  -> k_obs_tau_diag
      bool ok = (N % 32 == 0) && N >= 64 && !getenv("ALF_B200_NO_OBS_DIAG");
      for (int i = 0; i < N && ok; ++i) { const int i2 = (i + 32) % N; if (h->site_cell[i] < 0 || h->site_cell[i2] < 0 || h->site_orb[i] != h->site_orb[i2]) ok = false; }
      for (int i = 0; i < N && ok; ++i) for (int j = 0; j < N && ok; ++j)
        if (h->imj[h->site_cell[(i + 32) % N] + (size_t)h->site_cell[(j + 32) % N] * h->n_unit] != h->imj[h->site_cell[i] + (size_t)h->site_cell[j] * h->n_unit]) ok = false;
      lt.shift32 = ok ? 1 : 0;
    }
    for (int q = 0; q < 4; ++q) if (!obsS[q]) obsS[q] = dalloc<T>(n2 * NM);
    obst_smem = obs_tau_smem<T>(N, lt.n_unit, lt.norb);
    if (obst_smem > 227 * 1024) throw CudaError("obs_tau: lattice too large for the shared-memory bins");
    CK(alf_raise_smem(k_obs_tau<T, 0>));
    CK(alf_raise_smem(k_obs_tau<T, 1>));
    CK(alf_raise_smem(k_obs_tau_diag<T, 0>));
    CK(alf_raise_smem(k_obs_tau_diag<T, 1>));
  }
  void obsert(int nt_index) {      // where TAU_M / Tau_p call ham%ObserT(nt_index, GT0, G0T, G00, GTT, Phase): tau_m_mod.F90:115-124,151-177
    if (!h->obs_tau_on || nt_index < 0 || nt_index >= h->obst_ntau) return;
    obs_tau_setup();
    const T* src[4] = {GT0, G0T, G00, GTT}; const T* use[4];
    for (int q = 0; q < 4; ++q) {
      if (!h->symm) { use[q] = src[q]; continue; }
      use[q] = obsS[q];
      if (q == 2 && g00_sym_of == G00) continue;          // G(0,0) only changes at the stabilisation points: its symmetrised copy is reused
      if (dense_t) { CK(cudaMemcpyAsync(obsS[q], src[q], sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st)); hop_symm(obsS[q]); }
      else {     // Hop_mod_Symm out of place: the left half step reads the source and writes the copy, the right one works on the copy
        apply_ops(const_cast<T*>(src[q]), 0, MODE_TL_HALF, 0, 0, -1, nullptr, obsS[q]); apply_ops(obsS[q], 1, MODE_TR_HALFINV, 0, 0);
      }
      if (q == 2) g00_sym_of = G00;
    }
    if (lt.shift32) KL(KC_OBS, st, k_obs_tau_diag<T, 0><<<dim3(N / 32, C), 256, obst_smem, st>>>(use[0], use[1], use[2], use[3], n2, N, F, h->n_sun, h->d_phase, lt, nt_index, h->obst_ntau,
                                                                                 h->d_obst_acc, h->d_obst_bg, h->d_obst_cnt));
    else KL(KC_OBS, st, k_obs_tau<T, 0><<<C, 256, obst_smem, st>>>(use[0], use[1], use[2], use[3], n2, N, F, h->n_sun, h->d_phase, lt, nt_index, h->obst_ntau,
                                                        h->d_obst_acc, h->d_obst_bg, h->d_obst_cnt));
  }
  void compare_tau(const T* A, const T* B) {   // Control_Precision_tau, control_mod.F90:300-311
    KL(KC_EW, st, k_compare<T><<<dim3(NM, CMP_SPLIT), 256, 0, st>>>(A, B, n2, n2, d_cmp));
    KL(KC_EW, st, k_ctl_accum<<<(C + 127) / 128, 128, 0, st>>>(d_cmp, F, h->d_ctl, 1, C));
  }
  void tau_m() override {
    taum_alloc();
    const size_t bytes = sizeof(T) * n2 * NM; dim3 eg(ew_blocks(n2), NM);
    g00_sym_of = nullptr;
    CK(cudaMemcpyAsync(G00, G, bytes, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(GT0, G, bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(GTT, G, bytes, cudaMemcpyDeviceToDevice, st));
    KL(KC_EW, st, k_g0t_init<T><<<eg, 256, 0, st>>>(G0T, G, n2, N));
    taum_capture(0); obsert(0);
    set_udv_identity(udvr2);
    int NST = 1;
    for (int NT = 0; NT <= L - 1; ++NT) {
      const int NT1 = NT + 1;
      propr(GT0, NT1); proprm1(G0T, NT1); proprm1(GTT, NT1); propr(GTT, NT1);          // tau_m_mod.F90:141-149
      taum_capture(NT1); obsert(NT1);
      if (stab_nt[NST] == NT1) {
        wrapur_on(udvr2, stab_nt[NST - 1], NT1);
        la_cgr2_2<T>(w, w2, h->stab, udvr2, udvst[NST - 1], tmN[0], tmN[1], tmN[2], tmN[3], d_first);
        compare_tau(G, tmN[1]); compare_tau(GTT, tmN[2]); compare_tau(GT0, tmN[0]); compare_tau(G0T, tmN[3]);
        std::swap(GT0, tmN[0]); std::swap(G00, tmN[1]); std::swap(GTT, tmN[2]); std::swap(G0T, tmN[3]); g00_sym_of = nullptr;
        taum_capture(NT1, true);
        NST++;
      }
    }
  }
  // ---------------------------------------------------------------- Tau_p (Prog/tau_p_mod.F90:74-336, sequential update path) for all chains
  void tau_p(int NST_IN) override {
    if (!proj) throw CudaError("tau_p: the handle was not set up for the projective algorithm");
    taum_alloc();
    const size_t bytes = sizeof(T) * n2 * NM; dim3 eg(ew_blocks(n2), NM);
    copy_udv(udvr2, udvr);                                     // udvr_local
    CK(cudaMemcpyAsync(GTT, G, bytes, cudaMemcpyDeviceToDevice, st));
    int NT_ST = NST_IN;
    T* GRUP = tmN[0];
    auto restab = [&]() {                                      // Wrapur + CGRP + Control_Precision_tau (:131-142, :198-208)
      wrapur_on(udvr2, stab_nt[NT_ST], stab_nt[NT_ST + 1]);
      la_cgrp<T>(w, wp, udvr2, udvst[NT_ST], GRUP, d_z);       // udvst(nt_st + 1) in the reference's 1-based numbering
      compare_tau(GTT, GRUP);
    };
    for (int NT = stab_nt[NT_ST] + 1; NT <= thtrot + 1; ++NT) {
      proprm1(GTT, NT); propr(GTT, NT);
      if (NT_ST + 1 <= S && NT == stab_nt[NT_ST + 1]) { restab(); CK(cudaMemcpyAsync(GTT, GRUP, bytes, cudaMemcpyDeviceToDevice, st)); NT_ST++; }
    }
    g00_sym_of = nullptr;
    CK(cudaMemcpyAsync(G00, GTT, bytes, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(GT0, GTT, bytes, cudaMemcpyDeviceToDevice, st));
    KL(KC_EW, st, k_g0t_init<T><<<eg, 256, 0, st>>>(G0T, GTT, n2, N));                     // G0T = GTT - 1
    taum_capture(0); obsert(0);
    int NCHECK = 0;
    for (int NT = thtrot + 1; NT <= L - thtrot; ++NT) {
      const int NTAU = NT - thtrot - 1;
      if (NT_ST + 1 <= S && NT == stab_nt[NT_ST + 1] && NTAU != 0) {
        restab(); NT_ST++; NCHECK++;
        CK(cudaMemcpyAsync(GTT, GRUP, bytes, cudaMemcpyDeviceToDevice, st));
        gemm<T, 0, 0, 0>(st, N, N, N, GRUP, N, n2, GT0, N, n2, tmN[1], N, n2, NM); std::swap(GT0, tmN[1]);          // GT0 = G GT0
        KL(KC_EW, st, k_axpb_identity<T><<<eg, 256, 0, st>>>(tmN[2], GRUP, n2, N, -1.0, 1.0));                    // 1 - G
        gemm<T, 0, 0, 0>(st, N, N, N, G0T, N, n2, tmN[2], N, n2, tmN[1], N, n2, NM); std::swap(G0T, tmN[1]);        // G0T = G0T (1 - G)
        taum_capture(NT, true);
      }
      const int NT1 = NT + 1;
      propr(GT0, NT1); proprm1(G0T, NT1); proprm1(GTT, NT1); propr(GTT, NT1);
      taum_capture(NTAU + 1); obsert(NTAU + 1);
    }
    if (NCHECK == 0 && NT_ST + 1 <= S) {                       // fallback check of the reference (:262-279)
      for (int NT = L - thtrot + 2; NT <= stab_nt[NT_ST + 1]; ++NT) { proprm1(GTT, NT); propr(GTT, NT); }
      restab(); NT_ST++;
    }
  }

  // ---------------------------------------------------------------- global-in-slice moves (Prog/Wrapgr_mod.F90:247-433)
  int* d_mpos = nullptr;
  // shared memory of k_random_update; G is staged there when it fits (gm_stage)
  int gm_stage = 0, gm_stage_f = 0;
  size_t gm_smem() {
    const size_t base = sizeof(T) * ((size_t)4 * N + 2 * ALF_KMAX * ALF_KMAX) + 64, staged = base + sizeof(T) * (size_t)F * (N | 1) * N;
    gm_stage = (staged <= 220 * 1024 && !getenv("ALF_B200_NO_STAGE_G")) ? 1 : 0;
    size_t sz = gm_stage ? staged : base;
    gm_stage_f = (sz + (size_t)M + 16 <= 224 * 1024) ? 1 : 0;
    return sz + (gm_stage_f ? (size_t)M + 16 : 0);
  }
  void gm_alloc() { if (!d_mpos) { d_mpos = dalloc<int>(C); CK(cudaMemsetAsync(d_mpos, 0, sizeof(int) * C, st)); } }
  void gm_set_position(int m) override { gm_alloc(); KL(KC_EW, st, k_fill_int<<<ew_blocks(C), 256, 0, st>>>(d_mpos, C, m)); }
  void gm_get_position(int* m) override { gm_alloc(); sync(); CK(cudaMemcpy(m, d_mpos, sizeof(int) * C, cudaMemcpyDeviceToHost)); }
  void gm_random_update(int ntau, int n_moves, int maxlen, const int* len, const int* list0, const int8_t* val, const double* t0, const double* s0,
                        uint8_t* acc_out, int place_to) override {
    gm_alloc();
    const size_t np = (size_t)C * std::max(n_moves, 1), nl = np * std::max(maxlen, 1);
    struct Scratch { void* p = nullptr; ~Scratch() { if (p) cudaFree(p); } };       // freed on every exit path, including exceptions
    Scratch b_len, b_list, b_val, b_t0, b_s0, b_acc;
    CK(cudaMalloc(&b_len.p, sizeof(int) * np)); CK(cudaMalloc(&b_list.p, sizeof(int) * nl)); CK(cudaMalloc(&b_val.p, nl));
    CK(cudaMalloc(&b_t0.p, sizeof(double) * np)); CK(cudaMalloc(&b_s0.p, sizeof(double) * np)); CK(cudaMalloc(&b_acc.p, np));
    int *d_len = (int*)b_len.p, *d_list = (int*)b_list.p; int8_t* d_val = (int8_t*)b_val.p; double *d_t0 = (double*)b_t0.p, *d_s0 = (double*)b_s0.p; uint8_t* d_acc = (uint8_t*)b_acc.p;
    if (n_moves > 0) {
      CK(cudaMemcpyAsync(d_len, len, sizeof(int) * np, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(d_list, list0, sizeof(int) * nl, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(d_val, val, nl, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(d_t0, t0, sizeof(double) * np, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(d_s0, s0, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    }
    const size_t smem = gm_smem();
    CK(alf_raise_smem(k_random_update<T>));
    KL(KC_UPDATE, st, k_random_update<T><<<C, 512, smem, st>>>(G, G2, N, F, h->n_sun, M, d_vops + (h->has_gt ? (size_t)(ntau - 1) * M * F : 0), ft, h->d_fields, L, ntau, h->d_rng, h->d_phase, h->d_counters, d_mpos,
                                                               n_moves, std::max(maxlen, 1), d_len, d_list, d_val, d_t0, d_s0, d_acc, place_to, GmtDev{0, 0, nullptr, nullptr, {0, 0, nullptr, nullptr, nullptr, nullptr, nullptr}}, gm_stage, d_place_tab, d_place_pk, gm_stage_f));
    if (acc_out && n_moves > 0) { CK(cudaMemcpyAsync(acc_out, d_acc, np, cudaMemcpyDeviceToHost, st)); }
    sync();
  }

  // PROPR / PROPRM1 (Prog/tau_m_mod.F90:215-263): A <- B(nt) A ;  A <- A B(nt)^-1
  void propr(T* A, int nt) {
    if (!dense_t) apply_ops(A, 0, MODE_WRAPUR, nt, nt);
    else { dense_mult(A, 0, true); apply_ops(A, 0, MODE_WRAPUR, nt, nt); }
  }
  void proprm1(T* A, int nt) {
    if (!dense_t) apply_ops(A, 1, MODE_PROPRM1, nt, nt);
    else { dense_mult(A, 1, false); apply_ops(A, 1, MODE_PROPRM1, nt, nt); }
  }

  // ---------------------------------------------------------------- host access
  void get_green(int chain, int nf, int symm, cd* out) override {
    const T* src = G + n2 * ((long)chain * F + (nf - 1));
    std::vector<T> t(n2);
    if (symm) {   // Hop_mod_Symm on a copy (main.F90:761-764): do it for the whole batch in W[3]... only this matrix is read back
      CK(cudaMemcpyAsync(G2, G, sizeof(T) * n2 * NM, cudaMemcpyDeviceToDevice, st));
      hop_symm(G2); src = G2 + n2 * ((long)chain * F + (nf - 1));
    }
    CK(cudaMemcpyAsync(t.data(), src, sizeof(T) * n2, cudaMemcpyDeviceToHost, st)); sync();
    for (long i = 0; i < n2; ++i) out[i] = from_T<T>(t[i]);
  }
  void set_green(int chain, int nf, const cd* in) override {
    std::vector<T> t(n2); for (long i = 0; i < n2; ++i) t[i] = to_T<T>(in[i]);
    CK(cudaMemcpyAsync(G + n2 * ((long)chain * F + (nf - 1)), t.data(), sizeof(T) * n2, cudaMemcpyHostToDevice, st)); sync();
  }
  void get_udv(int which, int nst, int chain, int nf, cd* U, cd* D, cd* V) override {
    UdvDev<T>& u = which == 0 ? udvl : which == 1 ? udvr : udvst[nst - 1];
    const long b = (long)chain * F + (nf - 1);
    std::vector<T> t(n2); std::vector<double> d(N);
    CK(cudaMemcpyAsync(t.data(), u.U + n2 * b, sizeof(T) * n2, cudaMemcpyDeviceToHost, st)); sync(); for (long i = 0; i < n2; ++i) U[i] = from_T<T>(t[i]);
    CK(cudaMemcpyAsync(t.data(), u.V + n2 * b, sizeof(T) * n2, cudaMemcpyDeviceToHost, st)); sync(); for (long i = 0; i < n2; ++i) V[i] = from_T<T>(t[i]);
    CK(cudaMemcpyAsync(d.data(), u.D + (long)N * b, sizeof(double) * N, cudaMemcpyDeviceToHost, st)); sync(); for (int i = 0; i < N; ++i) D[i] = cd(d[i], 0.0);
  }
  // host UDV_State -> device (compat mode: WRAPUR / WRAPUL / CGR on the caller's states); det U is recomputed on the device
  void set_udv(int which, int nst, int chain, int nf, const cd* U, const cd* D, const cd* V) override {
    UdvDev<T>& u = which == 0 ? udvl : which == 1 ? udvr : udvst[nst - 1];
    const long b = (long)chain * F + (nf - 1);
    std::vector<T> t(n2, zero_<T>()); std::vector<double> d(N, 1.0);
    const long nu = proj ? (long)N * NP : n2; const int nd = proj ? NP : N;      // projector states carry U(Ndim, N_part), D(N_part) and no V
    for (long i = 0; i < nu; ++i) t[i] = to_T<T>(U[i]);
    CK(cudaMemcpyAsync(u.U + n2 * b, t.data(), sizeof(T) * n2, cudaMemcpyHostToDevice, st)); sync();
    if (!proj && V) { for (long i = 0; i < n2; ++i) t[i] = to_T<T>(V[i]); CK(cudaMemcpyAsync(u.V + n2 * b, t.data(), sizeof(T) * n2, cudaMemcpyHostToDevice, st)); sync(); }
    for (int i = 0; i < nd; ++i) d[i] = D[i].real();
    CK(cudaMemcpyAsync(u.D + (long)N * b, d.data(), sizeof(double) * N, cudaMemcpyHostToDevice, st)); sync();
    if (!proj) {      // det U of this one matrix: pivoted QR of a copy (batch of one inside the workspace)
      LaWork<T> w1; w1.alloc(N, 1, st);
      CK(cudaMemcpyAsync(w1.W[0], u.U + n2 * b, sizeof(T) * n2, cudaMemcpyDeviceToDevice, st));
      la_qrp<T>(w1, w1.W[0], N, N, w1.Dq);
      k_det_phase<<<1, 32, 0, st>>>(w1.qrout, u.det + b, 1); sync(); w1.release();
    }
  }
  void hop_apply(int which, int nf, cd* A) override {
    // applies to matrix slot (chain 0, flavor nf) of scratch G2; the whole batch is processed (test entry point)
    std::vector<T> t(n2); for (long i = 0; i < n2; ++i) t[i] = to_T<T>(A[i]);
    CK(cudaMemsetAsync(G2, 0, sizeof(T) * n2 * NM, st));
    CK(cudaMemcpyAsync(G2 + n2 * (nf - 1), t.data(), sizeof(T) * n2, cudaMemcpyHostToDevice, st));
    switch (which) { case 0: mmthr(G2); break; case 1: mmthr_m1(G2); break; case 2: mmthl(G2); break; case 3: mmthl_m1(G2); break; case 4: mmthlc(G2); break; case 5: hop_symm(G2); break; }
    CK(cudaMemcpyAsync(t.data(), G2 + n2 * (nf - 1), sizeof(T) * n2, cudaMemcpyDeviceToHost, st)); sync();
    for (long i = 0; i < n2; ++i) A[i] = from_T<T>(t[i]);
  }
};

// ---- kernel-level test helpers and the FP64 peak microbenchmark (templates must live outside extern "C")
template <typename T> static std::vector<T> h2T(const double* src, size_t n) { std::vector<T> v(n); for (size_t i = 0; i < n; ++i) v[i] = to_T<T>(cd(src[2 * i], src[2 * i + 1])); return v; }
template <typename T> static void T2h(const std::vector<T>& v, double* dst) { for (size_t i = 0; i < v.size(); ++i) { cd z = from_T<T>(v[i]); dst[2 * i] = z.real(); dst[2 * i + 1] = z.imag(); } }
template <typename T> struct DevBuf { T* p = nullptr; size_t n = 0; DevBuf(size_t nn) : n(nn) { CK(cudaMalloc(&p, sizeof(T) * (nn ? nn : 1))); } ~DevBuf() { cudaFree(p); }
  void up(const std::vector<T>& v) { CK(cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice)); } std::vector<T> down() { std::vector<T> v(n); CK(cudaMemcpy(v.data(), p, sizeof(T) * n, cudaMemcpyDeviceToHost)); return v; } };

template <typename T>
static void t_qdrp(int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases) {
  DevBuf<T> dA((size_t)m * n * batch), dtau((size_t)n * batch); DevBuf<int> dp((size_t)n * batch); DevBuf<double> dD((size_t)n * batch); DevBuf<QrOut> dq(batch);
  dA.up(h2T<T>(A, dA.n));
  launch_qrp<T, 1>(0, dA.p, m, n, m, (long)m * n, dtau.p, n, dp.p, n, dD.p, n, dq.p, batch);
  CK(cudaDeviceSynchronize());
  T2h<T>(dA.down(), A); T2h<T>(dtau.down(), tau);
  auto d = dD.down(); std::copy(d.begin(), d.end(), D);
  auto p = dp.down(); for (size_t i = 0; i < p.size(); ++i) jpvt[i] = p[i] + 1;
  auto q = dq.down(); for (int b = 0; b < batch; ++b) { phases[5 * b] = q[b].perm_sign; phases[5 * b + 1] = q[b].diag_phase.x; phases[5 * b + 2] = q[b].diag_phase.y; phases[5 * b + 3] = q[b].detq.x; phases[5 * b + 4] = q[b].detq.y; }
}
// blocked (windowed-pivoting) QR + explicit Q through the compact-WY application: returns QR-in-place, D, jpvt, tau, phases and Q
template <typename T>
static void t_qdrp_blk(int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases, double* Q) {
  DevBuf<T> dA((size_t)m * n * batch), dtau((size_t)n * batch), dQ((size_t)m * m * batch), dT((size_t)(n + 32) * 32 * batch); DevBuf<int> dp((size_t)n * batch); DevBuf<double> dD((size_t)n * batch); DevBuf<QrOut> dq(batch);
  dA.up(h2T<T>(A, dA.n));
  if (!qrblk_cfg<T>(m, n).ok) throw CudaError("blocked QR: matrix too large for shared memory");
  launch_qrp_blk<T>(0, dA.p, m, n, m, (long)m * n, dtau.p, n, dp.p, n, dD.p, n, dq.p, dT.p, batch);
  KL(KC_EW, 0, k_set_identity<T><<<dim3(ew_blocks((long)m * m), batch), 256>>>(dQ.p, m, (long)m * m, m, m));
  launch_apply_q<T>(0, dA.p, m, n, m, (long)m * n, dT.p, dQ.p, m, (long)m * m, m, 1, true, batch);
  CK(cudaDeviceSynchronize());
  T2h<T>(dA.down(), A); T2h<T>(dtau.down(), tau); T2h<T>(dQ.down(), Q);
  auto d = dD.down(); std::copy(d.begin(), d.end(), D);
  auto p = dp.down(); for (size_t i = 0; i < p.size(); ++i) jpvt[i] = p[i] + 1;
  auto q = dq.down(); for (int b = 0; b < batch; ++b) { phases[5 * b] = q[b].perm_sign; phases[5 * b + 1] = q[b].diag_phase.x; phases[5 * b + 2] = q[b].diag_phase.y; phases[5 * b + 3] = q[b].detq.x; phases[5 * b + 4] = q[b].detq.y; }
}
// UDV_Wrap_Pivot(A, U, D, V, NCON, N1, N2) on a batch of host matrices (complex layout at the boundary)
template <typename T>
static void t_udv_wrap_pivot(int n1, int n2, int batch, const double* A, double* U, double* D, double* V) {
  DevBuf<T> dA((size_t)n1 * n2 * batch), dU((size_t)n1 * n2 * batch), dV((size_t)n2 * n2 * batch); DevBuf<double> dD((size_t)n2 * batch);
  dA.up(h2T<T>(A, dA.n));
  la_udv_wrap_pivot<T>(0, dA.p, dU.p, dD.p, dV.p, n1, n2, batch); CK(cudaDeviceSynchronize());
  T2h<T>(dU.down(), U); T2h<T>(dV.down(), V); auto d = dD.down(); for (size_t i = 0; i < d.size(); ++i) { D[2 * i] = d[i]; D[2 * i + 1] = 0.0; }
}
template <typename T>
static void t_udv(int n, int batch, char side, double* U, double* D, double* V) {
  const size_t n2 = (size_t)n * n; LaWork<T> w; w.alloc(n, batch, 0);
  DevBuf<T> dU(n2 * batch), dV(n2 * batch); DevBuf<double> dD((size_t)n * batch); DevBuf<cplx> ddet(batch);
  dU.up(h2T<T>(U, dU.n)); dV.up(h2T<T>(V, dV.n)); { std::vector<double> d((size_t)n * batch); for (size_t i = 0; i < d.size(); ++i) d[i] = D[2 * i]; dD.up(d); }
  UdvDev<T> s; s.U = dU.p; s.V = dV.p; s.D = dD.p; s.det = ddet.p;
  la_decompose<T>(w, s, side); CK(cudaDeviceSynchronize());
  T2h<T>(dU.down(), U); T2h<T>(dV.down(), V); auto d = dD.down(); for (size_t i = 0; i < d.size(); ++i) { D[2 * i] = d[i]; D[2 * i + 1] = 0.0; }
  w.release();
}
template <typename T>
static void t_cgr(int n, int batch, int nvar, int stab, const double* UR, const double* DR, const double* VR, const double* UL, const double* DL, const double* VL,
                  const double* detUR, const double* detUL, double* G, double* phase) {
  const size_t n2 = (size_t)n * n; LaWork<T> w; w.alloc(n, batch, 0);
  DevBuf<T> dUR(n2 * batch), dVR(n2 * batch), dUL(n2 * batch), dVL(n2 * batch), dG(n2 * batch); DevBuf<double> dDR((size_t)n * batch), dDL((size_t)n * batch);
  DevBuf<cplx> detR(batch), detL(batch), dz(batch);
  dUR.up(h2T<T>(UR, n2 * batch)); dVR.up(h2T<T>(VR, n2 * batch)); dUL.up(h2T<T>(UL, n2 * batch)); dVL.up(h2T<T>(VL, n2 * batch));
  { std::vector<double> a((size_t)n * batch), b((size_t)n * batch); for (size_t i = 0; i < a.size(); ++i) { a[i] = DR[2 * i]; b[i] = DL[2 * i]; } dDR.up(a); dDL.up(b); }
  UdvDev<T> R, L; R.U = dUR.p; R.V = dVR.p; R.D = dDR.p; R.det = detR.p; L.U = dUL.p; L.V = dVL.p; L.D = dDL.p; L.det = detL.p;
  if (detUR && detUL) { std::vector<cplx> a(batch), b(batch); for (int i = 0; i < batch; ++i) { a[i] = cplx(detUR[2 * i], detUR[2 * i + 1]); b[i] = cplx(detUL[2 * i], detUL[2 * i + 1]); } detR.up(a); detL.up(b); }
  else {      // stand-alone call as the reference's CGR(PHASE, NVAR, GRUP, udvr, udvl): det U_R, det U_L (unit modulus) from a pivoted QR of a copy on the device
    for (int q = 0; q < 2; ++q) {
      CK(cudaMemcpyAsync(w.W[0], q ? dUL.p : dUR.p, sizeof(T) * n2 * batch, cudaMemcpyDeviceToDevice, 0));
      la_qrp<T>(w, w.W[0], n, n, w.Dq);
      k_det_phase<<<(batch + 127) / 128, 128>>>(w.qrout, q ? detL.p : detR.p, batch);
    }
  }
  la_cgr<T>(w, nvar, stab, R, L, dG.p, dz.p); CK(cudaDeviceSynchronize());
  T2h<T>(dG.down(), G); auto z = dz.down(); for (int i = 0; i < batch; ++i) { phase[2 * i] = z[i].x; phase[2 * i + 1] = z[i].y; }
  w.release();
}
template <typename T>
static void t_cgrp(int n, int np, int batch, const double* UR, const double* UL, double* G, double* phase) {
  const size_t n2 = (size_t)n * n; LaWork<T> w, wp; w.alloc(n, batch, 0); wp.alloc(np, batch, 0);
  DevBuf<T> dUR(n2 * batch), dUL(n2 * batch), dG(n2 * batch); DevBuf<cplx> dz(batch);
  // inputs are n x np per matrix; the device layout keeps the leading dimension n inside an n x n slot
  std::vector<T> a(n2 * batch, zero_<T>()), b(n2 * batch, zero_<T>());
  for (int q = 0; q < batch; ++q) for (size_t i = 0; i < (size_t)n * np; ++i) {
    const size_t src = (size_t)q * n * np + i;
    a[(size_t)q * n2 + i] = to_T<T>(cd(UR[2 * src], UR[2 * src + 1])); b[(size_t)q * n2 + i] = to_T<T>(cd(UL[2 * src], UL[2 * src + 1]));
  }
  dUR.up(a); dUL.up(b);
  UdvDev<T> R, L; R.U = dUR.p; L.U = dUL.p;
  la_cgrp<T>(w, wp, R, L, dG.p, dz.p); CK(cudaDeviceSynchronize());
  T2h<T>(dG.down(), G); auto z = dz.down(); for (int i = 0; i < batch; ++i) { phase[2 * i] = z[i].x; phase[2 * i + 1] = z[i].y; }
  w.release(); wp.release();
}
template <typename T>
static void t_cgr22(int n, int batch, int stab, const double* U2, const double* D2, const double* V2, const double* U1, const double* D1, const double* V1, double* out4) {
  const size_t n2 = (size_t)n * n; LaWork<T> w, w2; w.alloc(n, batch, 0); w2.alloc(2 * n, batch, 0);
  DevBuf<T> dU2(n2 * batch), dV2(n2 * batch), dU1(n2 * batch), dV1(n2 * batch), g0(n2 * batch), g1(n2 * batch), g2(n2 * batch), g3(n2 * batch);
  DevBuf<double> dD2((size_t)n * batch), dD1((size_t)n * batch); DevBuf<int> first(batch);
  dU2.up(h2T<T>(U2, n2 * batch)); dV2.up(h2T<T>(V2, n2 * batch)); dU1.up(h2T<T>(U1, n2 * batch)); dV1.up(h2T<T>(V1, n2 * batch));
  { std::vector<double> a((size_t)n * batch), b((size_t)n * batch); for (size_t i = 0; i < a.size(); ++i) { a[i] = D2[2 * i]; b[i] = D1[2 * i]; } dD2.up(a); dD1.up(b); }
  UdvDev<T> u2, u1; u2.U = dU2.p; u2.V = dV2.p; u2.D = dD2.p; u1.U = dU1.p; u1.V = dV1.p; u1.D = dD1.p;
  la_cgr2_2<T>(w, w2, stab, u2, u1, g0.p, g1.p, g2.p, g3.p, first.p); CK(cudaDeviceSynchronize());
  DevBuf<T>* gs[4] = {&g0, &g1, &g2, &g3};
  for (int q = 0; q < 4; ++q) T2h<T>(gs[q]->down(), out4 + 2 * q * n2 * batch);
  w.release(); w2.release();
}
template <typename T>
static void t_gemm(int ta, int tb, int m, int n, int k, int batch, const double* A, const double* B, double* Cc) {
  const int ar = ta ? k : m, ac = ta ? m : k, br = tb ? n : k, bc = tb ? k : n;
  DevBuf<T> dA((size_t)ar * ac * batch), dB((size_t)br * bc * batch), dC((size_t)m * n * batch);
  dA.up(h2T<T>(A, dA.n)); dB.up(h2T<T>(B, dB.n));
  if (!ta && !tb) gemm<T, 0, 0, 0>(0, m, n, k, dA.p, ar, (long)ar * ac, dB.p, br, (long)br * bc, dC.p, m, (long)m * n, batch);
  else if (ta && !tb) gemm<T, 1, 0, 0>(0, m, n, k, dA.p, ar, (long)ar * ac, dB.p, br, (long)br * bc, dC.p, m, (long)m * n, batch);
  else if (!ta && tb) gemm<T, 0, 1, 0>(0, m, n, k, dA.p, ar, (long)ar * ac, dB.p, br, (long)br * bc, dC.p, m, (long)m * n, batch);
  else gemm<T, 1, 1, 0>(0, m, n, k, dA.p, ar, (long)ar * ac, dB.p, br, (long)br * bc, dC.p, m, (long)m * n, batch);
  CK(cudaDeviceSynchronize()); T2h<T>(dC.down(), Cc);
}
