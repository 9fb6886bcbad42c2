// alf_b200.cu -- the C-ABI (include/alf_b200.h).  The engine templates are instantiated in alf_inst_real.cu / alf_inst_cplx.cu.
// There is no CPU fallback: without a CUDA device every entry point fails with ALF_ERROR_CUDA.
#include <dlfcn.h>
#include <nccl.h>        // types only: the symbols are resolved with dlopen (see NcclApi)
#include "alf_engine.cuh"
#include "alf_inst.h"

thread_local Prof* t_prof = nullptr;

__global__ void k_peak_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double x = 1.0000001, y = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y); a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_peak_dmma(double* out, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}

// ================================================================================================ C-ABI
#define API_BEGIN(h) if (!(h)) return ALF_ERROR_GENERIC; try { CK(cudaSetDevice((h)->device)); t_prof = &(h)->prof;
#define API_END(h) } catch (const CudaError& e) { (h)->err = e.what(); return ALF_ERROR_CUDA; } \
  catch (const std::exception& e) { (h)->err = e.what(); return ALF_ERROR_GENERIC; } return ALF_OK;
#define API_END_NORET(h) } catch (const CudaError& e) { (h)->err = e.what(); return ALF_ERROR_CUDA; } \
  catch (const std::exception& e) { (h)->err = e.what(); return ALF_ERROR_GENERIC; }

extern "C" {

int alf_b200_create(alf_b200_handle** out, int ndim, int n_fl, int n_sun, int ltrot, int nwrap, int n_opv, int n_opt,
                    int symm, int stab, int n_chains, int device) {
  if (!out) return ALF_ERROR_GENERIC;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) return ALF_ERROR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ALF_ERROR_CUDA;
  if (prop.major < 10) return ALF_ERROR_CUDA;        // sm_100a code only
  if (ndim < 1 || n_fl < 1 || ltrot < 1 || nwrap < 1 || n_chains < 1 || n_opv < 0 || n_opt < 0) return ALF_ERROR_GENERIC;
  alf_b200_handle* h = new alf_b200_handle();
  h->ndim = ndim; h->n_fl = n_fl; h->n_sun = n_sun; h->ltrot = ltrot; h->nwrap = nwrap; h->n_opv = n_opv; h->n_opt = n_opt;
  h->symm = symm; h->stab = stab; h->n_chains = n_chains; h->device = device;
  h->opv.resize((size_t)n_opv * n_fl); h->opt.resize((size_t)n_opt * n_fl); h->types.assign(n_opv, 0);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&h->stream) != cudaSuccess) { delete h; return ALF_ERROR_CUDA; }
  *out = h;
  return ALF_OK;
}

int alf_b200_destroy(alf_b200_handle* h) {
  if (!h) return ALF_OK;
  cudaSetDevice(h->device);
  if (t_prof == &h->prof) t_prof = nullptr;          // the launch-accounting pointer must not outlive its handle
  alf_b200_comm_destroy(h); if (h->d_ctlred) cudaFree(h->d_ctlred);
  h->eng.reset();
  if (h->d_fields_c) cudaFree(h->d_fields_c);
  if (h->d_fields) cudaFree(h->d_fields); if (h->d_rng) cudaFree(h->d_rng); if (h->d_phase) cudaFree(h->d_phase);
  if (h->d_obse_acc) cudaFree(h->d_obse_acc); if (h->d_obse_bg) cudaFree(h->d_obse_bg); if (h->d_obse_cnt) cudaFree(h->d_obse_cnt);
  if (h->d_obst_acc) cudaFree(h->d_obst_acc); if (h->d_obst_bg) cudaFree(h->d_obst_bg); if (h->d_obst_cnt) cudaFree(h->d_obst_cnt);
  if (h->d_counters) cudaFree(h->d_counters); if (h->d_ctl) cudaFree(h->d_ctl); if (h->d_acclog) cudaFree(h->d_acclog); if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_kin_idx) cudaFree(h->d_kin_idx); if (h->d_pot_idx) cudaFree(h->d_pot_idx); if (h->d_kin_coef) cudaFree(h->d_kin_coef); if (h->d_pot_coef) cudaFree(h->d_pot_coef);
  if (h->pin_fields) cudaFreeHost(h->pin_fields);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ALF_OK;
}
const char* alf_b200_last_error(const alf_b200_handle* h) { return h ? h->err.c_str() : "null handle"; }

static int fill_host_op(HostOp& op, int N, int nnz, int diag, int type, const int* P, const double* U, const double* E, double gr, double gi, double ar, double ai, int ndim) {
  if (N < 1 || !P || !U || !E) return ALF_ERROR_HAMILTONIAN;
  if (type < 0 || type > 4) return ALF_ERROR_HAMILTONIAN;                    // Op_set validation, Operator_mod.F90:275-281
  op.N = N; op.nnz = nnz; op.diag = diag; op.type = type; op.P.resize(N); op.U.resize((size_t)N * N); op.E.resize(N);
  for (int i = 0; i < N; ++i) { if (P[i] < 1 || P[i] > ndim) return ALF_ERROR_HAMILTONIAN; op.P[i] = P[i] - 1; op.E[i] = E[i]; }
  for (size_t i = 0; i < (size_t)N * N; ++i) op.U[i] = cd(U[2 * i], U[2 * i + 1]);
  op.g = cd(gr, gi); op.alpha = cd(ar, ai); op.set = true;
  return ALF_OK;
}
int alf_b200_set_op_v(alf_b200_handle* h, int n, int nf, int N, int nnz, int diag, int type, const int* P, const double* U, const double* E,
                      double g_re, double g_im, double a_re, double a_im) {
  if (!h || h->finalized || n < 1 || n > h->n_opv || nf < 1 || nf > h->n_fl) return ALF_ERROR_HAMILTONIAN;
  h->types[n - 1] = type;
  return fill_host_op(h->opv[(n - 1) + (size_t)h->n_opv * (nf - 1)], N, nnz, diag, type, P, U, E, g_re, g_im, a_re, a_im, h->ndim);
}
int alf_b200_set_op_v_gt(alf_b200_handle* h, int n, int nf, const double* g_t) {
  if (!h || h->finalized || n < 1 || n > h->n_opv || nf < 1 || nf > h->n_fl || !g_t) return ALF_ERROR_HAMILTONIAN;
  HostOp& op = h->opv[(n - 1) + (size_t)h->n_opv * (nf - 1)];
  if (!op.set) { h->err = "alf_b200_set_op_v_gt: call alf_b200_set_op_v for this vertex first"; return ALF_ERROR_HAMILTONIAN; }
  op.g_t.resize(h->ltrot); for (int t = 0; t < h->ltrot; ++t) op.g_t[t] = cd(g_t[2 * t], g_t[2 * t + 1]);
  h->has_gt = true;
  return ALF_OK;
}
int alf_b200_set_op_t(alf_b200_handle* h, int nc, int nf, int N, int diag, const int* P, const double* U, const double* E, double g_re, double g_im) {
  if (!h || h->finalized || nc < 1 || nc > h->n_opt || nf < 1 || nf > h->n_fl) return ALF_ERROR_HAMILTONIAN;
  return fill_host_op(h->opt[(nc - 1) + (size_t)h->n_opt * (nf - 1)], N, N, diag, 0, P, U, E, g_re, g_im, 0, 0, h->ndim);
}

// projective algorithm: Thtrot and the number of particles per flavor, then one trial wave function pair per flavor
// (WF_L(nf)%P, WF_R(nf)%P: Ndim x N_part, column-major complex; Prog/main.F90:366-376, 596-599)
int alf_b200_set_s0_ising(alf_b200_handle* h, int n_terms, const int* op_start, const int* term_start, const int* entry_op, const int* entry_dt,
                          const double* w, int open_boundaries, int propose_s0) {
  if (!h) return ALF_ERROR_GENERIC;
  if (h->finalized) { h->err = "alf_b200_set_s0_ising must be called before alf_b200_finalize_model"; return ALF_ERROR_GENERIC; }
  h->propose_s0 = propose_s0 ? 1 : 0;
  if (n_terms <= 0) { h->s0_on = false; return ALF_OK; }
  if (!op_start || !term_start || !entry_op || !entry_dt || !w || op_start[0] != 0 || op_start[h->n_opv] != n_terms || term_start[0] != 0) { h->err = "alf_b200_set_s0_ising: inconsistent tables"; return ALF_ERROR_GENERIC; }
  const int ne = term_start[n_terms];
  for (int e = 0; e < ne; ++e) if (entry_op[e] < 1 || entry_op[e] > h->n_opv || entry_dt[e] <= -h->ltrot || entry_dt[e] >= h->ltrot) { h->err = "alf_b200_set_s0_ising: entry out of range"; return ALF_ERROR_GENERIC; }
  h->s0_on = true; h->s0_open_bc = open_boundaries ? 1 : 0;
  h->s0_op_start.assign(op_start, op_start + h->n_opv + 1); h->s0_term_start.assign(term_start, term_start + n_terms + 1);
  h->s0_e_op.resize(ne); for (int e = 0; e < ne; ++e) h->s0_e_op[e] = entry_op[e] - 1;
  h->s0_e_dt.assign(entry_dt, entry_dt + ne); h->s0_w.assign(w, w + 2 * (size_t)n_terms);
  return ALF_OK;
}
int alf_b200_set_global_tau_sampling(alf_b200_handle* h, int nt_sequential_start, int nt_sequential_end, int n_global_tau) {
  if (!h) return ALF_ERROR_GENERIC;
  if (h->finalized) { h->err = "alf_b200_set_global_tau_sampling must be called before alf_b200_finalize_model"; return ALF_ERROR_GENERIC; }
  if (nt_sequential_start < 1 || nt_sequential_end > h->n_opv || nt_sequential_start > nt_sequential_end + 1 || n_global_tau < 0) { h->err = "alf_b200_set_global_tau_sampling: illegal range"; return ALF_ERROR_GENERIC; }
  h->nt_seq_start = nt_sequential_start; h->nt_seq_end = nt_sequential_end; h->n_global_tau = n_global_tau;
  return ALF_OK;
}
int alf_b200_set_measure_interval(alf_b200_handle* h, int lobs_st, int lobs_en) {
  if (!h) return ALF_ERROR_GENERIC;
  if (lobs_st < 0 || lobs_en < 0 || lobs_st > h->ltrot || lobs_en > h->ltrot) { h->err = "alf_b200_set_measure_interval: 0 <= LOBS_ST, LOBS_EN <= Ltrot"; return ALF_ERROR_GENERIC; }
  if (h->projector) {       // set_default_values_measuring_interval, QMC_runtime_var_mod.F90:163-180
    if (lobs_st != 0 && lobs_st < h->thtrot + 1) { h->err = "Measuring out of dedicating interval, LOBS_ST too small."; return ALF_ERROR_GENERIC; }
    if (lobs_en != 0 && lobs_en > h->ltrot - h->thtrot) { h->err = "Measuring out of dedicating interval, LOBS_EN too big."; return ALF_ERROR_GENERIC; }
  }
  h->lobs_st = lobs_st; h->lobs_en = lobs_en;
  return ALF_OK;
}
int alf_b200_set_global_move_tau_ising(alf_b200_handle* h, int n_sites, const int* move_start, const int* move_fields, int n_terms, const int* site_term_start,
                                       const int* term_start, const int* entry_op, const int* entry_dt, const double* w, int open_boundaries) {
  if (!h) return ALF_ERROR_GENERIC;
  if (h->finalized) { h->err = "alf_b200_set_global_move_tau_ising must be called before alf_b200_finalize_model"; return ALF_ERROR_GENERIC; }
  if (n_sites < 1 || !move_start || !move_fields || n_terms < 0 || !site_term_start || !term_start || !w || move_start[0] != 0 || site_term_start[0] != 0 ||
      site_term_start[n_sites] != n_terms || term_start[0] != 0) { h->err = "alf_b200_set_global_move_tau_ising: inconsistent tables"; return ALF_ERROR_GENERIC; }
  for (int I = 0; I < n_sites; ++I) {
    const int len = move_start[I + 1] - move_start[I];
    if (len < 1 || len > ALF_GM_MAXLEN) { h->err = "alf_b200_set_global_move_tau_ising: a move flips between 1 and 16 fields"; return ALF_ERROR_GENERIC; }
    for (int e = move_start[I]; e < move_start[I + 1]; ++e) {
      if (move_fields[e] < 1 || move_fields[e] > h->n_opv || (e > move_start[I] && move_fields[e] <= move_fields[e - 1])) { h->err = "alf_b200_set_global_move_tau_ising: Flip_list must be ascending (Wrapgr_sort) and in range"; return ALF_ERROR_GENERIC; }
      if (h->opv[(size_t)(move_fields[e] - 1)].type != 1) { h->err = "alf_b200_set_global_move_tau_ising: only Ising (type 1) fields"; return ALF_ERROR_UNSUPPORTED; }
    }
  }
  const int ne = term_start[n_terms];
  for (int e = 0; e < ne; ++e) if (entry_op[e] < 1 || entry_op[e] > h->n_opv || entry_dt[e] <= -h->ltrot || entry_dt[e] >= h->ltrot) { h->err = "alf_b200_set_global_move_tau_ising: entry out of range"; return ALF_ERROR_GENERIC; }
  h->gmt_on = true; h->gmt_n_sites = n_sites; h->gmt_open_bc = open_boundaries ? 1 : 0;
  h->gmt_move_start.assign(move_start, move_start + n_sites + 1); h->gmt_move_fields.resize(move_start[n_sites]);
  for (int e = 0; e < move_start[n_sites]; ++e) h->gmt_move_fields[e] = move_fields[e] - 1;
  h->gmt_op_start.assign(site_term_start, site_term_start + n_sites + 1); h->gmt_term_start.assign(term_start, term_start + n_terms + 1);
  h->gmt_e_op.resize(ne); for (int e = 0; e < ne; ++e) h->gmt_e_op[e] = entry_op[e] - 1;
  h->gmt_e_dt.assign(entry_dt, entry_dt + ne); h->gmt_w.assign(w, w + 2 * (size_t)n_terms);
  return ALF_OK;
}
int alf_b200_set_s0_gaussian(alf_b200_handle* h, int on) { if (!h) return ALF_ERROR_GENERIC; h->s0_gaussian = on != 0; return ALF_OK; }
int alf_b200_set_amplitude(alf_b200_handle* h, double amplitude) { if (!h) return ALF_ERROR_GENERIC; h->amplitude = amplitude; return ALF_OK; }
int alf_b200_set_projector(alf_b200_handle* h, int thtrot, int n_part) {
  if (!h || h->finalized || thtrot < 0 || n_part < 1 || n_part > h->ndim) return ALF_ERROR_HAMILTONIAN;
  h->projector = true; h->thtrot = thtrot; h->n_part = n_part;
  h->wf_l.assign(h->n_fl, std::vector<cd>()); h->wf_r.assign(h->n_fl, std::vector<cd>());
  return ALF_OK;
}
int alf_b200_set_trial_wf(alf_b200_handle* h, int nf, const double* PL, const double* PR) {
  if (!h || h->finalized || !h->projector || nf < 1 || nf > h->n_fl || !PL || !PR) return ALF_ERROR_HAMILTONIAN;
  const size_t n = (size_t)h->ndim * h->n_part;
  h->wf_l[nf - 1].resize(n); h->wf_r[nf - 1].resize(n);
  for (size_t i = 0; i < n; ++i) { h->wf_l[nf - 1][i] = cd(PL[2 * i], PL[2 * i + 1]); h->wf_r[nf - 1][i] = cd(PR[2 * i], PR[2 * i + 1]); }
  return ALF_OK;
}

int alf_b200_finalize_model(alf_b200_handle* h) {
  API_BEGIN(h)
  for (auto& o : h->opv) if (!o.set) { h->err = "finalize_model: an Op_V entry was never set"; return ALF_ERROR_HAMILTONIAN; }
  for (auto& o : h->opt) if (!o.set) { h->err = "finalize_model: an Op_T entry was never set"; return ALF_ERROR_HAMILTONIAN; }
  bool has_cont = false;
  for (auto& o : h->opv) {
    if (o.type == 3) { has_cont = true; if (o.N != 1) { h->err = "continuous fields (type 3) are supported for single-site vertices only in this build"; return ALF_ERROR_UNSUPPORTED; } }
    else if (o.type != 1 && o.type != 2) { h->err = "field types 1, 2 (discrete) and 3 (continuous, real) are supported in this build; type 4 is not"; return ALF_ERROR_UNSUPPORTED; }
  }
  if (h->has_gt && (h->n_global_tau > 0 || h->s0_on)) { h->err = "time-dependent couplings g_t together with Ising action tables / global moves are not supported in this build"; return ALF_ERROR_UNSUPPORTED; }
  if (has_cont && (h->n_global_tau > 0 || h->s0_on)) { h->err = "continuous fields together with Ising action tables / global moves are not supported in this build"; return ALF_ERROR_UNSUPPORTED; }
  // real instantiation iff every table the sweep touches is real
  bool cplx_needed = false;
  for (auto& o : h->opt) { if (o.g.imag() != 0.0) cplx_needed = true; for (auto& u : o.U) if (u.imag() != 0.0) cplx_needed = true; }
  for (auto& o : h->opv) { if (o.g.imag() != 0.0 || (o.g * o.alpha).imag() != 0.0) cplx_needed = true; for (auto& u : o.U) if (u.imag() != 0.0) cplx_needed = true;
    for (auto& gt : o.g_t) if (gt.imag() != 0.0 || (gt * o.alpha).imag() != 0.0) cplx_needed = true; }
  if (h->projector) {
    for (int f = 0; f < h->n_fl; ++f) {
      if (h->wf_l[f].size() != (size_t)h->ndim * h->n_part || h->wf_r[f].size() != (size_t)h->ndim * h->n_part) {
        h->err = "Projector is selected but there are no trial wave functions!"; return ALF_ERROR_HAMILTONIAN; }      // main.F90:366-372
      for (auto& z : h->wf_l[f]) if (z.imag() != 0.0) cplx_needed = true;
      for (auto& z : h->wf_r[f]) if (z.imag() != 0.0) cplx_needed = true;
    }
  }
  h->is_complex = cplx_needed;
  const long C = h->n_chains;
  CK(cudaMalloc(&h->d_fields, (size_t)C * h->ltrot * std::max(1, h->n_opv))); CK(cudaMemset(h->d_fields, 1, (size_t)C * h->ltrot * std::max(1, h->n_opv)));
  if (has_cont) { const size_t nf = (size_t)C * h->ltrot * h->n_opv; std::vector<double> one(nf, 1.0); CK(cudaMalloc(&h->d_fields_c, sizeof(double) * nf)); CK(cudaMemcpy(h->d_fields_c, one.data(), sizeof(double) * nf, cudaMemcpyHostToDevice)); }
  CK(cudaMalloc(&h->d_rng, sizeof(uint64_t) * 4 * C)); CK(cudaMalloc(&h->d_phase, sizeof(cplx) * C));
  CK(cudaMalloc(&h->d_counters, sizeof(unsigned long long) * 5 * C)); CK(cudaMemset(h->d_counters, 0, sizeof(unsigned long long) * 5 * C));   // [chain][4] control counters, then [chain] flushes
  CK(cudaMalloc(&h->d_ctl, sizeof(double) * 8 * C)); CK(cudaMemset(h->d_ctl, 0, sizeof(double) * 8 * C));
  h->obs_size = 16; CK(cudaMalloc(&h->d_obs, sizeof(double) * h->obs_size)); CK(cudaMemset(h->d_obs, 0, sizeof(double) * h->obs_size));
  { std::vector<int32_t> z(C, 0); int32_t* d; CK(cudaMalloc(&d, sizeof(int32_t) * C)); CK(cudaMemcpy(d, z.data(), sizeof(int32_t) * C, cudaMemcpyHostToDevice));
    KL(KC_EW, h->stream, k_ranset<<<(int)((C + 127) / 128), 128, 0, h->stream>>>(h->d_rng, d, (int)C)); CK(cudaStreamSynchronize(h->stream)); cudaFree(d); }
  KL(KC_EW, h->stream, k_fill_cplx<<<ew_blocks(C), 256, 0, h->stream>>>(h->d_phase, C, cplx(1.0, 0.0)));
  if (h->is_complex) h->eng.reset(alf_make_engine_cplx(h)); else h->eng.reset(alf_make_engine_real(h));
  CK(cudaStreamSynchronize(h->stream));
  h->finalized = true;
  API_END(h)
}
int alf_b200_is_complex(const alf_b200_handle* h) { return h && h->is_complex ? 1 : 0; }

#define NEED_FINAL(h) if (!(h)->finalized) { (h)->err = "model not finalized"; return ALF_ERROR_GENERIC; }

int alf_b200_set_seeds(alf_b200_handle* h, const int32_t* seeds) {
  API_BEGIN(h) NEED_FINAL(h)
  int32_t* d; CK(cudaMalloc(&d, sizeof(int32_t) * h->n_chains)); CK(cudaMemcpy(d, seeds, sizeof(int32_t) * h->n_chains, cudaMemcpyHostToDevice));
  KL(KC_EW, h->stream, k_ranset<<<(h->n_chains + 127) / 128, 128, 0, h->stream>>>(h->d_rng, d, h->n_chains)); CK(cudaStreamSynchronize(h->stream)); cudaFree(d);
  API_END(h)
}
int alf_b200_get_rng_state(alf_b200_handle* h, uint64_t* s) { API_BEGIN(h) NEED_FINAL(h) CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(s, h->d_rng, sizeof(uint64_t) * 4 * h->n_chains, cudaMemcpyDeviceToHost)); API_END(h) }
int alf_b200_set_rng_state(alf_b200_handle* h, const uint64_t* s) { API_BEGIN(h) NEED_FINAL(h) CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(h->d_rng, s, sizeof(uint64_t) * 4 * h->n_chains, cudaMemcpyHostToDevice)); API_END(h) }
int alf_b200_fields_set(alf_b200_handle* h) {
  API_BEGIN(h) NEED_FINAL(h)
  KL(KC_EW, h->stream, k_fields_set<<<(h->n_chains + 63) / 64, 64, 0, h->stream>>>(h->d_fields, h->d_rng, h->n_chains, (long)h->ltrot * h->n_opv));
  if (h->d_fields_c) { const long nf = (long)h->n_chains * h->ltrot * h->n_opv; KL(KC_EW, h->stream, k_i8_to_f64<<<ew_blocks(nf), 256, 0, h->stream>>>(h->d_fields, h->d_fields_c, nf)); }   // type 3 starts from +-1 too (Fields_mod.F90:600-602)
  API_END(h)
}
int alf_b200_set_fields(alf_b200_handle* h, const double* f) {
  API_BEGIN(h) NEED_FINAL(h)
  const size_t n = (size_t)h->n_chains * h->ltrot * h->n_opv; std::vector<int8_t> b(n);
  std::vector<double> bc(h->d_fields_c ? n : 0);
  for (size_t i = 0; i < n; ++i) { int t = h->types[i % h->n_opv];
    if (t == 3) { b[i] = 1; bc[i] = f[2 * i]; continue; }
    long s = std::lround(f[2 * i]);
    if (s == 0 || std::labs(s) > t) { h->err = "set_fields: field value outside the discrete range of its operator type"; return ALF_ERROR_FIELDS; } b[i] = (int8_t)s; if (h->d_fields_c) bc[i] = (double)s; }
  CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(h->d_fields, b.data(), n, cudaMemcpyHostToDevice));
  if (h->d_fields_c) CK(cudaMemcpy(h->d_fields_c, bc.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  API_END(h)
}
int alf_b200_get_fields(alf_b200_handle* h, double* f) {
  API_BEGIN(h) NEED_FINAL(h)
  const size_t n = (size_t)h->n_chains * h->ltrot * h->n_opv; std::vector<int8_t> b(n);
  CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(b.data(), h->d_fields, n, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) { f[2 * i] = (double)b[i]; f[2 * i + 1] = 0.0; }
  if (h->d_fields_c) { std::vector<double> bc(n); CK(cudaMemcpy(bc.data(), h->d_fields_c, sizeof(double) * n, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < n; ++i) if (h->types[i % h->n_opv] == 3) f[2 * i] = bc[i]; }
  API_END(h)
}

int alf_b200_init_sweep(alf_b200_handle* h) { API_BEGIN(h) NEED_FINAL(h) h->eng->init_sweep(); h->eng->sync(); API_END(h) }
int alf_b200_sweep(alf_b200_handle* h, int n_sweeps, int ltau) {
  API_BEGIN(h) NEED_FINAL(h)
  for (int s = 0; s < n_sweeps; ++s) h->eng->sweep(ltau);
  h->eng->sync();
  API_END(h)
}
int alf_b200_get_control(alf_b200_handle* h, double* out);
int alf_b200_get_obs(alf_b200_handle* h, double* out);
int alf_b200_sweep_host(alf_b200_handle* h, int n_sweeps, int ltau, const double* fields_in, double* fields_out, double* obs_out, double* control_out) {
  {
  API_BEGIN(h) NEED_FINAL(h)
  // fields cross PCIe as one int8 per (chain, nt, n) through a pinned staging buffer (nsigma%f of discrete fields holds +-1, +-2 only)
  const size_t n = (size_t)h->n_chains * h->ltrot * h->n_opv;
  if (!h->pin_fields) CK(cudaHostAlloc((void**)&h->pin_fields, n ? n : 1, cudaHostAllocDefault));
  if (fields_in) {
    std::vector<double> bc(h->d_fields_c ? n : 0);
    for (size_t i = 0; i < n; ++i) { int t = h->types[i % h->n_opv];
      if (t == 3) { h->pin_fields[i] = 1; bc[i] = fields_in[2 * i]; continue; }
      long s = std::lround(fields_in[2 * i]);
      if (s == 0 || std::labs(s) > t) { h->err = "sweep_host: field value outside the discrete range of its operator type"; return ALF_ERROR_FIELDS; } h->pin_fields[i] = (int8_t)s; if (h->d_fields_c) bc[i] = (double)s; }
    CK(cudaMemcpyAsync(h->d_fields, h->pin_fields, n, cudaMemcpyHostToDevice, h->stream));
    if (h->d_fields_c) { CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(h->d_fields_c, bc.data(), sizeof(double) * n, cudaMemcpyHostToDevice)); }     // continuous fields: 8 bytes each
  }
  for (int s = 0; s < n_sweeps; ++s) h->eng->sweep(ltau);
  if (fields_out) {
    CK(cudaMemcpyAsync(h->pin_fields, h->d_fields, n, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) { fields_out[2 * i] = (double)h->pin_fields[i]; fields_out[2 * i + 1] = 0.0; }
    if (h->d_fields_c) { std::vector<double> bc(n); CK(cudaMemcpy(bc.data(), h->d_fields_c, sizeof(double) * n, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < n; ++i) if (h->types[i % h->n_opv] == 3) fields_out[2 * i] = bc[i]; }
  } else h->eng->sync();
  API_END_NORET(h)
  }
  int rc;
  if (obs_out && (rc = alf_b200_get_obs(h, obs_out)) != ALF_OK) return rc;
  if (control_out && (rc = alf_b200_get_control(h, control_out)) != ALF_OK) return rc;
  return ALF_OK;
}

int alf_b200_wrapgrup(alf_b200_handle* h, int ntau) { API_BEGIN(h) NEED_FINAL(h) if (ntau < 0 || ntau > h->ltrot - 1) return ALF_ERROR_GENERIC; h->eng->wrapgrup(ntau); h->eng->sync(); API_END(h) }
int alf_b200_wrapgrdo(alf_b200_handle* h, int ntau) { API_BEGIN(h) NEED_FINAL(h) if (ntau < 1 || ntau > h->ltrot) return ALF_ERROR_GENERIC; h->eng->wrapgrdo(ntau); h->eng->sync(); API_END(h) }
int alf_b200_wrapur(alf_b200_handle* h, int ntau, int ntau1) { API_BEGIN(h) NEED_FINAL(h) h->eng->wrapur(ntau, ntau1); h->eng->sync(); API_END(h) }
int alf_b200_wrapul(alf_b200_handle* h, int ntau1, int ntau) { API_BEGIN(h) NEED_FINAL(h) h->eng->wrapul(ntau1, ntau); h->eng->sync(); API_END(h) }
int alf_b200_udv_reset(alf_b200_handle* h, int which, char side) { API_BEGIN(h) NEED_FINAL(h) h->eng->udv_reset(which, side); h->eng->sync(); API_END(h) }
int alf_b200_cgr(alf_b200_handle* h, int nvar) { API_BEGIN(h) NEED_FINAL(h) h->eng->cgr_call(nvar); h->eng->sync(); API_END(h) }
int alf_b200_langevin_forces(alf_b200_handle* h, double* forces) {
  API_BEGIN(h) NEED_FINAL(h)
  if (h->has_gt) { h->err = "time-dependent couplings g_t are not supported by the Langevin / HMC updates in this build"; return ALF_ERROR_UNSUPPORTED; }
  if (!forces) return ALF_ERROR_GENERIC;
  std::vector<cd> f((size_t)h->n_chains * h->ltrot * h->n_opv);
  h->eng->langevin_get_forces(f.data());
  for (size_t i = 0; i < f.size(); ++i) { forces[2 * i] = f[i].real(); forces[2 * i + 1] = f[i].imag(); }
  API_END(h)
}
int alf_b200_langevin_update(alf_b200_handle* h, double delta_t, double max_force, double* delta_t_running) {
  API_BEGIN(h) NEED_FINAL(h)
  if (!(delta_t > 0.0) || !(max_force > 0.0)) return ALF_ERROR_GENERIC;
  if (h->has_gt) { h->err = "time-dependent couplings g_t are not supported by the Langevin / HMC updates in this build"; return ALF_ERROR_UNSUPPORTED; }
  h->eng->langevin_update(delta_t, max_force, delta_t_running);
  API_END(h)
}
int alf_b200_hmc_update(alf_b200_handle* h, double delta_t, int leapfrog_steps, double* weight, uint8_t* accepted) {
  API_BEGIN(h) NEED_FINAL(h)
  if (!(delta_t > 0.0) || leapfrog_steps < 1) return ALF_ERROR_GENERIC;
  if (h->has_gt) { h->err = "time-dependent couplings g_t are not supported by the Langevin / HMC updates in this build"; return ALF_ERROR_UNSUPPORTED; }
  h->eng->hmc_update(delta_t, leapfrog_steps, weight, accepted);
  API_END(h)
}
int alf_b200_compute_fermion_det(alf_b200_handle* h, double* log_abs_det, double* phase_det) {
  API_BEGIN(h) NEED_FINAL(h)
  if (!log_abs_det || !phase_det) return ALF_ERROR_GENERIC;
  std::vector<cd> ph((size_t)h->n_chains * h->n_fl);
  h->eng->fermion_det(log_abs_det, ph.data());
  for (size_t i = 0; i < ph.size(); ++i) { phase_det[2 * i] = ph[i].real(); phase_det[2 * i + 1] = ph[i].imag(); }
  API_END(h)
}
int alf_b200_tau_m(alf_b200_handle* h) { API_BEGIN(h) NEED_FINAL(h) h->eng->tau_m(); h->eng->sync(); API_END(h) }
int alf_b200_tau_p(alf_b200_handle* h, int nst_in) { API_BEGIN(h) NEED_FINAL(h) if (nst_in < 0) return ALF_ERROR_GENERIC; h->eng->tau_p(nst_in); h->eng->sync(); API_END(h) }

// ---- lattice tables and device-side time-displaced lattice observables (what ham%ObserT accumulates through Predefined_Obs_tau_*)
int alf_b200_set_lattice(alf_b200_handle* h, int n_unit, int norb, const int* site_cell, const int* site_orb, const int* imj) {
  if (!h || n_unit < 1 || norb < 1 || !site_cell || !site_orb || !imj) return ALF_ERROR_GENERIC;
  h->n_unit = n_unit; h->norb = norb; h->site_cell.resize(h->ndim); h->site_orb.resize(h->ndim); h->imj.resize((size_t)n_unit * n_unit);
  for (int i = 0; i < h->ndim; ++i) {
    if (site_cell[i] > n_unit || (site_cell[i] > 0 && (site_orb[i] < 1 || site_orb[i] > norb))) { h->err = "set_lattice: List entry out of range"; return ALF_ERROR_GENERIC; }
    h->site_cell[i] = site_cell[i] - 1; h->site_orb[i] = site_orb[i] - 1;      // List(I1,1) <= 0: site not in the list (skipped, Predefined_Obs_mod.F90:365)
  }
  for (size_t i = 0; i < h->imj.size(); ++i) { if (imj[i] < 1 || imj[i] > n_unit) { h->err = "set_lattice: imj entry out of range"; return ALF_ERROR_GENERIC; } h->imj[i] = imj[i] - 1; }
  return ALF_OK;
}
int alf_b200_set_obs_scal_tables(alf_b200_handle* h, int n_kin, const int* kin_i, const int* kin_j, const int* kin_nf, const double* kin_coef,
                                 int n_pot, const int* pot_i1, const int* pot_nf1, const int* pot_i2, const int* pot_nf2, const double* pot_coef) {
  API_BEGIN(h) NEED_FINAL(h)
  if (n_kin < 0 || n_pot < 0) return ALF_ERROR_GENERIC;
  CK(cudaStreamSynchronize(h->stream));
  for (void* p : {(void*)h->d_kin_idx, (void*)h->d_pot_idx, (void*)h->d_kin_coef, (void*)h->d_pot_coef}) if (p) cudaFree(p);
  h->d_kin_idx = h->d_pot_idx = nullptr; h->d_kin_coef = h->d_pot_coef = nullptr; h->n_kin = h->n_pot = 0;
  auto bad = [&](int i) { return i < 1 || i > h->ndim; }; auto badf = [&](int f) { return f < 1 || f > h->n_fl; };
  std::vector<int> ki((size_t)3 * n_kin), pi((size_t)4 * n_pot); std::vector<cplx> kc(n_kin), pc(n_pot);
  for (int t = 0; t < n_kin; ++t) { if (bad(kin_i[t]) || bad(kin_j[t]) || badf(kin_nf[t])) { h->err = "set_obs_scal_tables: Kin index out of range"; return ALF_ERROR_GENERIC; }
    ki[3 * t] = kin_i[t] - 1; ki[3 * t + 1] = kin_j[t] - 1; ki[3 * t + 2] = kin_nf[t] - 1; kc[t] = cplx(kin_coef[2 * t], kin_coef[2 * t + 1]); }
  for (int t = 0; t < n_pot; ++t) { if (bad(pot_i1[t]) || bad(pot_i2[t]) || badf(pot_nf1[t]) || badf(pot_nf2[t])) { h->err = "set_obs_scal_tables: Pot index out of range"; return ALF_ERROR_GENERIC; }
    pi[4 * t] = pot_i1[t] - 1; pi[4 * t + 1] = pot_nf1[t] - 1; pi[4 * t + 2] = pot_i2[t] - 1; pi[4 * t + 3] = pot_nf2[t] - 1; pc[t] = cplx(pot_coef[2 * t], pot_coef[2 * t + 1]); }
  if (n_kin) { CK(cudaMalloc(&h->d_kin_idx, sizeof(int) * ki.size())); CK(cudaMalloc(&h->d_kin_coef, sizeof(cplx) * n_kin));
    CK(cudaMemcpy(h->d_kin_idx, ki.data(), sizeof(int) * ki.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(h->d_kin_coef, kc.data(), sizeof(cplx) * n_kin, cudaMemcpyHostToDevice)); }
  if (n_pot) { CK(cudaMalloc(&h->d_pot_idx, sizeof(int) * pi.size())); CK(cudaMalloc(&h->d_pot_coef, sizeof(cplx) * n_pot));
    CK(cudaMemcpy(h->d_pot_idx, pi.data(), sizeof(int) * pi.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(h->d_pot_coef, pc.data(), sizeof(cplx) * n_pot, cudaMemcpyHostToDevice)); }
  h->n_kin = n_kin; h->n_pot = n_pot;
  API_END(h)
}
static size_t obst_acc_len(const alf_b200_handle* h) { return (size_t)2 * OBST_NCH * h->obst_ntau * h->norb * h->norb * h->n_unit; }
static size_t obst_bg_len(const alf_b200_handle* h) { return (size_t)2 * 2 * h->obst_ntau * h->norb; }
int alf_b200_obs_tau_enable(alf_b200_handle* h, int on) {
  API_BEGIN(h) NEED_FINAL(h)
  if (on && !h->d_obst_acc) {
    if (h->n_unit <= 0) { h->err = "obs_tau_enable: call alf_b200_set_lattice first"; return ALF_ERROR_GENERIC; }
    if (h->n_fl > 2) { h->err = "obs_tau_enable: the device-side lattice observables support N_FL <= 2 (keep ham%ObserT on the host for more flavors)"; return ALF_ERROR_UNSUPPORTED; }
    h->obst_ntau = h->projector ? h->ltrot - 2 * h->thtrot + 1 : h->ltrot + 1;          // Ltau + 1 time points (Hubbard_smod.F90:628,666)
    CK(cudaMalloc(&h->d_obst_acc, sizeof(double) * obst_acc_len(h))); CK(cudaMalloc(&h->d_obst_bg, sizeof(double) * obst_bg_len(h))); CK(cudaMalloc(&h->d_obst_cnt, sizeof(double) * 2));
    CK(cudaMemsetAsync(h->d_obst_acc, 0, sizeof(double) * obst_acc_len(h), h->stream)); CK(cudaMemsetAsync(h->d_obst_bg, 0, sizeof(double) * obst_bg_len(h), h->stream));
    CK(cudaMemsetAsync(h->d_obst_cnt, 0, sizeof(double) * 2, h->stream));
    h->eng->obs_tau_setup();
  }
  h->obs_tau_on = on != 0;
  API_END(h)
}
// equal-time lattice observables (Predefined_Obs_eq_*), accumulated at every measured slice of the sweep
int alf_b200_obs_eq_enable(alf_b200_handle* h, int on) {
  API_BEGIN(h) NEED_FINAL(h)
  if (on && !h->d_obse_acc) {
    if (h->n_unit <= 0) { h->err = "obs_eq_enable: call alf_b200_set_lattice first"; return ALF_ERROR_GENERIC; }
    if (h->n_fl > 2) { h->err = "obs_eq_enable: the device-side lattice observables support N_FL <= 2 (keep ham%Obser on the host for more flavors)"; return ALF_ERROR_UNSUPPORTED; }
    const size_t na = (size_t)2 * OBST_NCH * h->norb * h->norb * h->n_unit, nb = (size_t)2 * 2 * h->norb;
    CK(cudaMalloc(&h->d_obse_acc, sizeof(double) * na)); CK(cudaMalloc(&h->d_obse_bg, sizeof(double) * nb)); CK(cudaMalloc(&h->d_obse_cnt, sizeof(double) * 2));
    CK(cudaMemsetAsync(h->d_obse_acc, 0, sizeof(double) * na, h->stream)); CK(cudaMemsetAsync(h->d_obse_bg, 0, sizeof(double) * nb, h->stream));
    CK(cudaMemsetAsync(h->d_obse_cnt, 0, sizeof(double) * 2, h->stream));
    h->eng->obs_tau_setup();
  }
  h->obs_eq_on = on != 0;
  API_END(h)
}
int alf_b200_get_obs_eq(alf_b200_handle* h, double* acc, double* bg, double* cnt) {
  API_BEGIN(h) NEED_FINAL(h) if (!h->d_obse_acc) return ALF_ERROR_GENERIC;
  const size_t na = (size_t)2 * OBST_NCH * h->norb * h->norb * h->n_unit, nb = (size_t)2 * 2 * h->norb;
  CK(cudaStreamSynchronize(h->stream));
  if (acc) CK(cudaMemcpy(acc, h->d_obse_acc, sizeof(double) * na, cudaMemcpyDeviceToHost));
  if (bg) CK(cudaMemcpy(bg, h->d_obse_bg, sizeof(double) * nb, cudaMemcpyDeviceToHost));
  if (cnt) CK(cudaMemcpy(cnt, h->d_obse_cnt, sizeof(double) * 2, cudaMemcpyDeviceToHost));
  API_END(h)
}
int alf_b200_obs_tau_reset(alf_b200_handle* h) {
  API_BEGIN(h) NEED_FINAL(h) if (!h->d_obst_acc) return ALF_ERROR_GENERIC;
  CK(cudaMemsetAsync(h->d_obst_acc, 0, sizeof(double) * obst_acc_len(h), h->stream)); CK(cudaMemsetAsync(h->d_obst_bg, 0, sizeof(double) * obst_bg_len(h), h->stream));
  CK(cudaMemsetAsync(h->d_obst_cnt, 0, sizeof(double) * 2, h->stream));
  API_END(h)
}
int alf_b200_obs_tau_dims(const alf_b200_handle* h, int* n_channels, int* ntau, int* norb, int* n_unit) {
  if (!h || !h->d_obst_acc) return ALF_ERROR_GENERIC;
  if (n_channels) *n_channels = OBST_NCH; if (ntau) *ntau = h->obst_ntau; if (norb) *norb = h->norb; if (n_unit) *n_unit = h->n_unit;
  return ALF_OK;
}
int alf_b200_get_obs_tau(alf_b200_handle* h, double* acc, double* bg, double* cnt) {
  API_BEGIN(h) NEED_FINAL(h) if (!h->d_obst_acc) return ALF_ERROR_GENERIC;
  CK(cudaStreamSynchronize(h->stream));
  if (acc) CK(cudaMemcpy(acc, h->d_obst_acc, sizeof(double) * obst_acc_len(h), cudaMemcpyDeviceToHost));
  if (bg) CK(cudaMemcpy(bg, h->d_obst_bg, sizeof(double) * obst_bg_len(h), cudaMemcpyDeviceToHost));
  if (cnt) CK(cudaMemcpy(cnt, h->d_obst_cnt, sizeof(double) * 2, cudaMemcpyDeviceToHost));
  API_END(h)
}

// ---- global-in-slice moves: Wrapgr_PlaceGR / Wrapgr_Random_update (Prog/Wrapgr_mod.F90:247-433) with host-supplied proposals
int alf_b200_wrapgr_set_position(alf_b200_handle* h, int m) { API_BEGIN(h) NEED_FINAL(h) if (m < 0 || m > h->n_opv) return ALF_ERROR_GENERIC; h->eng->gm_set_position(m); API_END(h) }
int alf_b200_wrapgr_get_position(alf_b200_handle* h, int* m) { API_BEGIN(h) NEED_FINAL(h) h->eng->gm_get_position(m); API_END(h) }
int alf_b200_wrapgr_placegr(alf_b200_handle* h, int m1, int ntau) {
  API_BEGIN(h) NEED_FINAL(h) if (m1 < 0 || m1 > h->n_opv || ntau < 1 || ntau > h->ltrot) return ALF_ERROR_GENERIC;
  h->eng->gm_random_update(ntau, 0, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, m1); API_END(h)
}
int alf_b200_wrapgr_random_update(alf_b200_handle* h, int ntau, int n_moves, int maxlen, const int* flip_length, const int* flip_list, const double* flip_value,
                                  const double* t0_ratio, const double* s0_ratio, uint8_t* accepted, int place_to) {
  API_BEGIN(h) NEED_FINAL(h)
  if (ntau < 1 || ntau > h->ltrot || n_moves < 0 || maxlen < 1 || maxlen > ALF_GM_MAXLEN || place_to > h->n_opv) return ALF_ERROR_GENERIC;
  const size_t np = (size_t)h->n_chains * n_moves;
  std::vector<int> l0(np * maxlen, 0); std::vector<int8_t> v8(np * maxlen, 1);
  for (size_t p = 0; p < np; ++p) {
    const int len = flip_length[p];
    if (len < 0 || len > maxlen) { h->err = "wrapgr_random_update: Flip_length out of range"; return ALF_ERROR_GENERIC; }
    std::vector<std::pair<int, long>> e(len);
    for (int c = 0; c < len; ++c) {
      const int n = flip_list[p * maxlen + c]; const long s = std::lround(flip_value[2 * (p * maxlen + c)]);
      if (n < 1 || n > h->n_opv) { h->err = "wrapgr_random_update: Flip_list entry outside 1..size(Op_V,1)"; return ALF_ERROR_GENERIC; }
      if (s == 0 || std::labs(s) > h->types[n - 1]) { h->err = "wrapgr_random_update: Flip_value outside the discrete range of its operator type"; return ALF_ERROR_FIELDS; }
      e[c] = std::make_pair(n - 1, s);
    }
    std::stable_sort(e.begin(), e.end(), [](const std::pair<int, long>& a, const std::pair<int, long>& b) { return a.first < b.first; });   // Wrapgr_sort
    for (int c = 0; c < len; ++c) { l0[p * maxlen + c] = e[c].first; v8[p * maxlen + c] = (int8_t)e[c].second; }
  }
  h->eng->gm_random_update(ntau, n_moves, maxlen, flip_length, l0.data(), v8.data(), t0_ratio, s0_ratio, accepted, place_to);
  API_END(h)
}

int alf_b200_get_green(alf_b200_handle* h, int chain, int nf, int symmetrize, double* out) {
  API_BEGIN(h) NEED_FINAL(h) if (chain < 0 || chain >= h->n_chains || nf < 1 || nf > h->n_fl) return ALF_ERROR_GENERIC;
  h->eng->get_green(chain, nf, symmetrize, reinterpret_cast<cd*>(out)); API_END(h)
}
int alf_b200_set_green(alf_b200_handle* h, int chain, int nf, const double* in) {
  API_BEGIN(h) NEED_FINAL(h) if (chain < 0 || chain >= h->n_chains || nf < 1 || nf > h->n_fl) return ALF_ERROR_GENERIC;
  h->eng->set_green(chain, nf, reinterpret_cast<const cd*>(in)); API_END(h)
}
int alf_b200_get_phase(alf_b200_handle* h, double* out) { API_BEGIN(h) NEED_FINAL(h) CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(out, h->d_phase, sizeof(cplx) * h->n_chains, cudaMemcpyDeviceToHost)); API_END(h) }
int alf_b200_get_udv(alf_b200_handle* h, int which, int nst, int chain, int nf, double* U, double* D, double* V) {
  API_BEGIN(h) NEED_FINAL(h) h->eng->get_udv(which, nst, chain, nf, reinterpret_cast<cd*>(U), reinterpret_cast<cd*>(D), reinterpret_cast<cd*>(V)); API_END(h)
}
int alf_b200_set_udv(alf_b200_handle* h, int which, int nst, int chain, int nf, const double* U, const double* D, const double* V) {
  API_BEGIN(h) NEED_FINAL(h)
  if (chain < 0 || chain >= h->n_chains || nf < 1 || nf > h->n_fl || which < 0 || which > 2 || !U || !D) return ALF_ERROR_GENERIC;
  h->eng->set_udv(which, nst, chain, nf, reinterpret_cast<const cd*>(U), reinterpret_cast<const cd*>(D), reinterpret_cast<const cd*>(V)); API_END(h)
}
int alf_b200_get_control(alf_b200_handle* h, double* out) {
  API_BEGIN(h) NEED_FINAL(h)
  const int C = h->n_chains; std::vector<double> c((size_t)C * 8); std::vector<unsigned long long> k((size_t)C * 5);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(c.data(), h->d_ctl, sizeof(double) * 8 * C, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(k.data(), h->d_counters, sizeof(unsigned long long) * 5 * C, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 16; ++i) out[i] = 0.0;
  for (int ch = 0; ch < C; ++ch) {   // the reductions of Control_Print (control_mod.F90:397-452): SUM for counters/means, MAX for maxima
    out[0] += c[ch * 8 + 0]; out[1] = std::max(out[1], c[ch * 8 + 1]); out[2] += c[ch * 8 + 2]; out[3] = std::max(out[3], c[ch * 8 + 3]);
    out[4] += c[ch * 8 + 4]; out[5] = std::max(out[5], c[ch * 8 + 5]); out[6] += c[ch * 8 + 6];
    out[7] += (double)k[ch * 4 + 0]; out[8] += (double)k[ch * 4 + 1]; out[9] += (double)k[ch * 4 + 2]; out[10] += (double)k[ch * 4 + 3];
    int fl = (int)c[ch * 8 + 7]; if (fl & 1) out[11] = 1.0; if (fl & 2) out[12] = 1.0;
    out[13] += (double)k[(size_t)4 * C + ch];
  }
  API_END(h)
}
// ---- NCCL bin / control reduction (symbols resolved at run time: the library loads without NCCL on a single-GPU host)
namespace {
struct NcclApi {
  void* lib = nullptr; bool tried = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr; ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (tried) return lib != nullptr;
    tried = true;
    for (const char* nm : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
#define NCCL_SYM(F) *(void**)(&F) = dlsym(lib, "nccl" #F); if (!F) { lib = nullptr; return false; }
    NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(Reduce) NCCL_SYM(GroupStart) NCCL_SYM(GroupEnd) NCCL_SYM(GetErrorString)
#undef NCCL_SYM
    return true;
  }
};
NcclApi g_nccl;
#define NCK(h, x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { (h)->err = std::string("NCCL: ") + g_nccl.GetErrorString(r_); return ALF_ERROR_CUDA; } } while (0)
}
int alf_b200_comm_unique_id(char* id) {
  if (!id || !g_nccl.load()) return ALF_ERROR_CUDA;
  ncclUniqueId u; if (g_nccl.GetUniqueId(&u) != ncclSuccess) return ALF_ERROR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size"); std::memcpy(id, &u, 128); return ALF_OK;
}
int alf_b200_comm_init(alf_b200_handle* h, int nranks, int rank, const char* id) {
  API_BEGIN(h) NEED_FINAL(h)
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) return ALF_ERROR_GENERIC;
  if (!g_nccl.load()) { h->err = "comm_init: libnccl.so.2 not found"; return ALF_ERROR_CUDA; }
  if (h->comm) { g_nccl.CommDestroy((ncclComm_t)h->comm); h->comm = nullptr; }
  ncclUniqueId u; std::memcpy(&u, id, 128); ncclComm_t c = nullptr;
  NCK(h, g_nccl.CommInitRank(&c, nranks, u, rank));
  h->comm = c; h->comm_nranks = nranks; h->comm_rank = rank;
  if (!h->d_ctlred) CK(cudaMalloc(&h->d_ctlred, sizeof(double) * 32));
  API_END(h)
}
int alf_b200_comm_destroy(alf_b200_handle* h) {
  if (!h) return ALF_OK;
  if (h->comm && g_nccl.lib) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); g_nccl.CommDestroy((ncclComm_t)h->comm); }
  h->comm = nullptr; h->comm_nranks = 1; h->comm_rank = 0; return ALF_OK;
}
int alf_b200_reduce_bins(alf_b200_handle* h, int root) {
  API_BEGIN(h) NEED_FINAL(h)
  if (!h->comm || h->comm_nranks <= 1) return ALF_OK;
  if (root < 0 || root >= h->comm_nranks) return ALF_ERROR_GENERIC;
  ncclComm_t c = (ncclComm_t)h->comm;
  struct Buf { double* p; size_t n; };
  std::vector<Buf> bufs; bufs.push_back({h->d_obs, (size_t)h->obs_size});
  if (h->d_obst_acc) { bufs.push_back({h->d_obst_acc, obst_acc_len(h)}); bufs.push_back({h->d_obst_bg, obst_bg_len(h)}); bufs.push_back({h->d_obst_cnt, 2}); }
  if (h->d_obse_acc) { bufs.push_back({h->d_obse_acc, (size_t)2 * OBST_NCH * h->norb * h->norb * h->n_unit}); bufs.push_back({h->d_obse_bg, (size_t)2 * 2 * h->norb}); bufs.push_back({h->d_obse_cnt, 2}); }
  NCK(h, g_nccl.GroupStart());
  for (const Buf& b : bufs) NCK(h, g_nccl.Reduce(b.p, b.p, b.n, ncclDouble, ncclSum, root, c, h->stream));
  NCK(h, g_nccl.GroupEnd());
  CK(cudaStreamSynchronize(h->stream));
  h->n_reduce_calls++;
  API_END(h)
}
int alf_b200_reduce_control(alf_b200_handle* h, int root, double* out) {
  if (!h || !out) return ALF_ERROR_GENERIC;
  int rc = alf_b200_get_control(h, out); if (rc != ALF_OK) return rc;
  API_BEGIN(h)
  if (!h->comm || h->comm_nranks <= 1) return ALF_OK;
  ncclComm_t c = (ncclComm_t)h->comm;
  CK(cudaMemcpyAsync(h->d_ctlred, out, sizeof(double) * 16, cudaMemcpyHostToDevice, h->stream)); CK(cudaMemcpyAsync(h->d_ctlred + 16, out, sizeof(double) * 16, cudaMemcpyHostToDevice, h->stream));
  NCK(h, g_nccl.GroupStart());
  NCK(h, g_nccl.Reduce(h->d_ctlred, h->d_ctlred, 16, ncclDouble, ncclSum, root, c, h->stream));
  NCK(h, g_nccl.Reduce(h->d_ctlred + 16, h->d_ctlred + 16, 16, ncclDouble, ncclMax, root, c, h->stream));
  NCK(h, g_nccl.GroupEnd());
  double r[32]; CK(cudaMemcpyAsync(r, h->d_ctlred, sizeof(double) * 32, cudaMemcpyDeviceToHost, h->stream)); CK(cudaStreamSynchronize(h->stream));
  if (h->comm_rank == root) { for (int i = 0; i < 16; ++i) out[i] = r[i]; for (int i : {1, 3, 5, 11, 12}) out[i] = r[16 + i]; }      // maxima as in Control_Print
  API_END(h)
}
int alf_b200_accept_log(alf_b200_handle* h, int enable) {
  API_BEGIN(h) NEED_FINAL(h)
  h->acclog_on = enable != 0; h->acclog_pos = 0;
  const long want = 2L * h->ltrot * h->n_opv * (enable > 0 ? enable : 1);    // enable = number of sweeps to record
  if (enable && h->d_acclog && want > h->acclog_per_chain) { CK(cudaFree(h->d_acclog)); h->d_acclog = nullptr; }
  if (enable && !h->d_acclog) { h->acclog_per_chain = want; CK(cudaMalloc(&h->d_acclog, (size_t)h->acclog_per_chain * h->n_chains)); }
  if (h->d_acclog) CK(cudaMemset(h->d_acclog, 255, (size_t)h->acclog_per_chain * h->n_chains));
  API_END(h)
}
// layout on device: [slice visit v][chain][step]; returned as out[chain][v*M + step]
int alf_b200_get_accept_log(alf_b200_handle* h, uint8_t* out, long cap, long* n_per_chain) {
  API_BEGIN(h) NEED_FINAL(h)
  if (!h->d_acclog) return ALF_ERROR_GENERIC;
  const long C = h->n_chains, M = h->n_opv, n = h->acclog_pos;
  if (n_per_chain) *n_per_chain = n;
  std::vector<uint8_t> raw((size_t)n * C);
  CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(raw.data(), h->d_acclog, raw.size(), cudaMemcpyDeviceToHost));
  for (long c = 0; c < C; ++c) for (long v = 0; v < n / M; ++v) for (long s = 0; s < M; ++s) { long o = c * n + v * M + s; if (o < cap) out[o] = raw[(size_t)(v * M) * C + c * M + s]; }
  API_END(h)
}
int alf_b200_taum_capture(alf_b200_handle* h, int every) { if (!h) return ALF_ERROR_GENERIC; h->taum_every = every; h->taum_host.clear(); h->taum_fresh_host.clear(); return ALF_OK; }
int alf_b200_get_taum(alf_b200_handle* h, int chain, double* out, long cap, long* n) {
  if (!h) return ALF_ERROR_GENERIC;
  if (chain < 0 || chain >= (int)h->taum_host.size()) { if (n) *n = 0; return ALF_OK; }
  const auto& v = h->taum_host[chain]; if (n) *n = (long)v.size();
  if (out) std::memcpy(out, v.data(), sizeof(cd) * std::min<long>((long)v.size(), cap));
  return ALF_OK;
}
int alf_b200_get_taum_fresh(alf_b200_handle* h, int chain, double* out, long cap, long* n) {
  if (!h) return ALF_ERROR_GENERIC;
  if (chain < 0 || chain >= (int)h->taum_fresh_host.size()) { if (n) *n = 0; return ALF_OK; }
  const auto& v = h->taum_fresh_host[chain]; if (n) *n = (long)v.size();
  if (out) std::memcpy(out, v.data(), sizeof(cd) * std::min<long>((long)v.size(), cap));
  return ALF_OK;
}
int alf_b200_obs_size(const alf_b200_handle* h) { return h ? h->obs_size : 0; }
int alf_b200_obs_reset(alf_b200_handle* h) { API_BEGIN(h) NEED_FINAL(h) CK(cudaMemsetAsync(h->d_obs, 0, sizeof(double) * h->obs_size, h->stream)); API_END(h) }
int alf_b200_obs_device_ptr(alf_b200_handle* h, double** p, long* n) { if (!h || !h->finalized) return ALF_ERROR_GENERIC; *p = h->d_obs; *n = h->obs_size; return ALF_OK; }
int alf_b200_get_obs(alf_b200_handle* h, double* out) { API_BEGIN(h) NEED_FINAL(h) CK(cudaStreamSynchronize(h->stream)); CK(cudaMemcpy(out, h->d_obs, sizeof(double) * h->obs_size, cudaMemcpyDeviceToHost)); API_END(h) }

int alf_b200_get_stream(alf_b200_handle* h, void** stream) { if (!h || !stream) return ALF_ERROR_GENERIC; *stream = (void*)h->stream; return ALF_OK; }
int alf_b200_kernel_timing(alf_b200_handle* h, unsigned mask) { API_BEGIN(h) CK(cudaStreamSynchronize(h->stream)); h->prof.reset(); h->prof.timing_mask = mask; API_END(h) }
int alf_b200_get_kernel_stats(alf_b200_handle* h, double* ms, long* launches) {
  API_BEGIN(h) CK(cudaStreamSynchronize(h->stream)); h->prof.collect();
  for (int c = 0; c < KC_COUNT; ++c) { if (ms) ms[c] = h->prof.ms[c]; if (launches) launches[c] = h->prof.launches[c]; }
  API_END(h)
}
int alf_b200_get_kernel_flops(alf_b200_handle* h, double* flops) {
  API_BEGIN(h) for (int c = 0; c < KC_COUNT; ++c) flops[c] = h->prof.flops[c]; API_END(h)
}
int alf_b200_hop_apply(alf_b200_handle* h, int which, int nf, double* A) { API_BEGIN(h) NEED_FINAL(h) h->eng->hop_apply(which, nf, reinterpret_cast<cd*>(A)); API_END(h) }

// ---- kernel-level test entry points and FP64 peak microbenchmark
int alf_b200_test_qdrp(int device, int is_complex, int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_qdrp_cplx(m, n, batch, A, D, jpvt, tau, phases); else alf_t_qdrp_real(m, n, batch, A, D, jpvt, tau, phases); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_qdrp: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_qdrp_blocked(int device, int is_complex, int m, int n, int batch, double* A, double* D, int* jpvt, double* tau, double* phases, double* Q) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_qdrp_blk_cplx(m, n, batch, A, D, jpvt, tau, phases, Q); else alf_t_qdrp_blk_real(m, n, batch, A, D, jpvt, tau, phases, Q); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_qdrp_blocked: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_udv_decompose(int device, int is_complex, int n, int batch, char side, double* U, double* D, double* V) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_udv_cplx(n, batch, side, U, D, V); else alf_t_udv_real(n, batch, side, U, D, V); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_udv_decompose: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_udv_wrap_pivot(int device, int is_complex, int n1, int n2, int batch, const double* A, double* U, double* D, double* V) {
  if (n1 < 1 || n2 < 1 || n2 > n1 || n1 > 576 || batch < 1 || !A || !U || !D || !V) return ALF_ERROR_GENERIC;
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_udv_wrap_pivot_cplx(n1, n2, batch, A, U, D, V); else alf_t_udv_wrap_pivot_real(n1, n2, batch, A, U, D, V); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_udv_wrap_pivot: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_cgr(int device, int is_complex, int n, int batch, int nvar, int stab, const double* UR, const double* DR, const double* VR,
                      const double* UL, const double* DL, const double* VL, const double* detUR, const double* detUL, double* G, double* phase) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_cgr_cplx(n, batch, nvar, stab, UR, DR, VR, UL, DL, VL, detUR, detUL, G, phase);
        else alf_t_cgr_real(n, batch, nvar, stab, UR, DR, VR, UL, DL, VL, detUR, detUL, G, phase); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_cgr: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_cgrp(int device, int is_complex, int n, int n_part, int batch, const double* UR, const double* UL, double* G, double* phase) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_cgrp_cplx(n, n_part, batch, UR, UL, G, phase); else alf_t_cgrp_real(n, n_part, batch, UR, UL, G, phase); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_cgrp: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_cgr2_2(int device, int is_complex, int n, int batch, int stab, const double* U2, const double* D2, const double* V2,
                         const double* U1, const double* D1, const double* V1, double* out4) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_cgr22_cplx(n, batch, stab, U2, D2, V2, U1, D1, V1, out4); else alf_t_cgr22_real(n, batch, stab, U2, D2, V2, U1, D1, V1, out4); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_cgr2_2: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}
int alf_b200_test_gemm(int device, int is_complex, int ta, int tb, int m, int n, int k, int batch, const double* A, const double* B, double* C) {
  try { t_prof = nullptr; CK(cudaSetDevice(device)); if (is_complex) alf_t_gemm_cplx(ta, tb, m, n, k, batch, A, B, C); else alf_t_gemm_real(ta, tb, m, n, k, batch, A, B, C); }
  catch (const std::exception& e) { fprintf(stderr, "alf_b200_test_gemm: %s\n", e.what()); return ALF_ERROR_CUDA; } return ALF_OK;
}

int alf_b200_fp64_peak(int device, double* dfma_tflops, double* dmma_tflops) {
  try {
    CK(cudaSetDevice(device)); cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 20000;
    double* d; CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); float ms;
    for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); k_peak_dfma<<<blocks, threads>>>(d, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); }
    CK(cudaEventElapsedTime(&ms, e0, e1)); *dfma_tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 2; ++rep) { CK(cudaEventRecord(e0)); k_peak_dmma<<<blocks, threads>>>(d, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); }
    CK(cudaEventElapsedTime(&ms, e0, e1)); *dmma_tflops = 2.0 * 256.0 * 4.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
  } catch (const std::exception& e) { fprintf(stderr, "alf_b200_fp64_peak: %s\n", e.what()); return ALF_ERROR_CUDA; }
  return ALF_OK;
}


}  // extern "C"
