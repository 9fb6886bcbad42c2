// Host-side launch helpers for the batched linear algebra (decompose, CGR) shared by the sweep engine and the
// kernel-level test entry points.
#pragma once
#include <cstdio>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include "alf_la.cuh"
#include "alf_qrblk.cuh"
#include "alf_qrblk2.cuh"

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  throw CudaError(std::string(#call) + " -> " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); } } while (0)
#define CKL() CK(cudaGetLastError())

// Opt-in dynamic shared memory: the attribute is per-function state shared by every handle of the process (two handles with
// different models may launch the same kernel from two host threads), so it is raised once per device and function to the most
// the function can ever get (227 KB opt-in limit minus its static shared memory) instead of to what one launch needs.
template <class F>
static cudaError_t alf_raise_smem(F* f) {
  static std::mutex mu; static std::set<std::pair<int, const void*>> done;
  int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, (const void*)f})) return cudaSuccess;
  cudaFuncAttributes a; e = cudaFuncGetAttributes(&a, f); if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, ALF_MAX_DYN_SMEM - (int)a.sharedSizeBytes);
  if (e == cudaSuccess) done.insert({dev, (const void*)f});
  return e;
}

// ---- launch accounting: every kernel launch goes through a KScope so that bench.py can report gpu_launches and the
// per-kernel device time (CUDA events on the handle's stream) its roofline entry needs.
enum { KC_UPDATE = 0, KC_OPS, KC_QRP, KC_FORMQ, KC_GEMM, KC_TRSM, KC_EW, KC_OBS, KC_COUNT };
struct Prof {
  unsigned timing_mask = 0;                       // bit c set: record CUDA events around launches of category c
  long launches[KC_COUNT] = {0}; double ms[KC_COUNT] = {0};
  double flops[KC_COUNT] = {0};                    // algorithmic FP64 flops of the dense kernels (real flops; a complex FMA counts 8)
  void add_flops(int c, double f) { flops[c] += f; }
  struct Rec { int cat; cudaEvent_t a, b; };
  std::vector<Rec> recs; std::vector<cudaEvent_t> pool;
  cudaEvent_t get() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
  void collect() {                                // caller has synchronised the stream
    for (auto& r : recs) { float t = 0.f; if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) ms[r.cat] += t; pool.push_back(r.a); pool.push_back(r.b); }
    recs.clear();
  }
  void reset() { collect(); for (int c = 0; c < KC_COUNT; ++c) { launches[c] = 0; ms[c] = 0.0; flops[c] = 0.0; } }
  ~Prof() { collect(); for (auto e : pool) cudaEventDestroy(e); }
};
extern thread_local Prof* t_prof;     // defined in alf_b200.cu, set by every C-ABI entry point
struct KScope {
  Prof* p; int cat; cudaStream_t st; cudaEvent_t a, b; bool on = false;
  KScope(int c, cudaStream_t s) : p(t_prof), cat(c), st(s) {
    if (!p) return;
    p->launches[c]++;
    if (p->timing_mask & (1u << c)) { on = true; a = p->get(); b = p->get(); cudaEventRecord(a, st); }
  }
  ~KScope() { if (on) { cudaEventRecord(b, st); p->recs.push_back({cat, a, b}); } }
};
#define KL(CAT, ST, ...) do { KScope ks_(CAT, ST); __VA_ARGS__; CKL(); } while (0)
template <typename T> static inline double flop_scale() { return std::is_same<T, double>::value ? 1.0 : 4.0; }
static inline void count_flops(int cat, double f) { if (t_prof) t_prof->add_flops(cat, f); }

// grid.x of the element-wise kernels: four elements per thread (their grid-stride loops are unrolled by 4, so every thread has four loads in flight)
static inline int ew_blocks(long n) { long b = (n + 1023) / 1024; return (int)(b > 2048 ? 2048 : (b < 1 ? 1 : b)); }

template <typename T>
struct UdvDev {            // batch of UDV_State objects (Prog/udv_state_mod.F90:85-110), one per (chain, flavor)
  T* U = nullptr; T* V = nullptr; double* D = nullptr; cplx* det = nullptr;   // det = det(U), tracked instead of DET_C(U_R^H U_L)
};

template <typename T>
struct LaWork {
  int N = 0, NM = 0; cudaStream_t st = 0;
  T* W[4] = {nullptr, nullptr, nullptr, nullptr};
  T* tau = nullptr; int* jpvt = nullptr; double* Dq = nullptr; QrOut* qrout = nullptr; cplx* sc_phase = nullptr; cplx* sc_beta = nullptr;
  T* Tbuf = nullptr;      // compact-WY factors of the blocked QR: [matrix][panel][QRB_NB x QRB_NB]
  T* Rinv = nullptr;      // inverted 32 x 32 diagonal blocks of the blocked triangular solve: [matrix][(N + 32) * 32]
  long sRinv() const { return (long)(N + 32) * 32; }
  long n2() const { return (long)N * N; }
  void alloc(int n, int nm, cudaStream_t s) {
    N = n; NM = nm; st = s;
    for (int i = 0; i < 4; ++i) CK(cudaMalloc(&W[i], sizeof(T) * n2() * NM));
    CK(cudaMalloc(&tau, sizeof(T) * (long)N * NM)); CK(cudaMalloc(&jpvt, sizeof(int) * (long)N * NM));
    CK(cudaMalloc(&Dq, sizeof(double) * (long)N * NM)); CK(cudaMalloc(&qrout, sizeof(QrOut) * NM));
    CK(cudaMalloc(&sc_phase, sizeof(cplx) * NM)); CK(cudaMalloc(&sc_beta, sizeof(cplx) * NM));
    CK(cudaMalloc(&Tbuf, sizeof(T) * (size_t)(N + 32) * 32 * NM));
    CK(cudaMalloc(&Rinv, sizeof(T) * (size_t)(N + 32) * 32 * NM));
  }
  void release() {
    for (int i = 0; i < 4; ++i) if (W[i]) cudaFree(W[i]);
    if (tau) cudaFree(tau); if (jpvt) cudaFree(jpvt); if (Dq) cudaFree(Dq); if (qrout) cudaFree(qrout);
    if (sc_phase) cudaFree(sc_phase); if (sc_beta) cudaFree(sc_beta); if (Tbuf) cudaFree(Tbuf); Tbuf = nullptr; if (Rinv) cudaFree(Rinv); Rinv = nullptr;
    for (int i = 0; i < 4; ++i) W[i] = nullptr; tau = nullptr; jpvt = nullptr; Dq = nullptr; qrout = nullptr; sc_phase = sc_beta = nullptr;
  }
};

template <typename T, int TA, int TB, int MASK>
static void gemm(cudaStream_t st, int M, int N, int K, const T* A, int lda, long sA, const T* B, int ldb, long sB, T* C, int ldc, long sC, int batch) {
  dim3 grid(((M + GEMM_BM - 1) / GEMM_BM) * ((N + GEMM_BN - 1) / GEMM_BN), batch);
  count_flops(KC_GEMM, flop_scale<T>() * 2.0 * M * N * K * batch * (MASK ? 0.5 : 1.0));
  KL(KC_GEMM, st, k_gemm<T, TA, TB, MASK><<<grid, 256, 0, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC));
}

static const size_t kSmemStageLimit = 200 * 1024;

template <typename T, int PIVOT>
static void launch_qrp(cudaStream_t st, T* A, int m, int n, int ld, long sA, T* tau, long sTau, int* jpvt, long sP, double* D, long sD, QrOut* out, int batch) {
  size_t extra = sizeof(T) * m + sizeof(double) * (n + (n > 32 ? n : 32)) + sizeof(int) * n + 64;
  size_t stage = sizeof(T) * (size_t)m * n;
  bool do_stage = (stage + extra) <= kSmemStageLimit;
  size_t smem = extra + (do_stage ? stage : 0);
#define QRP_LAUNCH(MAXR, STG) do { \
    CK(alf_raise_smem(k_qrp<T, MAXR, PIVOT, STG>)); \
    k_qrp<T, MAXR, PIVOT, STG><<<batch, 512, smem, st>>>(A, m, n, ld, sA, tau, sTau, jpvt, sP, D, sD, out); } while (0)
#define QRP_DISPATCH(MAXR) do { if (do_stage) QRP_LAUNCH(MAXR, 1); else QRP_LAUNCH(MAXR, 0); } while (0)
  count_flops(KC_QRP, flop_scale<T>() * (2.0 * m * n * n - 2.0 * n * n * n / 3.0) * batch);
  if (m <= 64 && n <= 64 && n <= m && !getenv("ALF_B200_NO_SMALL_QR")) {      // thread-per-column kernel, several matrices per SM
    const size_t sm = qrp_small_smem(sizeof(T), m);
    CK(alf_raise_smem(k_qrp_small<T, PIVOT>));
    KL(KC_QRP, st, k_qrp_small<T, PIVOT><<<batch, 64, sm, st>>>(A, m, n, ld, sA, tau, sTau, jpvt, sP, D, sD, out));
    return;
  }
  { KScope ks_(KC_QRP, st);
  if (m <= 64) QRP_DISPATCH(2); else if (m <= 128) QRP_DISPATCH(4); else if (m <= 288) QRP_DISPATCH(9); else if (m <= 576) QRP_DISPATCH(18);
  else throw CudaError("k_qrp: matrices with more than 576 rows are not supported in this build");
  CKL(); }
#undef QRP_DISPATCH
#undef QRP_LAUNCH
}

template <typename T>
static void launch_formq(cudaStream_t st, T* A, int m, int n, int ld, long sA, const T* tau, long sTau, const cplx* colscale, int batch) {
  size_t extra = sizeof(T) * m + 64;
  size_t stage = sizeof(T) * (size_t)m * n;
  bool do_stage = (stage + extra) <= kSmemStageLimit;
  size_t smem = extra + (do_stage ? stage : 0);
#define FQ_LAUNCH(MAXR, STG) do { \
    CK(alf_raise_smem(k_formq<T, MAXR, STG>)); \
    k_formq<T, MAXR, STG><<<batch, 512, smem, st>>>(A, m, n, ld, sA, tau, sTau, colscale); } while (0)
#define FQ_DISPATCH(MAXR) do { if (do_stage) FQ_LAUNCH(MAXR, 1); else FQ_LAUNCH(MAXR, 0); } while (0)
  count_flops(KC_FORMQ, flop_scale<T>() * (2.0 * m * n * n - 2.0 * n * n * n / 3.0) * batch);
  { KScope ks_(KC_FORMQ, st);
  if (m <= 64) FQ_DISPATCH(2); else if (m <= 128) FQ_DISPATCH(4); else if (m <= 288) FQ_DISPATCH(9); else if (m <= 576) FQ_DISPATCH(18);
  else throw CudaError("k_formq: matrices with more than 576 rows are not supported in this build");
  CKL(); }
#undef FQ_DISPATCH
#undef FQ_LAUNCH
}

// real matrices: blocked DMMA solve; rinv = workspace of ((n + 31) / 32) * 1024 doubles per matrix (stride sI)
template <int LOWER>
static void launch_trsm_blk(cudaStream_t st, const double* R, int ldr, long sR, double* B, int ldb, long sB, int n, int nrhs, const double* dinv, long sD,
                            double* rinv, long sI, int batch, int zero_cols_from = 0, int zero_rows = 0, const double* dout = nullptr) {
  const int np = (n + 31) & ~31, nb = np / 32;
  count_flops(KC_TRSM, 1.0 * n * n * nrhs * batch);
  const size_t smem = sizeof(double) * (size_t)ld_pad(np) * TRSMB_CW;
  KL(KC_TRSM, st, k_tri_inv_blocks<LOWER><<<dim3(nb, batch), 32, 0, st>>>(R, ldr, sR, n, rinv, sI));
  CK(alf_raise_smem(k_trsm_blk<LOWER>));
  const int nthr = (2 * (smem + 1024) <= 227 * 1024) ? 256 : 512;      // one resident CTA per SM only: give it 16 warps
  KL(KC_TRSM, st, k_trsm_blk<LOWER><<<dim3((nrhs + TRSMB_CW - 1) / TRSMB_CW, batch), nthr, smem, st>>>(R, ldr, sR, rinv, sI, B, ldb, sB, n, nrhs, dinv, sD, zero_cols_from, zero_rows, dout));
}
template <int LOWER>
static void launch_trsm_blk_c(cudaStream_t st, const cplx* R, int ldr, long sR, cplx* B, int ldb, long sB, int n, int nrhs, const double* dinv, long sD,
                              cplx* rinv, long sI, int batch) {
  const int np = (n + 31) & ~31, nb = np / 32;
  count_flops(KC_TRSM, 4.0 * n * n * nrhs * batch);
  const size_t smem = sizeof(cplx) * (size_t)(np + 1) * TRSMB_CWC;
  KL(KC_TRSM, st, k_tri_inv_blocks_c<LOWER><<<dim3(nb, batch), 32, 0, st>>>(R, ldr, sR, n, rinv, sI));
  CK(alf_raise_smem(k_trsm_blk_c<LOWER>));
  const int nthr = (2 * (smem + 1024) <= 227 * 1024) ? 256 : 512;
  KL(KC_TRSM, st, k_trsm_blk_c<LOWER><<<dim3((nrhs + TRSMB_CWC - 1) / TRSMB_CWC, batch), nthr, smem, st>>>(R, ldr, sR, rinv, sI, B, ldb, sB, n, nrhs, dinv, sD));
}
static inline bool la_force_old_trsm() { static int v = -1; if (v < 0) v = getenv("ALF_B200_OLD_TRSM") ? 1 : 0; return v == 1; }

template <typename T, int LOWER = 0>
static void launch_trsm(cudaStream_t st, const T* R, int ldr, long sR, T* B, int ldb, long sB, int n, int nrhs, const double* dinv, long sD, int batch,
                        T* rinv = nullptr, long sI = 0) {
  if constexpr (std::is_same<T, double>::value) {
    if (rinv && !la_force_old_trsm() && sizeof(double) * (size_t)ld_pad((n + 31) & ~31) * TRSMB_CW <= 220 * 1024) {
      launch_trsm_blk<LOWER>(st, R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD, rinv, sI, batch); return;
    }
  } else {
    if (rinv && !la_force_old_trsm() && sizeof(cplx) * (size_t)(((n + 31) & ~31) + 1) * TRSMB_CWC <= 220 * 1024) {
      launch_trsm_blk_c<LOWER>(st, R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD, rinv, sI, batch); return;
    }
  }
  dim3 grid((nrhs + TRSM_COLS - 1) / TRSM_COLS, batch);
  count_flops(KC_TRSM, flop_scale<T>() * n * n * nrhs * batch);
  size_t smem = sizeof(T) * n;
  KScope ks_(KC_TRSM, st);
  if (n <= 64) k_trsm_lun<T, 2, LOWER><<<grid, TRSM_WARPS * 32, smem, st>>>(R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD);
  else if (n <= 128) k_trsm_lun<T, 4, LOWER><<<grid, TRSM_WARPS * 32, smem, st>>>(R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD);
  else if (n <= 288) k_trsm_lun<T, 9, LOWER><<<grid, TRSM_WARPS * 32, smem, st>>>(R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD);
  else if (n <= 576) k_trsm_lun<T, 18, LOWER><<<grid, TRSM_WARPS * 32, smem, st>>>(R, ldr, sR, B, ldb, sB, n, nrhs, dinv, sD);
  else throw CudaError("k_trsm_lun: n > 576 not supported in this build");
  CKL();
}

// ---- blocked path (matrices that do not fit one SM's shared memory): windowed pivoted QR + compact-WY application
struct QrBlkCfg { int NB, TC; size_t smem; bool ok; };
template <typename T>
static QrBlkCfg qrblk_cfg(int m, int n) {
  QrBlkCfg c; c.ok = false;
  const int nbs[2] = {32, 16};
  for (int i = 0; i < 2; ++i) { c.NB = nbs[i]; c.TC = 8; c.smem = qrblk_smem<T>(m, n, c.NB, c.TC); if (c.smem <= 226 * 1024) { c.ok = true; break; } }
  return c;
}
static inline bool la_force_unblocked() { static int v = -1; if (v < 0) v = getenv("ALF_B200_UNBLOCKED_QR") ? 1 : 0; return v == 1; }
// true if the matrix is factored by the blocked kernel (the whole-matrix-in-shared-memory kernel is kept for small matrices)
template <typename T>
static bool use_blocked_qr(int m, int n) {
  if (la_force_unblocked()) return false;
  size_t extra = sizeof(T) * m + sizeof(double) * (n + (n > 32 ? n : 32)) + sizeof(int) * n + 64;
  return sizeof(T) * (size_t)m * n + extra > kSmemStageLimit && m <= 1024 && qrblk_cfg<T>(m, n).ok;
}
static inline bool la_force_qr1() { static int v = -1; if (v < 0) v = getenv("ALF_B200_QR1") ? 1 : 0; return v == 1; }
// real matrices with m >= n, m <= 576: register-resident panels + strip updates (alf_qrblk2.cuh)
template <typename T> static bool use_qr2(int m, int n) { return std::is_same<T, double>::value && m >= n && m <= 576 && !la_force_qr1() && qr2_smem(m, n, 16) <= 226 * 1024; }
template <typename T>
static void launch_qrp_blk(cudaStream_t st, T* A, int m, int n, int ld, long sA, T* tau, long sTau, int* jpvt, long sP, double* D, long sD, QrOut* out,
                           T* Tbuf, int batch) {
  count_flops(KC_QRP, flop_scale<T>() * (2.0 * m * n * n - 2.0 * n * n * n / 3.0) * batch);
  if constexpr (std::is_same<T, double>::value) {
    if (use_qr2<T>(m, n)) {
      const long sT = (long)(n + 32) * 32;
      // one CTA of 16 warps per matrix (two CTAs of 8 warps per SM, four panel columns per warp, measured slower twice: 1.94 vs 1.72 ms in round 1,
      // 1.79 vs 1.35 ms with round 2's panel loop -- the column step is bound by the instruction stream of a warp, not by latency)
      const size_t smem = qr2_smem(m, n, 16);
#define QR2_LAUNCH(MAXR, CPW) do { CK(alf_raise_smem(k_qrp_reg<MAXR, CPW>)); \
        KL(KC_QRP, st, k_qrp_reg<MAXR, CPW><<<batch, 32 * (QR2_NB / CPW), smem, st>>>(A, m, n, ld, sA, tau, sTau, jpvt, sP, D, sD, out, Tbuf, sT)); } while (0)
      if (m <= 128) QR2_LAUNCH(4, 2); else if (m <= 256) QR2_LAUNCH(8, 2); else if (m <= 288) QR2_LAUNCH(9, 2); else QR2_LAUNCH(18, 2);
#undef QR2_LAUNCH
      return;
    }
  }
  const QrBlkCfg c = qrblk_cfg<T>(m, n);
  CK(alf_raise_smem(k_qrp_blk<T>));
  KL(KC_QRP, st, k_qrp_blk<T><<<batch, 512, c.smem, st>>>(A, m, n, ld, sA, tau, sTau, jpvt, sP, D, sD, out, Tbuf, (long)(n + 32) * 32, c.NB, c.TC));
}
// X <- Q^H X (mode 0) / Q X (mode 1); ident: X holds the identity on entry (mode 1 only: forms Q)
template <typename T>
static void launch_apply_q(cudaStream_t st, const T* QR, int m, int n, int ld, long sQ, const T* Tbuf, T* X, int ldx, long sX, int ncols, int mode, bool ident, int batch) {
  // ZUNMQR count 4 m n ncols - 2 n^2 ncols; forming Q from the identity (ZUNGQR) touches half of it
  count_flops(KC_FORMQ, flop_scale<T>() * (4.0 * m * n * ncols - 2.0 * n * n * ncols) * (ident ? 0.5 : 1.0) * batch);
  if constexpr (std::is_same<T, double>::value) {
    if (use_qr2<T>(m, n)) {
      if (applyq2p_smem(m, 16) <= 226 * 1024 && !getenv("ALF_B200_NO_Q_PREFETCH")) {      // double-buffered reflector panels (cp.async prefetch)
        const size_t smemp = applyq2p_smem(m, 16); const int cpcp = 128; dim3 gridp((ncols + cpcp - 1) / cpcp, batch);
        const long sTp = (long)(n + 32) * 32;
        KScope ks_(KC_FORMQ, st);
#define AQ2P_LAUNCH(MD, ID) do { CK(alf_raise_smem(k_apply_q2p<MD, ID>)); \
          k_apply_q2p<MD, ID><<<gridp, 512, smemp, st>>>(QR, m, n, ld, sQ, Tbuf, sTp, X, ldx, sX, ncols, cpcp); } while (0)
        if (mode == 0) AQ2P_LAUNCH(0, 0); else if (ident) AQ2P_LAUNCH(1, 1); else AQ2P_LAUNCH(1, 0);
#undef AQ2P_LAUNCH
        CKL();
        return;
      }
      const bool small = applyq2_smem(m, 8) + 1024 <= 113 * 1024;        // two 8-warp CTAs per SM, else one 16-warp CTA
      const size_t smem = applyq2_smem(m, small ? 8 : 16); const int cpc2 = small ? 64 : 128; dim3 grid2((ncols + cpc2 - 1) / cpc2, batch);
      const long sT2 = (long)(n + 32) * 32;
      KScope ks_(KC_FORMQ, st);
#define AQ2_LAUNCH(MD, ID, NT) do { CK(alf_raise_smem(k_apply_q2<MD, ID, NT>)); \
        k_apply_q2<MD, ID, NT><<<grid2, NT, smem, st>>>(QR, m, n, ld, sQ, Tbuf, sT2, X, ldx, sX, ncols, cpc2); } while (0)
      if (small) { if (mode == 0) AQ2_LAUNCH(0, 0, 256); else if (ident) AQ2_LAUNCH(1, 1, 256); else AQ2_LAUNCH(1, 0, 256); }
      else { if (mode == 0) AQ2_LAUNCH(0, 0, 512); else if (ident) AQ2_LAUNCH(1, 1, 512); else AQ2_LAUNCH(1, 0, 512); }
#undef AQ2_LAUNCH
      CKL();
      return;
    }
  }
  const QrBlkCfg c = qrblk_cfg<T>(m, n);
  const int cpc = 64; dim3 grid((ncols + cpc - 1) / cpc, batch);
  const long sT = (long)(n + 32) * 32;
  KScope ks_(KC_FORMQ, st);
  if (mode == 0) { CK(alf_raise_smem(k_apply_q<T, 0, 0>));
    k_apply_q<T, 0, 0><<<grid, 512, c.smem, st>>>(QR, m, n, ld, sQ, Tbuf, sT, X, ldx, sX, ncols, c.NB, c.TC, cpc); }
  else if (ident) { CK(alf_raise_smem(k_apply_q<T, 1, 1>));
    k_apply_q<T, 1, 1><<<grid, 512, c.smem, st>>>(QR, m, n, ld, sQ, Tbuf, sT, X, ldx, sX, ncols, c.NB, c.TC, cpc); }
  else { CK(alf_raise_smem(k_apply_q<T, 1, 0>));
    k_apply_q<T, 1, 0><<<grid, 512, c.smem, st>>>(QR, m, n, ld, sQ, Tbuf, sT, X, ldx, sX, ncols, c.NB, c.TC, cpc); }
  CKL();
}
// pivoted QR dispatcher used by the sweep
template <typename T>
static void la_qrp(LaWork<T>& w, T* A, int m, int n, double* D) {
  if (use_blocked_qr<T>(m, n)) launch_qrp_blk<T>(w.st, A, m, n, m, (long)m * n, w.tau, n, w.jpvt, n, D, n, w.qrout, w.Tbuf, w.NM);
  else launch_qrp<T, 1>(w.st, A, m, n, m, (long)m * n, w.tau, n, w.jpvt, n, D, n, w.qrout, w.NM);
}

// pivoted QR of m x n matrices stored with leading dimension ld and batch stride sA (tau / jpvt strides = w.N, D stride sD)
template <typename T>
static void la_qrp_gen(LaWork<T>& w, T* A, int m, int n, int ld, long sA, double* D, long sD) {
  if (use_blocked_qr<T>(m, n)) launch_qrp_blk<T>(w.st, A, m, n, ld, sA, w.tau, w.N, w.jpvt, w.N, D, sD, w.qrout, w.Tbuf, w.NM);
  else launch_qrp<T, 1>(w.st, A, m, n, ld, sA, w.tau, w.N, w.jpvt, w.N, D, sD, w.qrout, w.NM);
}

// phase bookkeeping of decompose (udv_state_mod.F90:480-492, 578): Phase = prod R_ii * sign(perm), conjugated for side L;
// beta = 1/Phase scales row 1 of R, Phase scales column 1 of U.  det(U_new) = det(Q) * Phase.
static __global__ void k_decomp_phase(const QrOut* __restrict__ q, int side_l, cplx* __restrict__ ph, cplx* __restrict__ beta, cplx* __restrict__ det, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= n) return;
  cplx p = q[b].perm_sign * q[b].diag_phase;
  if (side_l) p = conj_(p);
  ph[b] = p; beta[b] = cplx(1.0, 0.0) / p; det[b] = q[b].detq * p;
}

// decompose_UDV_state (Prog/udv_state_mod.F90:448-582, default branch) for a batch
template <typename T>
static void la_decompose(LaWork<T>& w, UdvDev<T>& s, char side) {
  const int N = w.N, NM = w.NM; const long n2 = w.n2(); cudaStream_t st = w.st;
  const bool left = (side == 'l' || side == 'L');
  KL(KC_EW, st, k_colscale<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(s.U, N, n2, N, N, s.D, N));
  const bool blk = use_blocked_qr<T>(N, N);
  la_qrp<T>(w, s.U, N, N, s.D);
  KL(KC_EW, st, k_decomp_phase<<<(NM + 127) / 128, 128, 0, st>>>(w.qrout, left ? 1 : 0, w.sc_phase, w.sc_beta, s.det, NM));
  KL(KC_EW, st, k_row0scale<T><<<dim3((N + 255) / 256, NM), 256, 0, st>>>(s.U, N, n2, N, w.sc_beta));
  if (!left) {
    KL(KC_EW, st, k_permcopy<T, 1><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(w.W[0], N, n2, s.V, N, n2, N, N, w.jpvt, N));
    gemm<T, 0, 0, 1>(st, N, N, N, s.U, N, n2, w.W[0], N, n2, s.V, N, n2, NM);          // V = R * (P^T V)
  } else {
    KL(KC_EW, st, k_permcopy<T, 2><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(w.W[0], N, n2, s.V, N, n2, N, N, w.jpvt, N));
    gemm<T, 0, 1, 2>(st, N, N, N, w.W[0], N, n2, s.U, N, n2, s.V, N, n2, NM);          // V = (V P) * R^H
  }
  if (!blk) launch_formq<T>(st, s.U, N, N, N, n2, w.tau, N, w.sc_phase, NM);
  else {   // Q = H_1 ... H_n applied to the identity block-wise, then column 1 scaled by the phase (udv_state_mod.F90:576-578)
    KL(KC_EW, st, k_set_identity<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(w.W[1], N, n2, N, N));
    launch_apply_q<T>(st, s.U, N, N, N, n2, w.Tbuf, w.W[1], N, n2, N, 1, true, NM);
    KL(KC_EW, st, k_copy_col0scale<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(s.U, w.W[1], n2, N, w.sc_phase, n2));
  }
}

// decompose_UDV_state for the projective algorithm (no V: udv_state_mod.F90:473-493,576-578): U (N x N_part, leading dimension N)
// <- Q of the pivoted QR, first column scaled by the phase; D <- |R_ii|.
template <typename T>
static void la_decompose_proj(LaWork<T>& w, UdvDev<T>& s, char side, int NP) {
  const int N = w.N, NM = w.NM; const long n2 = w.n2(); cudaStream_t st = w.st;
  const bool left = (side == 'l' || side == 'L');
  const bool blk = use_blocked_qr<T>(N, NP);
  la_qrp_gen<T>(w, s.U, N, NP, N, n2, s.D, N);
  KL(KC_EW, st, k_decomp_phase<<<(NM + 127) / 128, 128, 0, st>>>(w.qrout, left ? 1 : 0, w.sc_phase, w.sc_beta, s.det, NM));
  if (!blk) launch_formq<T>(st, s.U, N, NP, N, n2, w.tau, N, w.sc_phase, NM);
  else {
    KL(KC_EW, st, k_set_identity<T><<<dim3(ew_blocks((long)N * NP), NM), 256, 0, st>>>(w.W[1], N, n2, N, NP));
    launch_apply_q<T>(st, s.U, N, NP, N, n2, w.Tbuf, w.W[1], N, n2, NP, 1, true, NM);
    KL(KC_EW, st, k_copy_col0scale<T><<<dim3(ew_blocks((long)N * NP), NM), 256, 0, st>>>(s.U, w.W[1], n2, N, w.sc_phase, (long)N * NP));
  }
}

// phase of det(U_L^H U_R) from its pivoted QR: det(Q) prod R_ii/|R_ii| sign(P)   (CGRP, Prog/cgr1_mod.F90:497-506)
static __global__ void k_cgrp_z(const QrOut* __restrict__ q, cplx* __restrict__ z, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= n) return;
  z[b] = (q[b].detq * q[b].diag_phase) * q[b].perm_sign;
}
template <typename T>
__global__ void k_one_minus(T* __restrict__ G, long sM, int n) {
  const int b = blockIdx.y; G += (long)b * sM;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)((unsigned)e / (unsigned)n), i = (int)((unsigned)e - (unsigned)j * (unsigned)n);
    G[e] = ((i == j) ? one_<T>() : zero_<T>()) - G[e];
  }
}

// per-matrix phase factor of det(1 + B_R B_L) without Op_phase (Prog/cgr1_mod.F90:300-349)
static __global__ void k_cgr_z(const QrOut* __restrict__ q, const cplx* __restrict__ detR, const cplx* __restrict__ detL, int nvar, cplx* __restrict__ z, int n) {
  int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= n) return;
  cplx dp = q[b].diag_phase, dq = q[b].detq;
  if (nvar != 1) { dp = conj_(dp); dq = conj_(dq); }
  z[b] = ((detR[b] * conj_(detL[b])) * q[b].perm_sign) * (dp * dq);
}

// CGR (Prog/cgr1_mod.F90:176-447): G = (1 + B_R B_L)^-1 and the phase factor z per matrix.  stab = 3 selects the
// scale-separated branch.  Gout must not alias any input.
template <typename T>
static void la_cgr(LaWork<T>& w, int nvar, int stab, const UdvDev<T>& R, const UdvDev<T>& L, T* Gout, cplx* z) {
  const int N = w.N, NM = w.NM; const long n2 = w.n2(); cudaStream_t st = w.st;
  dim3 eg(ew_blocks(n2), NM);
  gemm<T, 1, 0, 0>(st, N, N, N, R.U, N, n2, L.U, N, n2, w.W[0], N, n2, NM);            // RHS = U_R^H U_L
  gemm<T, 0, 0, 0>(st, N, N, N, R.V, N, n2, L.V, N, n2, w.W[1], N, n2, NM);            // TPUP = V_R V_L
  if (stab == 3) { if (nvar == 1) KL(KC_EW, st, k_cgr_tpup<T, 1, 0><<<eg, 256, 0, st>>>(w.W[2], w.W[1], w.W[0], n2, N, R.D, L.D, N));
                   else KL(KC_EW, st, k_cgr_tpup<T, 1, 1><<<eg, 256, 0, st>>>(w.W[2], w.W[1], w.W[0], n2, N, R.D, L.D, N)); }
  else { if (nvar == 1) KL(KC_EW, st, k_cgr_tpup<T, 0, 0><<<eg, 256, 0, st>>>(w.W[2], w.W[1], w.W[0], n2, N, R.D, L.D, N));
         else KL(KC_EW, st, k_cgr_tpup<T, 0, 1><<<eg, 256, 0, st>>>(w.W[2], w.W[1], w.W[0], n2, N, R.D, L.D, N)); }
  CKL();
  const bool blk = use_blocked_qr<T>(N, N);
  la_qrp<T>(w, w.W[2], N, N, w.Dq);
  KL(KC_EW, st, k_cgr_z<<<(NM + 127) / 128, 128, 0, st>>>(w.qrout, R.det, L.det, nvar, z, NM));
  // X0 = U_R^H (nvar 1) or U_L^H (nvar 2), with the D_+^-1 row scaling of the STAB3 branch
  const UdvDev<T>& A0 = (nvar == 1) ? R : L;
  const UdvDev<T>& A1 = (nvar == 1) ? L : R;
  if (!blk) {
    // explicit Q in W[1]
    KL(KC_EW, st, k_permcopy<T, 0><<<eg, 256, 0, st>>>(w.W[1], N, n2, w.W[2], N, n2, N, N, nullptr, 0));
    launch_formq<T>(st, w.W[1], N, N, N, n2, w.tau, N, nullptr, NM);
    KL(KC_EW, st, k_transpose_conj<T><<<dim3(((N + 31) / 32) * ((N + 31) / 32), NM), 256, 0, st>>>(w.W[0], N, n2, A0.U, N, n2, N));
    if (stab == 3) { KL(KC_EW, st, k_sep_scale<T, 1><<<eg, 256, 0, st>>>(w.W[0], n2, N, A0.D, N)); }
    gemm<T, 1, 0, 0>(st, N, N, N, w.W[1], N, n2, w.W[0], N, n2, w.W[3], N, n2, NM);      // X = Q^H X0
  } else {   // ZUNMQR: the block reflectors are applied to X0 directly
    KL(KC_EW, st, k_transpose_conj<T><<<dim3(((N + 31) / 32) * ((N + 31) / 32), NM), 256, 0, st>>>(w.W[3], N, n2, A0.U, N, n2, N));
    if (stab == 3) { KL(KC_EW, st, k_sep_scale<T, 1><<<eg, 256, 0, st>>>(w.W[3], n2, N, A0.D, N)); }
    launch_apply_q<T>(st, w.W[2], N, N, N, n2, w.Tbuf, w.W[3], N, n2, N, 0, false, NM);
  }
  launch_trsm<T>(st, w.W[2], N, n2, w.W[3], N, n2, N, N, w.Dq, N, NM, w.Rinv, w.sRinv());                 // X = R^-1 D^-1 X
  KL(KC_EW, st, k_permcopy<T, 3><<<eg, 256, 0, st>>>(w.W[0], N, n2, w.W[3], N, n2, N, N, w.jpvt, N));   // X2 = P X
  if (stab == 3) { KL(KC_EW, st, k_sep_scale<T, 1><<<eg, 256, 0, st>>>(w.W[0], n2, N, A1.D, N)); }
  if (nvar == 1) gemm<T, 0, 0, 0>(st, N, N, N, L.U, N, n2, w.W[0], N, n2, Gout, N, n2, NM);        // G = U_L X2
  else gemm<T, 1, 1, 0>(st, N, N, N, w.W[0], N, n2, R.U, N, n2, Gout, N, n2, NM);                 // G = X2^H U_R^H
}


// A^-1 for a batch of well-conditioned N x N matrices (replaces INV = ZGETRF + ZGETRI, Libraries/Modules/mymats_mod.F90:546-580,
// at its only hot-path call site, V1INV of CGR2_2, Prog/cgr2_2_mod.F90:349) through the pivoted QR already on the device:
// A P = Q D R  =>  A^-1 = P R^-1 D^-1 Q^H.   A is destroyed; Ainv must not alias A.  Uses w.W[0], w.W[1].
template <typename T>
static void la_inverse(LaWork<T>& w, T* A, T* Ainv) {
  const int N = w.N, NM = w.NM; const long n2 = w.n2(); cudaStream_t st = w.st; dim3 eg(ew_blocks(n2), NM);
  la_qrp<T>(w, A, N, N, w.Dq);
  if (!use_blocked_qr<T>(N, N)) {
    KL(KC_EW, st, k_permcopy<T, 0><<<eg, 256, 0, st>>>(w.W[0], N, n2, A, N, n2, N, N, nullptr, 0));
    launch_formq<T>(st, w.W[0], N, N, N, n2, w.tau, N, nullptr, NM);
    KL(KC_EW, st, k_transpose_conj<T><<<dim3(((N + 31) / 32) * ((N + 31) / 32), NM), 256, 0, st>>>(w.W[1], N, n2, w.W[0], N, n2, N));      // Q^H
  } else {
    KL(KC_EW, st, k_set_identity<T><<<eg, 256, 0, st>>>(w.W[1], N, n2, N, N));
    launch_apply_q<T>(st, A, N, N, N, n2, w.Tbuf, w.W[1], N, n2, N, 0, false, NM);                            // Q^H
  }
  launch_trsm<T>(st, A, N, n2, w.W[1], N, n2, N, N, w.Dq, N, NM, w.Rinv, w.sRinv());                                            // R^-1 D^-1 Q^H
  KL(KC_EW, st, k_permcopy<T, 3><<<eg, 256, 0, st>>>(Ainv, N, n2, w.W[1], N, n2, N, N, w.jpvt, N));          // rows scattered by P
}

// CGRP (Prog/cgr1_mod.F90:464-515) for a batch: G = 1 - U_R (U_L^H U_R)^-1 U_L^H and z = phase of det(U_L^H U_R).
// w: N-sized workspace, wp: N_part-sized workspace.  The LU of the reference is replaced by the pivoted QR already on the device.
template <typename T>
static void la_cgrp(LaWork<T>& w, LaWork<T>& wp, const UdvDev<T>& R, const UdvDev<T>& L, T* Gout, cplx* z) {
  const int N = w.N, NP = wp.N, NM = w.NM; const long n2 = w.n2(), np2 = wp.n2(); cudaStream_t st = w.st;
  gemm<T, 1, 0, 0>(st, NP, NP, N, L.U, N, n2, R.U, N, n2, wp.W[2], NP, np2, NM);                   // S = U_L^H U_R
  la_inverse<T>(wp, wp.W[2], wp.W[3]);                                                             // S^-1
  KL(KC_EW, st, k_cgrp_z<<<(NM + 127) / 128, 128, 0, st>>>(wp.qrout, z, NM));
  gemm<T, 0, 1, 0>(st, NP, N, NP, wp.W[3], NP, np2, L.U, N, n2, w.W[0], NP, (long)NP * N, NM);      // rMat = S^-1 U_L^H
  gemm<T, 0, 0, 0>(st, N, N, NP, R.U, N, n2, w.W[0], NP, (long)NP * N, Gout, N, n2, NM);            // U_R rMat
  KL(KC_EW, st, k_one_minus<T><<<dim3(ew_blocks(n2), NM), 256, 0, st>>>(Gout, n2, N));
}

// CGR2_2 (Prog/cgr2_2_mod.F90:318-425) for a batch: w is the N-sized workspace, w2 the 2N-sized one (w2.NM == w.NM).
// udv2 = right propagation (side R), udv1 = left propagation (side L).  Outputs must not alias inputs.
template <typename T>
static void la_cgr2_2(LaWork<T>& w, LaWork<T>& w2, int stab, const UdvDev<T>& udv2, const UdvDev<T>& udv1, T* GT0, T* G00, T* GTT, T* G0T, int* first) {
  const int N = w.N, NM = w.NM, N2 = 2 * N; const long n2 = w.n2(), n22 = w2.n2(); cudaStream_t st = w.st;
  dim3 eg(ew_blocks(n2), NM), eg2(ew_blocks(n22), NM);
  // V1INV in w.W[3] (V1 copied to w.W[2] because la_inverse destroys its input)
  KL(KC_EW, st, k_permcopy<T, 0><<<eg, 256, 0, st>>>(w.W[2], N, n2, udv1.V, N, n2, N, N, nullptr, 0));
  la_inverse<T>(w, w.W[2], w.W[3]);
  // HLPB1 = HLPB2^H in w2.W[0]; right-hand side HLP in w2.W[1]
  if (N % 32 == 0) {      // every access coalesced through 32 x 32 tiles
    const dim3 gt((2 * N / 32) * (2 * N / 32), NM);
    if (stab == 3) KL(KC_EW, st, k_cgr22_build_tiled<T, 1><<<gt, 256, 0, st>>>(w2.W[0], w2.W[1], n22, w.W[3], udv1.U, udv1.D, udv2.U, udv2.V, udv2.D, n2, N, N, first));
    else KL(KC_EW, st, k_cgr22_build_tiled<T, 0><<<gt, 256, 0, st>>>(w2.W[0], w2.W[1], n22, w.W[3], udv1.U, udv1.D, udv2.U, udv2.V, udv2.D, n2, N, N, first));
  } else
  if (stab == 3) KL(KC_EW, st, k_cgr22_build<T, 1><<<eg2, 256, 0, st>>>(w2.W[0], w2.W[1], n22, w.W[3], udv1.U, udv1.D, udv2.U, udv2.V, udv2.D, n2, N, N, first));
  else KL(KC_EW, st, k_cgr22_build<T, 0><<<eg2, 256, 0, st>>>(w2.W[0], w2.W[1], n22, w.W[3], udv1.U, udv1.D, udv2.U, udv2.V, udv2.D, n2, N, N, first));
  la_qrp<T>(w2, w2.W[0], N2, N2, w2.Dq);
  // HLP <- P^T-row gather (ZLAPMR forward), L = R^H, HLP <- L^-1 HLP, rows / D3, HLP <- Q HLP
  KL(KC_EW, st, k_permcopy<T, 1><<<eg2, 256, 0, st>>>(w2.W[2], N2, n22, w2.W[1], N2, n22, N2, N2, w2.jpvt, N2));
  KL(KC_EW, st, k_transpose_conj<T><<<dim3(((N2 + 31) / 32) * ((N2 + 31) / 32), NM), 256, 0, st>>>(w2.W[3], N2, n22, w2.W[0], N2, n22, N2));   // full conj-transpose; only its lower triangle (R^H) is read
  bool fused = false;
  if constexpr (std::is_same<T, double>::value) {
    if (!la_force_old_trsm() && sizeof(double) * (size_t)ld_pad(N2) * TRSMB_CW <= 220 * 1024) {
      // the division by D3 rides on the write-back (the row permutation has destroyed the block-diagonal zero pattern of HLP,
      // so no block rows can be skipped: zero_rows = 0)
      launch_trsm_blk<1>(st, w2.W[3], N2, n22, w2.W[2], N2, n22, N2, N2, nullptr, N2, w2.Rinv, w2.sRinv(), NM, 0, 0, w2.Dq);
      fused = true;
    }
  }
  if (!fused) {
    launch_trsm<T, 1>(st, w2.W[3], N2, n22, w2.W[2], N2, n22, N2, N2, nullptr, 0, NM, w2.Rinv, w2.sRinv());
    KL(KC_EW, st, k_rowscale_inv<T><<<eg2, 256, 0, st>>>(w2.W[2], N2, n22, N2, N2, w2.Dq, N2));
  }
  if (!use_blocked_qr<T>(N2, N2)) {
    launch_formq<T>(st, w2.W[0], N2, N2, N2, n22, w2.tau, N2, nullptr, NM);
    gemm<T, 0, 0, 0>(st, N2, N2, N2, w2.W[0], N2, n22, w2.W[2], N2, n22, w2.W[1], N2, n22, NM);
    KL(KC_EW, st, k_cgr22_blocks<T><<<eg, 256, 0, st>>>(w2.W[1], n22, GT0, G00, GTT, G0T, n2, N, first));
  } else {
    launch_apply_q<T>(st, w2.W[0], N2, N2, N2, n22, w2.Tbuf, w2.W[2], N2, n22, N2, 1, false, NM);             // HLP <- Q HLP (ZUNMQR 'L','N')
    KL(KC_EW, st, k_cgr22_blocks<T><<<eg, 256, 0, st>>>(w2.W[2], n22, GT0, G00, GTT, G0T, n2, N, first));
  }
}
