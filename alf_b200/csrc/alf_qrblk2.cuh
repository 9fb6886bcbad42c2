// Second-generation blocked pivoted QR for REAL matrices that do not fit one SM's shared memory (N = 256: 512 KB), and the
// matching block-reflector application.  Same mathematics and the same outputs as k_qrp_blk / k_apply_q (alf_qrblk.cuh:
// windowed column pivoting, compact-WY trailing update; replaces ZGEQP3 / ZUNGQR / ZUNMQR of Prog/QDRP_decompose_mod.F90:78-100,
// Prog/udv_state_mod.F90:569-578, Prog/cgr1_mod.F90:350-445, Prog/cgr2_2_mod.F90:155-191), restructured around the two things the
// ncu source view showed the first version waiting on (profiles/r1_ncu_stab.md): block-wide barriers in the panel
// factorisation and in the 8-column tiles of the trailing update.
//  * Panel factorisation: the 32 panel columns live in REGISTERS, two columns per warp (16 warps), lane l holding rows
//    l, l+32, ...  Dot products and norms are warp-shuffle reductions, in-panel pivoting is a logical permutation (no data
//    movement), the reflector is broadcast through shared memory: two barriers per column instead of five.
//  * Trailing update / Q application: every warp owns an 8-column strip of the target and runs the whole chain
//    W = V^T C, W2 = op(T) W, C -= V W2 on the FP64 tensor cores (DMMA m8n8k4) with C read from and written to global memory
//    (L2) in fragment layout: no barrier inside the update, and the exact column norms for the next pivot window fall out of
//    the accumulator registers.
#pragma once
#include "alf_qrblk.cuh"

#define QR2_NB 32
#ifdef ALF_QR_PROF      // experimental build only: clock64 accounting of the phases of k_qrp_reg (summed over CTAs, thread 0)
__device__ unsigned long long g_qr_prof[16];
#define QRP_T0 long long qp_t = clock64();
#define QRP_ACC(i) if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_qr_prof[i], (unsigned long long)(t_ - qp_t)); qp_t = t_; }
#else
#define QRP_T0
#define QRP_ACC(i)
#endif
#define QR2_LDW 36
#define QR2_WSC (QR2_LDW * 8)      // per-warp scratch: 32 x 8 block, leading dimension 36

// nw = warps per CTA.  The Gram matrix of the panel aliases the per-warp scratch (used in disjoint phases).
static inline int qr2_vlen(int m) { return (m <= 128) ? 128 : (m <= 256) ? 256 : (m <= 288) ? 288 : 576; }      // 32 * MAXR of the instantiation used for m rows
static size_t qr2_smem(int m, int n, int nw) {
  const int mp = qr2_vlen(m), ldv = ld_pad((m + 7) & ~7);
  const size_t wsc = (size_t)nw * QR2_WSC > (size_t)QR2_LDW * QR2_NB ? (size_t)nw * QR2_WSC : (size_t)QR2_LDW * QR2_NB;
  return sizeof(double) * ((size_t)ldv * QR2_NB + QR2_LDW * QR2_NB + wsc + mp + 4 * QR2_NB + n) + sizeof(int) * (4 * (size_t)n + 64) + 64;
}
static size_t applyq2_smem(int m, int nw) {
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  return sizeof(double) * ((size_t)ldv * QR2_NB + QR2_LDW * QR2_NB + (size_t)nw * QR2_WSC) + 64;
}

// X(r0:m, strip) <- (I - V op(T) V^T) X(r0:m, strip) for the 8-column strips c_begin + 8 (warp + k nw) < c_end of this CTA.
// Vs: (m - r0) x 32 explicit unit-lower trapezoid, zero padded to mvp rows and 32 columns; Ts: 32 x 32 upper triangular (ld 36).
// CONJT = 1: op(T) = T^T (Q^T from the left), 0: op(T) = T.  vn != nullptr: 2-norm of X(r0 + nbk : m, c) -> vn[c].
// Both GEMM phases are software-pipelined: the C fragments of the next group of k-steps (phase 1) / row blocks (phase 3) are
// already in flight while the tensor cores work on the current group, so a warp never waits a full L2 round trip per group
// (round 1's ncu source view: 49 % long-scoreboard stall on the DMMA line of phase 3, which had a one-block prefetch only).
#define QR2_KU 8      // k-steps (of 4 rows) per load group in phase 1
#define QR2_RB 4      // row blocks (of 8 rows) per load group in phase 3: four independent accumulator chains
template <int CONJT>
__device__ __forceinline__ void apply_panel_strips(double* __restrict__ X, int ldx, int r0, int m, int c_begin, int c_end, const double* __restrict__ Vs,
                                                   int ldv, const double* __restrict__ Ts, int nbk, double* __restrict__ Wsc, double* __restrict__ vn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int mv = m - r0, mvp = (mv + 7) & ~7;
  double* Wm = Wsc + warp * QR2_WSC;
  for (int c0 = c_begin + 8 * warp; c0 < c_end; c0 += 8 * nw) {
    // ---- W = V^T C  (32 x 8)
    double w[4][2];
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) { w[ib][0] = 0.0; w[ib][1] = 0.0; }
    const bool colb = (c0 + g) < c_end;
    const double* xb = X + (long)(c0 + g) * ldx + r0 + q;
    const int nks = mvp / 4;
    double bc[QR2_KU], bn[QR2_KU];
#pragma unroll
    for (int u = 0; u < QR2_KU; ++u) bc[u] = (colb && 4 * u + q < mv) ? xb[4 * u] : 0.0;
    for (int ks0 = 0; ks0 < nks; ks0 += QR2_KU) {
      if (ks0 + QR2_KU < nks) {
#pragma unroll
        for (int u = 0; u < QR2_KU; ++u) { const int ks = ks0 + QR2_KU + u; bn[u] = (colb && 4 * ks + q < mv) ? xb[4 * ks] : 0.0; }
      }
#pragma unroll
      for (int u = 0; u < QR2_KU; ++u) {
        const int ks = ks0 + u;
        if (ks < nks) {                                       // warp-uniform; rows beyond mvp of Vs are not initialised
          const double* vp = Vs + (4 * ks + q) + (long)g * ldv;
#pragma unroll
          for (int ib = 0; ib < 4; ++ib) dmma884(w[ib][0], w[ib][1], vp[(long)(8 * ib) * ldv], bc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < QR2_KU; ++u) bc[u] = bn[u];
    }
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) { Wm[(8 * ib + g) + (2 * q) * QR2_LDW] = w[ib][0]; Wm[(8 * ib + g) + (2 * q + 1) * QR2_LDW] = w[ib][1]; }
    __syncwarp();
    // ---- W2 = op(T) W  (32 x 8)
    double bw[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) bw[ks] = Wm[(4 * ks + q) + g * QR2_LDW];
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) { w[ib][0] = 0.0; w[ib][1] = 0.0; }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int ib = 0; ib < 4; ++ib) {
        const double a = CONJT ? Ts[(4 * ks + q) + (8 * ib + g) * QR2_LDW] : Ts[(8 * ib + g) + (4 * ks + q) * QR2_LDW];
        dmma884(w[ib][0], w[ib][1], a, bw[ks]);
      }
    __syncwarp();
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) { Wm[(8 * ib + g) + (2 * q) * QR2_LDW] = -w[ib][0]; Wm[(8 * ib + g) + (2 * q + 1) * QR2_LDW] = -w[ib][1]; }
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) bw[ks] = Wm[(4 * ks + q) + g * QR2_LDW];       // -W2 as B fragments
    __syncwarp();
    // ---- C -= V W2, QR2_RB blocks of 8 rows at a time; exact norms of the rows below the panel
    const bool ca = (c0 + 2 * q) < c_end, cb2 = (c0 + 2 * q + 1) < c_end;
    double* xc = X + (long)(c0 + 2 * q) * ldx + r0 + g;
    double n0 = 0.0, n1 = 0.0;
    const int nrb = mvp / 8;
    double pa[QR2_RB][2], pb[QR2_RB][2];
#pragma unroll
    for (int t = 0; t < QR2_RB; ++t) { const bool ok = 8 * t + g < mv; pa[t][0] = (ok && ca) ? xc[8 * t] : 0.0; pa[t][1] = (ok && cb2) ? xc[8 * t + ldx] : 0.0; }
    for (int rb = 0; rb < nrb; rb += QR2_RB) {
      if (rb + QR2_RB < nrb) {
#pragma unroll
        for (int t = 0; t < QR2_RB; ++t) { const int rr = 8 * (rb + QR2_RB + t); const bool ok = rr + g < mv; pb[t][0] = (ok && ca) ? xc[rr] : 0.0; pb[t][1] = (ok && cb2) ? xc[rr + ldx] : 0.0; }
      }
      const double* vp = Vs + (8 * rb + g) + (long)q * ldv;
      if (rb + QR2_RB <= nrb) {                               // full group: four interleaved accumulator chains
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
          for (int t = 0; t < QR2_RB; ++t) dmma884(pa[t][0], pa[t][1], vp[8 * t + (long)(4 * ks) * ldv], bw[ks]);
      } else {
#pragma unroll
        for (int t = 0; t < QR2_RB; ++t) if (rb + t < nrb) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) dmma884(pa[t][0], pa[t][1], vp[8 * t + (long)(4 * ks) * ldv], bw[ks]);
        }
      }
#pragma unroll
      for (int t = 0; t < QR2_RB; ++t) {
        const int r = 8 * (rb + t) + g; const bool rok = r < mv;
        if (rok && ca) xc[8 * (rb + t)] = pa[t][0];
        if (rok && cb2) xc[8 * (rb + t) + ldx] = pa[t][1];
        if (rok && r >= nbk) { n0 = fma(pa[t][0], pa[t][0], n0); n1 = fma(pa[t][1], pa[t][1], n1); }
        pa[t][0] = pb[t][0]; pa[t][1] = pb[t][1];
      }
    }
    if (vn) {
      n0 += __shfl_xor_sync(0xffffffffu, n0, 4); n1 += __shfl_xor_sync(0xffffffffu, n1, 4);
      n0 += __shfl_xor_sync(0xffffffffu, n0, 8); n1 += __shfl_xor_sync(0xffffffffu, n1, 8);
      n0 += __shfl_xor_sync(0xffffffffu, n0, 16); n1 += __shfl_xor_sync(0xffffffffu, n1, 16);
      if (g == 0) { if (ca) vn[c0 + 2 * q] = sqrt(n0); if (cb2) vn[c0 + 2 * q + 1] = sqrt(n1); }
    }
  }
}

// Column loop of one panel for NL LIVE row blocks: the registers hold rows k0 + 32 r + lane, r < NL, only (k0 is a multiple of 32, so the pivot rows
// of the whole panel sit in block 0, pivot row j in lane j, and the rows above k0 -- finished R entries -- take no part).  Written after the
// instrumented build and the SASS-level stall view of the first versions (4300 cycles per column: the warp owning the pivot ran ~770 branchy
// instructions while 15 warps waited; per-row-block branches compiled to chains of convergence barriers; half of the FP64 work was spent on
// finished rows):
//  * the owner only STORES the raw pivot column x to shared memory; every warp derives the reflector itself: beta = -sign(alpha) |x|,
//    v = scal x~ with x~ = x except x~(j) = alpha - beta = 1 / scal, so H c = c - x~ (gamma x~.c) with ONE division gamma = tau scal^2 =
//    -1 / (beta (alpha - beta)); the finished column stays RAW in its registers, scal and beta are applied at the write-back;
//  * straight-line code over exactly the live blocks (one instantiation per NL), row predicates in block 0 only;
//  * the pivot search is two redux.sync on the bit patterns of the (non-negative) norms instead of five shuffle rounds.
template <int MAXR, int CPW, int NL>
__device__ __forceinline__ void qr2_panel_cols(double (&cr)[CPW][MAXR], const int nbk, const int sb, double* __restrict__ v_s, double* __restrict__ pn,
                                               double* __restrict__ tau_s, double* __restrict__ sc_s, double* __restrict__ be_s, int* __restrict__ pos_slot,
                                               double* __restrict__ s_detq) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned used = 0u;
  QRP_T0
  for (int j = 0; j < nbk; ++j) {
    int p;
    {   // arg max of pn over the unused slots, lowest slot on ties: key = bits(pn) + 1 (0 for used / empty slots), 64-bit max as two 32-bit redux
      const bool cand0 = lane < nbk && !((used >> lane) & 1u);
      const unsigned long long key = cand0 ? ((unsigned long long)__double_as_longlong(pn[lane]) + 1ull) : 0ull;
      const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
      const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
      const bool c1 = khi == mhi;
      const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? klo : 0u);
      p = __ffs(__ballot_sync(0xffffffffu, c1 && klo == mlo)) - 1;
    }
    used |= 1u << p;
    const double cn_pre = sqrt(pn[p]);      // |x| of the pivot column: off the critical path (overlaps the owner's store and the barrier)
    QRP_ACC(8)
    if (warp == p / CPW) {      // owner: raw pivot column -> shared memory (local rows j .. ; everything else of v_s is zero)
      const int hp = p % CPW;
#pragma unroll
      for (int r = 0; r < NL; ++r) {
        double x = cr[0][r];
#pragma unroll
        for (int h = 1; h < CPW; ++h) x = (h == hp) ? cr[h][r] : x;
        if (r > 0 || lane >= j) v_s[lane + 32 * r] = x;
      }
      if (lane == 0) { pos_slot[j] = p; if (j > 0) v_s[j - 1] = 0.0; }
    }
    __syncthreads();
    QRP_ACC(9)
    {
      // the reflector from the raw column (every warp, redundantly): LAPACK's ZLARFG for real data.  cn = 2-norm of the pivot column below row
      // j - 1, exact (from the previous step's update); cn == 0: H = I (ZLARFG's xnorm == 0, alpha == 0 case)
      const double alpha = v_s[j], cn = cn_pre;
      const double beta = (cn == 0.0) ? alpha : -copysign(cn, alpha);
      const double dd = alpha - beta;                          // x~(j) = 1 / scal
      const double gam = (cn == 0.0) ? 0.0 : -1.0 / (beta * dd);
      if (tid == 0) {
        tau_s[j] = (cn == 0.0) ? 0.0 : (beta - alpha) / beta; sc_s[p] = (cn == 0.0) ? 0.0 : 1.0 / dd; be_s[p] = beta;
        if (cn != 0.0) { s_detq[0] = -s_detq[0]; s_detq[1] = -s_detq[1]; }      // det of a real reflector = -1 (Prog/cgr1_mod.F90:338-347)
      }
      bool doh[CPW];
#pragma unroll
      for (int h = 0; h < CPW; ++h) doh[h] = (sb + h < nbk) && !((used >> (sb + h)) & 1u);
      double xv[NL];
      double w[CPW], wb[CPW], qn[CPW], qb[CPW];
#pragma unroll
      for (int h = 0; h < CPW; ++h) { w[h] = 0.0; wb[h] = 0.0; qn[h] = 0.0; qb[h] = 0.0; }
#pragma unroll
      for (int r = 0; r < NL; ++r) {
        xv[r] = v_s[lane + 32 * r];
        if (r == 0) xv[0] = (lane == j) ? dd : xv[0];
#pragma unroll
        for (int h = 0; h < CPW; ++h) { if (r & 1) wb[h] = fma(xv[r], cr[h][r], wb[h]); else w[h] = fma(xv[r], cr[h][r], w[h]); }
      }
      if (NL > 1) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) w[h] += wb[h];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < CPW; ++h) w[h] += __shfl_xor_sync(0xffffffffu, w[h], o);
#pragma unroll
      for (int h = 0; h < CPW; ++h) w[h] = doh[h] ? w[h] * gam : 0.0;      // finished columns keep their raw reflector entries
#pragma unroll
      for (int r = 0; r < NL; ++r) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) {
          cr[h][r] = fma(-xv[r], w[h], cr[h][r]);
          const double t = (r > 0 || lane > j) ? cr[h][r] : 0.0;          // exact norm of the rows below the pivot row
          if (r & 1) qb[h] = fma(t, t, qb[h]); else qn[h] = fma(t, t, qn[h]);
        }
      }
      if (NL > 1) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) qn[h] += qb[h];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < CPW; ++h) qn[h] += __shfl_xor_sync(0xffffffffu, qn[h], o);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) if (doh[h]) pn[sb + h] = qn[h];
      }
    }
    QRP_ACC(10)
    __syncthreads();
    QRP_ACC(11)
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_qrp_reg: windowed column-pivoted blocked Householder QR of an m x n real matrix (m >= n, m <= 32 MAXR), in place, then
// D(i) = |R(i,i)|, R(i, i:) /= D(i).  One CTA per matrix.  Outputs as k_qrp_blk (Tbuf: 32 x 32 factor per panel).
// ------------------------------------------------------------------------------------------------------------------------
template <int MAXR, int CPW>      // CPW = panel columns per warp: 2 -> 16 warps (one CTA per SM), 4 -> 8 warps (two CTAs per SM)
__global__ void __launch_bounds__(32 * (QR2_NB / CPW), CPW == 4 ? 2 : 1) k_qrp_reg(double* __restrict__ A, int m, int n, int ld, long sA, double* __restrict__ tau, long sTau,
                                                    int* __restrict__ jpvt, long sP, double* __restrict__ D, long sD, QrOut* __restrict__ out,
                                                    double* __restrict__ Tbuf, long sT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NB = QR2_NB, LDW = QR2_LDW;
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  A += (long)b * sA; tau += (long)b * sTau; jpvt += (long)b * sP; D += (long)b * sD; Tbuf += (long)b * sT;
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  double* Vs = reinterpret_cast<double*>(smem_raw);
  double* Ts = Vs + (long)ldv * NB;
  double* Wsc = Ts + LDW * NB;
  double* Gs = Wsc;                   // Gram matrix: dead before the first strip update of the panel
  double* v_s = Wsc + ((nw * QR2_WSC > LDW * NB) ? nw * QR2_WSC : LDW * NB);
  double* tau_s = v_s + 32 * MAXR;
  double* pn = tau_s + NB;
  double* sc_s = pn + NB;             // per panel slot: 1 / (alpha - beta) and beta of its reflector
  double* be_s = sc_s + NB;
  double* vn = be_s + NB;
  int* ipv = reinterpret_cast<int*>(vn + n);
  int* rank_s = ipv + n;
  int* lista = rank_s + n;
  int* listb = lista + n;
  int* pos_slot = listb + n;          // [32] slot at logical panel position j; [32..63] scratch
  __shared__ int s_na;
  __shared__ double s_detq[2];

  QRP_T0
  for (int c = warp; c < n; c += nw) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) { const double x = A[i + (long)c * ld]; s = fma(x, x, s); }
    s = warp_sum(s);
    if (lane == 0) { vn[c] = sqrt(s); ipv[c] = c; }
  }
  if (tid == 0) { s_detq[0] = 1.0; s_detq[1] = 0.0; }
  __syncthreads();
  const int kmax = n;                 // m >= n
  QRP_ACC(0)
  for (int k0 = 0; k0 < kmax; k0 += NB) {
    const int nbk = min(NB, kmax - k0), mv = m - k0, mvp = (mv + 7) & ~7;
    // ---- (1) the nbk remaining columns of largest norm form the panel (rank by counting; ties -> lower index first)
    for (int c = k0 + tid; c < n; c += nthr) {
      const double x = vn[c]; int r = 0;
      for (int c2 = k0; c2 < n; ++c2) { const double y = vn[c2]; r += (y > x || (y == x && c2 < c)) ? 1 : 0; }
      rank_s[c] = r;
    }
    __syncthreads();
    if (warp == 0) {      // lista: selected columns outside the panel range; listb: unselected columns inside it (same count)
      int na = 0, nb2 = 0;
      for (int c0 = k0; c0 < n; c0 += 32) {
        const int c = c0 + lane; const bool in = c < n;
        const bool sel = in && rank_s[c] < nbk, front = in && c < k0 + nbk;
        const unsigned ma = __ballot_sync(0xffffffffu, sel && !front), mb = __ballot_sync(0xffffffffu, !sel && front);
        if (sel && !front) lista[na + __popc(ma & ((1u << lane) - 1))] = c;
        if (!sel && front) listb[nb2 + __popc(mb & ((1u << lane) - 1))] = c;
        na += __popc(ma); nb2 += __popc(mb);
      }
      if (lane == 0) s_na = na;
    }
    __syncthreads();
    // ---- (2) bring the selected columns into the panel range (column swaps in global memory)
    const int na = s_na;
    for (int pi = warp; pi < na; pi += nw) {
      const int ca = lista[pi], cb = listb[pi];
      {   // all loads of the two columns are issued before the first store (the read-modify-write loop serialised on the global latency)
        double xa[MAXR], xb[MAXR];
#pragma unroll
        for (int r = 0; r < MAXR; ++r) { const int i = lane + 32 * r; xa[r] = (i < m) ? A[i + (long)ca * ld] : 0.0; xb[r] = (i < m) ? A[i + (long)cb * ld] : 0.0; }
#pragma unroll
        for (int r = 0; r < MAXR; ++r) { const int i = lane + 32 * r; if (i < m) { A[i + (long)ca * ld] = xb[r]; A[i + (long)cb * ld] = xa[r]; } }
      }
      if (lane == 0) { const int t = ipv[ca]; ipv[ca] = ipv[cb]; ipv[cb] = t; vn[ca] = vn[cb]; }
    }
    __syncthreads();
    QRP_ACC(1)
    // ---- (3) the LIVE rows (k0 .. m - 1) of the panel columns -> registers: warp w owns slots CPW w .. CPW w + CPW - 1, register block r holds rows
    // k0 + 32 r + lane; exact norms.  (Rows above k0 stay in global memory; they only move with the in-panel permutation, step 5.)
    double cr[CPW][MAXR];
    const int sb = CPW * warp;
    const int nl = (mv + 31) >> 5;                    // live row blocks of this panel
    {
      double a[CPW];
#pragma unroll
      for (int h = 0; h < CPW; ++h) a[h] = 0.0;
#pragma unroll
      for (int r = 0; r < MAXR; ++r) {
        const int i = k0 + lane + 32 * r;
#pragma unroll
        for (int h = 0; h < CPW; ++h) {
          cr[h][r] = (r < nl && i < m && sb + h < nbk) ? A[i + (long)(k0 + sb + h) * ld] : 0.0;
          a[h] = fma(cr[h][r], cr[h][r], a[h]);
        }
      }
#pragma unroll
      for (int h = 0; h < CPW; ++h) { a[h] = warp_sum(a[h]); if (lane == 0) pn[sb + h] = a[h]; }      // pn: SQUARED norms (the square root is taken for the pivot only)
    }
    for (int e = tid; e < 32 * MAXR; e += nthr) v_s[e] = 0.0;
    __syncthreads();
    QRP_ACC(2)
    // ---- (4) exact column-pivoted Householder QR of the panel, columns in registers: one instantiation per number of live row blocks
#define QR2_CASE(NLV) case NLV: if constexpr (NLV <= MAXR) qr2_panel_cols<MAXR, CPW, NLV>(cr, nbk, sb, v_s, pn, tau_s, sc_s, be_s, pos_slot, s_detq); break;
    switch (nl) {
      QR2_CASE(1) QR2_CASE(2) QR2_CASE(3) QR2_CASE(4) QR2_CASE(5) QR2_CASE(6) QR2_CASE(7) QR2_CASE(8) QR2_CASE(9) QR2_CASE(10) QR2_CASE(11) QR2_CASE(12)
      QR2_CASE(13) QR2_CASE(14) QR2_CASE(15) QR2_CASE(16) QR2_CASE(17) QR2_CASE(18)
      default: break;
    }
#undef QR2_CASE
    QRP_ACC(3)
    // ---- (5) write the factored panel back at the logical positions; V panel (explicit unit lower trapezoid) -> shared memory
    int pos[CPW];
    {
      const int ps = (lane < nbk) ? pos_slot[lane] : -1;
#pragma unroll
      for (int h = 0; h < CPW; ++h) { const unsigned mk = __ballot_sync(0xffffffffu, ps == sb + h); pos[h] = mk ? (__ffs(mk) - 1) : -1; }
    }
    for (int e = tid; e < (mvp - mv) * NB; e += nthr) { const int r = mv + e % (mvp - mv), c = e / (mvp - mv); Vs[r + (long)c * ldv] = 0.0; }
    for (int e = tid; e < mvp * (NB - nbk); e += nthr) { const int r = e % mvp, c = nbk + e / mvp; Vs[r + (long)c * ldv] = 0.0; }
    double sch[CPW], beh[CPW];                            // the finished columns are still raw below their pivot row: v = scal x, R(pos, pos) = beta
#pragma unroll
    for (int h = 0; h < CPW; ++h) { sch[h] = (sb + h < nbk) ? sc_s[sb + h] : 0.0; beh[h] = (sb + h < nbk) ? be_s[sb + h] : 0.0; }
    // live rows: from the registers to the logical positions, with the scaling of the reflector entries and beta on the diagonal
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int rr = lane + 32 * r;                       // local row
      if (r < nl && k0 + rr < m) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) if (pos[h] >= 0) {
          const double val = (rr > pos[h]) ? cr[h][r] * sch[h] : ((rr == pos[h]) ? beh[h] : cr[h][r]);
          A[(k0 + rr) + (long)(k0 + pos[h]) * ld] = val;
          Vs[rr + (long)pos[h] * ldv] = (rr > pos[h]) ? val : ((rr == pos[h]) ? 1.0 : 0.0);
        }
      }
    }
    // rows above k0 (entries of R computed by earlier panels) follow the in-panel permutation: four row blocks at a time through registers
    for (int rb0 = 0; rb0 < (k0 >> 5); rb0 += 4) {
      double up[CPW][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * (rb0 + t);
#pragma unroll
        for (int h = 0; h < CPW; ++h) up[h][t] = (i < k0 && pos[h] >= 0) ? A[i + (long)(k0 + sb + h) * ld] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * (rb0 + t);
#pragma unroll
        for (int h = 0; h < CPW; ++h) if (i < k0 && pos[h] >= 0) A[i + (long)(k0 + pos[h]) * ld] = up[h][t];
      }
      __syncthreads();
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < CPW; ++h) if (pos[h] >= 0) lista[pos[h]] = ipv[k0 + sb + h];
    }
    if (tid < nbk) tau[k0 + tid] = tau_s[tid];
    if (tid >= nbk && tid < NB) tau_s[tid] = 0.0;
    __syncthreads();
    if (tid < nbk) ipv[k0 + tid] = lista[tid];
    QRP_ACC(4)
    // ---- (6) Gram matrix of the reflectors, then the triangular factor T (ZLARFT forward / columnwise) by one warp
    blk_gemm<1, 0>(NB, NB, mvp, Vs, ldv, Vs, ldv, Gs, LDW, 1.0);
    __syncthreads();
    if (warp == 0) {
      // T (ZLARFT, forward / columnwise): T(:, j) = -tau_j T(:, 0:j) G(0:j, j), lane i owns row i.  Right-looking form: as soon as column l is final its
      // contribution T(i, l) G(l, j) goes to the accumulators of ALL later columns -- independent FMAs instead of one dependent chain per column.
      double acc[NB];
#pragma unroll
      for (int l = 0; l < NB; ++l) acc[l] = 0.0;
#pragma unroll
      for (int l = 0; l < NB; ++l) {
        const double tl = tau_s[l];
        const double t = (lane < l) ? -tl * acc[l] : ((lane == l) ? tl : 0.0);
        Ts[lane + l * LDW] = t;
        Tbuf[(long)(k0 / NB) * NB * NB + lane + l * NB] = t;
#pragma unroll
        for (int jj = l + 1; jj < NB; ++jj) acc[jj] = fma(t, Gs[l + jj * LDW], acc[jj]);
      }
    }
    __syncthreads();
    QRP_ACC(5)
    // ---- (7) trailing update with exact recomputation of the remaining column norms
    apply_panel_strips<1>(A, ld, k0, m, k0 + nbk, n, Vs, ldv, Ts, nbk, Wsc, vn);
    __syncthreads();
    QRP_ACC(6)
  }
  // ---- D(i) = |R(i,i)|, R(i, i:) /= D(i); phases (QDRP_decompose_mod.F90:86-100, Pivot_phase :103-126)
  int negd = 0;                                        // sign of prod R_ii
  for (int i = tid; i < kmax; i += nthr) { const double r = A[i + (long)i * ld], x = fabs(r); D[i] = x; v_s[i] = 1.0 / x; negd ^= (r < 0.0) ? 1 : 0; }
  const int neg_total = __syncthreads_count(negd) & 1;
  // the upper trapezoid as one flat loop over all threads (independent, coalesced read-modify-writes)
  for (int c0 = 8 * warp; c0 < n; c0 += 8 * nw) {        // eight columns per warp and pass: all loads before the first store
    for (int i0 = 0; i0 < kmax && i0 <= c0 + 7; i0 += 32) {
      const int i = i0 + lane; double x[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) x[t] = (c0 + t < n && i < kmax && i <= c0 + t) ? A[i + (long)(c0 + t) * ld] : 0.0;
      const double sc = (i < kmax) ? v_s[i] : 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) if (c0 + t < n && i < kmax && i <= c0 + t) A[i + (long)(c0 + t) * ld] = x[t] * sc;
    }
  }
  for (int c = tid; c < n; c += nthr) jpvt[c] = ipv[c];
  // permutation parity = (-1)^(n - number of cycles): every thread walks the cycle of its element; the smallest element of a cycle counts it
  int leaders = 0;
  for (int i = tid; i < n; i += nthr) { int mn = i, x = ipv[i]; while (x != i) { mn = min(mn, x); x = ipv[x]; } leaders += (mn == i) ? 1 : 0; }
  __shared__ int s_cyc;
  if (tid == 0) s_cyc = 0;
  __syncthreads();
  leaders += __shfl_xor_sync(0xffffffffu, leaders, 16); leaders += __shfl_xor_sync(0xffffffffu, leaders, 8); leaders += __shfl_xor_sync(0xffffffffu, leaders, 4);
  leaders += __shfl_xor_sync(0xffffffffu, leaders, 2); leaders += __shfl_xor_sync(0xffffffffu, leaders, 1);
  if (lane == 0 && leaders) atomicAdd(&s_cyc, leaders);
  __syncthreads();
  if (tid == 0) { out[b].perm_sign = ((n - s_cyc) & 1) ? -1.0 : 1.0; out[b].diag_phase = cplx(neg_total ? -1.0 : 1.0, 0.0); out[b].detq = cplx(s_detq[0], s_detq[1]); }
  QRP_ACC(7)
}

// ------------------------------------------------------------------------------------------------------------------------
// k_apply_q2: X <- Q^T X (MODE 0: panels forward with T^T; ZUNMQR 'L','C') or X <- Q X (MODE 1: panels backward with T;
// ZUNMQR 'L','N'; with X = 1 this is ZUNGQR).  grid = (column groups of cols_per_cta, batch), 256 threads, two CTAs per SM.
// ------------------------------------------------------------------------------------------------------------------------
template <int MODE, int IDENT, int NTHR>      // NTHR = 256: two CTAs per SM (m <= 288); 512: one CTA per SM for taller matrices
__global__ void __launch_bounds__(NTHR, NTHR == 256 ? 2 : 1) k_apply_q2(const double* __restrict__ QR, int m, int n, int ld, long sQ, const double* __restrict__ Tbuf, long sT,
                                                  double* __restrict__ X, int ldx, long sX, int ncols, int cols_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NB = QR2_NB, LDW = QR2_LDW;
  const int b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
  QR += (long)b * sQ; Tbuf += (long)b * sT; X += (long)b * sX;
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  double* Vs = reinterpret_cast<double*>(smem_raw);
  double* Ts = Vs + (long)ldv * NB;
  double* Wsc = Ts + LDW * NB;
  const int cb = blockIdx.x * cols_per_cta, ce = min(ncols, cb + cols_per_cta);
  const int kmax = (m < n) ? m : n, npan = (kmax + NB - 1) / NB;
  for (int pp = 0; pp < npan; ++pp) {
    const int pi = MODE ? (npan - 1 - pp) : pp, k0 = pi * NB, nbk = min(NB, kmax - k0);
    int c_lo = cb;
    if (IDENT) c_lo = max(cb, k0 & ~7);
    if (c_lo >= ce) continue;
    load_vpanel<double>(QR, ld, m, k0, nbk, NB, Vs, ldv);
    for (int e = tid; e < LDW * NB; e += nthr) { const int r = e % LDW, c = e / LDW; Ts[e] = (r < NB) ? Tbuf[(long)pi * NB * NB + r + (long)c * NB] : 0.0; }
    __syncthreads();
    apply_panel_strips<MODE ? 0 : 1>(X, ldx, k0, m, c_lo, ce, Vs, ldv, Ts, nbk, Wsc, nullptr);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_apply_q2p: k_apply_q2 with the reflector panel DOUBLE-BUFFERED in shared memory: the V and T factors of panel p + 1 are
// requested with cp.async (LDGSTS, no register staging) before the strips of panel p are worked on, so their L2 / HBM latency
// hides behind the tensor-core work (round 2's ncu source view of k_apply_q2: 31 % of the warp-stall samples were long-scoreboard
// waits on the panel load).  One CTA of 16 warps per SM, 128 columns of X per CTA.  Used when two panels fit (m <= ~330).
// ------------------------------------------------------------------------------------------------------------------------
static size_t applyq2p_smem(int m, int nw) {
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  return sizeof(double) * (2 * ((size_t)ldv * QR2_NB + QR2_LDW * QR2_NB) + (size_t)nw * QR2_WSC) + 64;
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc) : "memory");
}
template <int MODE, int IDENT>
__global__ void __launch_bounds__(512, 1) k_apply_q2p(const double* __restrict__ QR, int m, int n, int ld, long sQ, const double* __restrict__ Tbuf, long sT,
                                                    double* __restrict__ X, int ldx, long sX, int ncols, int cols_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NB = QR2_NB, LDW = QR2_LDW;
  const int b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  QR += (long)b * sQ; Tbuf += (long)b * sT; X += (long)b * sX;
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  double* Vb[2]; double* Tb[2];
  Vb[0] = reinterpret_cast<double*>(smem_raw); Vb[1] = Vb[0] + (long)ldv * NB;
  Tb[0] = Vb[1] + (long)ldv * NB; Tb[1] = Tb[0] + LDW * NB;
  double* Wsc = Tb[1] + LDW * NB;
  const int cb = blockIdx.x * cols_per_cta, ce = min(ncols, cb + cols_per_cta);
  const int kmax = (m < n) ? m : n, npan = (kmax + NB - 1) / NB;
  auto panel_of = [&](int pp, int& pi, int& k0, int& nbk, int& c_lo) {
    pi = MODE ? (npan - 1 - pp) : pp; k0 = pi * NB; nbk = min(NB, kmax - k0);
    c_lo = cb; if (IDENT) c_lo = max(cb, k0 & ~7);
  };
  auto issue = [&](int pp, int buf) {          // raw copy: rows k0 .. m-1 of the panel's columns (R entries on and above the diagonal are fixed up on arrival)
    int pi, k0, nbk, c_lo; panel_of(pp, pi, k0, nbk, c_lo);
    const int mv = m - k0;
    for (int c = warp; c < nbk; c += nw) { const double* src = QR + k0 + (long)(k0 + c) * ld; double* dst = Vb[buf] + (long)c * ldv; for (int r = lane; r < mv; r += 32) cp_async8(dst + r, src + r); }
    for (int e = tid; e < NB * NB; e += nthr) cp_async8(Tb[buf] + (e & 31) + (e >> 5) * LDW, Tbuf + (long)pi * NB * NB + e);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int e = tid; e < 2 * LDW * NB; e += nthr) { const int r = (e % (LDW * NB)) % LDW; if (r >= NB) Tb[0][e] = 0.0; }      // pad rows 32 .. 35 of both T buffers (Tb[1] follows Tb[0])
  int first = 0;
  for (; first < npan; ++first) { int pi, k0, nbk, c_lo; panel_of(first, pi, k0, nbk, c_lo); if (c_lo < ce) break; }
  if (first >= npan) return;
  issue(first, 0);
  for (int pp = first; pp < npan; ++pp) {
    const int buf = (pp - first) & 1;
    int pi, k0, nbk, c_lo; panel_of(pp, pi, k0, nbk, c_lo);
    const int mv = m - k0, mvp = (mv + 7) & ~7;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double* Vs = Vb[buf];
    for (int e = tid; e < NB * NB; e += nthr) { const int r = e & 31, c = e >> 5; if (c < nbk && r <= c && r < mvp) Vs[r + (long)c * ldv] = (r == c) ? 1.0 : 0.0; }
    for (int e = tid; e < (mvp - mv) * NB; e += nthr) { const int r = mv + e % (mvp - mv), c = e / (mvp - mv); Vs[r + (long)c * ldv] = 0.0; }
    if (nbk < NB) for (int e = tid; e < mvp * (NB - nbk); e += nthr) { const int r = e % mvp, c = nbk + e / mvp; Vs[r + (long)c * ldv] = 0.0; }
    __syncthreads();
    if (pp + 1 < npan) issue(pp + 1, buf ^ 1);
    if (c_lo < ce) apply_panel_strips<MODE ? 0 : 1>(X, ldx, k0, m, c_lo, ce, Vs, ldv, Tb[buf], nbk, Wsc, nullptr);
    __syncthreads();
  }
}
