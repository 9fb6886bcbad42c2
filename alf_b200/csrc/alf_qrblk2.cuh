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
#define QR2_LDW 36
#define QR2_WSC (QR2_LDW * 8)      // per-warp scratch: 32 x 8 block, leading dimension 36

// nw = warps per CTA.  The Gram matrix of the panel aliases the per-warp scratch (used in disjoint phases).
static size_t qr2_smem(int m, int n, int nw) {
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp);
  const size_t wsc = (size_t)nw * QR2_WSC > (size_t)QR2_LDW * QR2_NB ? (size_t)nw * QR2_WSC : (size_t)QR2_LDW * QR2_NB;
  return sizeof(double) * ((size_t)ldv * QR2_NB + QR2_LDW * QR2_NB + wsc + mp + 2 * QR2_NB + n) + sizeof(int) * (4 * (size_t)n + 64) + 64;
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
  double* tau_s = v_s + mp;
  double* pn = tau_s + NB;
  double* vn = pn + NB;
  int* ipv = reinterpret_cast<int*>(vn + n);
  int* rank_s = ipv + n;
  int* lista = rank_s + n;
  int* listb = lista + n;
  int* pos_slot = listb + n;          // [32] slot at logical panel position j; [32..63] scratch
  __shared__ int s_na;
  __shared__ double s_detq[2];

  for (int c = warp; c < n; c += nw) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) { const double x = A[i + (long)c * ld]; s = fma(x, x, s); }
    s = warp_sum(s);
    if (lane == 0) { vn[c] = sqrt(s); ipv[c] = c; }
  }
  if (tid == 0) { s_detq[0] = 1.0; s_detq[1] = 0.0; }
  __syncthreads();
  const int kmax = n;                 // m >= n
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
      for (int r = lane; r < m; r += 32) { const double x = A[r + (long)ca * ld], y = A[r + (long)cb * ld]; A[r + (long)ca * ld] = y; A[r + (long)cb * ld] = x; }
      if (lane == 0) { const int t = ipv[ca]; ipv[ca] = ipv[cb]; ipv[cb] = t; vn[ca] = vn[cb]; }
    }
    __syncthreads();
    // ---- (3) panel columns -> registers: warp w owns slots CPW w .. CPW w + CPW - 1; exact norms below row k0
    double cr[CPW][MAXR];
    const int sb = CPW * warp;
    {
      double a[CPW];
#pragma unroll
      for (int h = 0; h < CPW; ++h) a[h] = 0.0;
#pragma unroll
      for (int r = 0; r < MAXR; ++r) {
        const int i = lane + 32 * r;
#pragma unroll
        for (int h = 0; h < CPW; ++h) {
          cr[h][r] = (i < m && sb + h < nbk) ? A[i + (long)(k0 + sb + h) * ld] : 0.0;
          if (i >= k0) a[h] = fma(cr[h][r], cr[h][r], a[h]);
        }
      }
#pragma unroll
      for (int h = 0; h < CPW; ++h) { a[h] = warp_sum(a[h]); if (lane == 0) pn[sb + h] = sqrt(a[h]); }
    }
    __syncthreads();
    // ---- (4) exact column-pivoted Householder QR of the panel, columns in registers
    unsigned used = 0u;
    for (int j = 0; j < nbk; ++j) {
      const int prow = k0 + j;
      int p;
      {
        double best = (lane < nbk && !((used >> lane) & 1u)) ? pn[lane] : -1.0; int bi = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        p = bi;
      }
      used |= 1u << p;
      if (tid == 0) pos_slot[j] = p;
      if (warp == p / CPW) {
        const int hp = p % CPW;
        // The 2-norm of the pivot column below row prow - 1 is already known (pn, exact, from the previous step's update), so
        // beta = -sign(alpha) pn needs no reduction on the critical path.  pn == 0: H = I (ZLARFG's xnorm == 0, alpha == 0 case).
        double al = 0.0;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          const int i = lane + 32 * r;
          double x = cr[0][r];
#pragma unroll
          for (int h = 1; h < CPW; ++h) if (h == hp) x = cr[h][r];
          if (i == prow) al = x;
        }
        const double alpha = __shfl_sync(0xffffffffu, al, prow & 31);
        const double cn = pn[p];
        double tj, scal, beta;
        if (cn == 0.0) { tj = 0.0; scal = 0.0; beta = alpha; }
        else { beta = -copysign(cn, alpha); tj = (beta - alpha) / beta; scal = 1.0 / (alpha - beta); }
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          if (32 * r + 31 < prow) continue;
          const int i = lane + 32 * r;
          if (i < m && i >= prow) {
            double x = cr[0][r];
#pragma unroll
            for (int h = 1; h < CPW; ++h) if (h == hp) x = cr[h][r];
            if (i > prow) { x *= scal; v_s[i] = x; } else { x = beta; v_s[i] = 1.0; }
#pragma unroll
            for (int h = 0; h < CPW; ++h) if (h == hp) cr[h][r] = x;
          }
        }
        if (lane == 0) {
          tau_s[j] = tj;
          if (tj != 0.0) { s_detq[0] = -s_detq[0]; s_detq[1] = -s_detq[1]; }      // det of a real reflector = -1 (Prog/cgr1_mod.F90:338-347)
        }
      }
      __syncthreads();
      {
        const double tj = tau_s[j];
        bool doh[CPW]; bool any = false;
#pragma unroll
        for (int h = 0; h < CPW; ++h) { doh[h] = (sb + h < nbk) && !((used >> (sb + h)) & 1u); any = any || doh[h]; }
        if (any) {
          // rows above the pivot row are finished: row blocks r with 32 r + 31 < prow are skipped by the whole warp (on average half of
          // them); the dot products and norms run as two interleaved chains (even / odd row blocks)
          double w[CPW], wb[CPW], qn[CPW], qb[CPW];
#pragma unroll
          for (int h = 0; h < CPW; ++h) { w[h] = 0.0; wb[h] = 0.0; qn[h] = 0.0; qb[h] = 0.0; }
          if (tj != 0.0) {
#pragma unroll
            for (int r = 0; r < MAXR; ++r) {
              if (32 * r + 31 < prow) continue;
              const int i = lane + 32 * r;
              if (i >= prow && i < m) {
                const double v = v_s[i];
#pragma unroll
                for (int h = 0; h < CPW; ++h) { if (r & 1) wb[h] = fma(v, cr[h][r], wb[h]); else w[h] = fma(v, cr[h][r], w[h]); }
              }
            }
#pragma unroll
            for (int h = 0; h < CPW; ++h) w[h] += wb[h];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
              for (int h = 0; h < CPW; ++h) w[h] += __shfl_xor_sync(0xffffffffu, w[h], o);
#pragma unroll
            for (int h = 0; h < CPW; ++h) w[h] *= tj;
          }
#pragma unroll
          for (int r = 0; r < MAXR; ++r) {
            if (32 * r + 31 < prow) continue;
            const int i = lane + 32 * r;
            if (i >= prow && i < m) {
              const double v = (tj != 0.0) ? v_s[i] : 0.0;
#pragma unroll
              for (int h = 0; h < CPW; ++h) {
                if (doh[h]) cr[h][r] = fma(-v, w[h], cr[h][r]);
                if (i > prow) { if (r & 1) qb[h] = fma(cr[h][r], cr[h][r], qb[h]); else qn[h] = fma(cr[h][r], cr[h][r], qn[h]); }
              }
            }
          }
#pragma unroll
          for (int h = 0; h < CPW; ++h) qn[h] += qb[h];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int h = 0; h < CPW; ++h) qn[h] += __shfl_xor_sync(0xffffffffu, qn[h], o);
          if (lane == 0) {
#pragma unroll
            for (int h = 0; h < CPW; ++h) if (doh[h]) pn[sb + h] = sqrt(qn[h]);
          }
        }
      }
      __syncthreads();
    }
    // ---- (5) write the factored panel back at the logical positions; V panel (explicit unit lower trapezoid) -> shared memory
    int pos[CPW];
    {
      const int ps = (lane < nbk) ? pos_slot[lane] : -1;
#pragma unroll
      for (int h = 0; h < CPW; ++h) { const unsigned mk = __ballot_sync(0xffffffffu, ps == sb + h); pos[h] = mk ? (__ffs(mk) - 1) : -1; }
    }
    for (int e = tid; e < (mvp - mv) * NB; e += nthr) { const int r = mv + e % (mvp - mv), c = e / (mvp - mv); Vs[r + (long)c * ldv] = 0.0; }
    for (int e = tid; e < mvp * (NB - nbk); e += nthr) { const int r = e % mvp, c = nbk + e / mvp; Vs[r + (long)c * ldv] = 0.0; }
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int i = lane + 32 * r;
      if (i < m) {
#pragma unroll
        for (int h = 0; h < CPW; ++h) if (pos[h] >= 0) {
          A[i + (long)(k0 + pos[h]) * ld] = cr[h][r];
          if (i >= k0) { const int rr = i - k0; Vs[rr + (long)pos[h] * ldv] = (rr > pos[h]) ? cr[h][r] : ((rr == pos[h]) ? 1.0 : 0.0); }
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < CPW; ++h) if (pos[h] >= 0) lista[pos[h]] = ipv[k0 + sb + h];
    }
    if (tid < nbk) tau[k0 + tid] = tau_s[tid];
    if (tid >= nbk && tid < NB) tau_s[tid] = 0.0;
    __syncthreads();
    if (tid < nbk) ipv[k0 + tid] = lista[tid];
    // ---- (6) Gram matrix of the reflectors, then the triangular factor T (ZLARFT forward / columnwise) by one warp
    blk_gemm<1, 0>(NB, NB, mvp, Vs, ldv, Vs, ldv, Gs, LDW, 1.0);
    __syncthreads();
    if (warp == 0) {
      double trow[NB];
#pragma unroll
      for (int l = 0; l < NB; ++l) trow[l] = 0.0;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const double tj = tau_s[j];
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int l = 0; l < j; ++l) { if (l & 1) x1 = fma(trow[l], Gs[l + j * LDW], x1); else x0 = fma(trow[l], Gs[l + j * LDW], x0); }
        const double t = (lane < j) ? -tj * (x0 + x1) : ((lane == j) ? tj : 0.0);
        trow[j] = t;
        Ts[lane + j * LDW] = t;
        Tbuf[(long)(k0 / NB) * NB * NB + lane + j * NB] = t;
      }
    }
    __syncthreads();
    // ---- (7) trailing update with exact recomputation of the remaining column norms
    apply_panel_strips<1>(A, ld, k0, m, k0 + nbk, n, Vs, ldv, Ts, nbk, Wsc, vn);
    __syncthreads();
  }
  // ---- D(i) = |R(i,i)|, R(i, i:) /= D(i); phases (QDRP_decompose_mod.F90:86-100, Pivot_phase :103-126)
  for (int i = tid; i < kmax; i += nthr) { const double x = fabs(A[i + (long)i * ld]); D[i] = x; v_s[i] = x; }
  __syncthreads();
  for (int c = warp; c < n; c += nw) {
    const int top = min(c, kmax - 1);
    for (int i = lane; i <= top; i += 32) A[i + (long)c * ld] = A[i + (long)c * ld] * (1.0 / v_s[i]);
  }
  for (int c = tid; c < n; c += nthr) jpvt[c] = ipv[c];
  __syncthreads();
  if (warp == 0) {
    // sign of prod R_ii (after the scaling R_ii = +-1)
    int neg = 0;
    for (int i = lane; i < kmax; i += 32) neg ^= (A[i + (long)i * ld] < 0.0) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) neg ^= __shfl_xor_sync(0xffffffffu, neg, o);
    if (lane == 0) {
      double sg = 1.0;      // permutation parity: cycles of even length flip the sign
      for (int i = 0; i < n; ++i) rank_s[i] = 0;
      for (int i = 0; i < n; ++i) if (rank_s[i] == 0) {
        int next = i, L = 0;
        while (rank_s[next] == 0) { ++L; rank_s[next] = 1; next = ipv[next]; }
        if ((L & 1) == 0) sg = -sg;
      }
      out[b].perm_sign = sg; out[b].diag_phase = cplx(neg ? -1.0 : 1.0, 0.0); out[b].detq = cplx(s_detq[0], s_detq[1]);
    }
  }
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
