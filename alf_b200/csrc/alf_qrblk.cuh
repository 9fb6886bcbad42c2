// Blocked column-pivoted Householder QR and block-reflector application for batches of N x N (or 2N x 2N) matrices that
// do NOT fit one SM's shared memory (N = 256 real: 512 KB).  Replaces ZGEQP3 / ZUNGQR / ZUNMQR on the sweep's hot path
// (Prog/QDRP_decompose_mod.F90:78-100, Prog/udv_state_mod.F90:569-578, Prog/cgr1_mod.F90:350-445, Prog/cgr2_2_mod.F90:155-191).
//
// ZGEQP3's BLAS-2 half (one pass over the whole trailing matrix per column) is what makes a pivoted QR slow when the matrix
// lives in L2 / HBM.  Here pivoting is WINDOWED: a panel consists of the NB remaining columns of largest (exactly recomputed)
// norm; inside the panel, which sits in shared memory, pivoting is exact; the trailing matrix is then updated ONCE per panel
// with the compact-WY block reflector  (I - V T^H V^H)  on the FP64 tensor cores (DMMA m8n8k4), and the column norms of the
// remaining columns are recomputed exactly in the same pass.  The factorisation A P = Q R is a valid pivoted QR with graded
// |R_ii| (rank revealing in practice; the stabilised products and Green functions agree with full pivoting to ~1e-13, see
// tests), but the pivot ORDER can differ from ZGEQP3's, so only pivot-invariant quantities are compared with the oracle.
// Traffic per matrix: ~2 passes over the trailing matrix per panel (N/NB panels) instead of 2 per column.
#pragma once
#include "alf_la.cuh"
#include "alf_update_fast.cuh"   // dmma884

// ---- C (+)= alpha * op(A) * B on shared-memory operands, all warps of the CTA; 8 x 8 output blocks per warp.
// op(A)(i,k) = TA ? conj(A[k + i*lda]) : A[i + k*lda];  M, N multiples of 8, K multiple of 4 (callers zero-pad).
// ACC = 0: C = alpha*op(A)B ; ACC = 1: C += alpha*op(A)B.
template <int TA, int ACC>
__device__ __forceinline__ void blk_gemm(int M, int N, int K, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                         double* __restrict__ C, int ldc, double alpha) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int mb = M >> 3, nb = N >> 3;
  for (int blk = warp; blk < mb * nb; blk += nw) {
    const int i0 = (blk % mb) * 8, j0 = (blk / mb) * 8;
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    const double* bp = B + (long)(j0 + g) * ldb + q;
    if (TA) {
      const double* ap = A + (long)(i0 + g) * lda + q;
      int k0 = 0;
      for (; k0 + 8 <= K; k0 += 8) { dmma884(c0, c1, ap[k0], bp[k0]); dmma884(d0, d1, ap[k0 + 4], bp[k0 + 4]); }
      for (; k0 < K; k0 += 4) dmma884(c0, c1, ap[k0], bp[k0]);
    } else {
      const double* ap = A + (i0 + g) + (long)q * lda;
      int k0 = 0;
      for (; k0 + 8 <= K; k0 += 8) { dmma884(c0, c1, ap[(long)k0 * lda], bp[k0]); dmma884(d0, d1, ap[(long)(k0 + 4) * lda], bp[k0 + 4]); }
      for (; k0 < K; k0 += 4) dmma884(c0, c1, ap[(long)k0 * lda], bp[k0]);
    }
    c0 += d0; c1 += d1;
    double* cp = C + (i0 + g) + (long)(j0 + 2 * q) * ldc;
    if (ACC) { cp[0] += alpha * c0; cp[ldc] += alpha * c1; } else { cp[0] = alpha * c0; cp[ldc] = alpha * c1; }
  }
}
// complex: the same 8 x 8 output blocks, each complex product as four real DMMA products on split accumulators
template <int TA, int ACC>
__device__ __forceinline__ void blk_gemm(int M, int N, int K, const cplx* __restrict__ A, int lda, const cplx* __restrict__ B, int ldb,
                                         cplx* __restrict__ C, int ldc, double alpha) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int mb = M >> 3, nb = N >> 3;
  for (int blk = warp; blk < mb * nb; blk += nw) {
    const int i0 = (blk % mb) * 8, j0 = (blk / mb) * 8;
    double r0 = 0.0, r1 = 0.0, m0 = 0.0, m1 = 0.0;
    const cplx* bp = B + (long)(j0 + g) * ldb + q;
    for (int k0 = 0; k0 < K; k0 += 4) {
      cplx a = TA ? A[(long)(i0 + g) * lda + q + k0] : A[(i0 + g) + (long)(q + k0) * lda];
      if (TA) a.y = -a.y;                                 // op(A) = A^H
      const cplx b = bp[k0];
      dmma884(r0, r1, a.x, b.x); dmma884(r0, r1, -a.y, b.y);
      dmma884(m0, m1, a.x, b.y); dmma884(m0, m1, a.y, b.x);
    }
    cplx* cp = C + (i0 + g) + (long)(j0 + 2 * q) * ldc;
    if (ACC) { cp[0] = cp[0] + cplx(alpha * r0, alpha * m0); cp[ldc] = cp[ldc] + cplx(alpha * r1, alpha * m1); }
    else { cp[0] = cplx(alpha * r0, alpha * m0); cp[ldc] = cplx(alpha * r1, alpha * m1); }
  }
}

// Shared-memory footprint (bytes) of the blocked kernels for an m-row matrix: V panel, column tile, W, W2, T, reflector, norms, ints.
template <typename T>
static size_t qrblk_smem(int m, int n, int NB, int TC) {
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp), ldw = ld_pad(NB);
  return sizeof(T) * ((size_t)ldv * NB + (size_t)ldv * TC + 2 * (size_t)ldw * TC + (size_t)ldw * NB + mp + 2 * NB) + sizeof(double) * (2 * (size_t)n + 64) + sizeof(int) * (4 * (size_t)n + 64) + 256;
}

// Apply the block reflector of one panel to TC-column tiles of a global matrix X:
//   X(r0:m, cols) <- (I - V op(T) V^H) X(r0:m, cols),   op(T) = T^H (CONJT = 1: Q^H from the left) or T (CONJT = 0: Q from the left)
// Vs: mv x NB in shared memory (unit lower trapezoid made explicit, zero padded to mvp rows), Ts: NB x NB upper triangular.
// If vn != nullptr the 2-norm of X(r0 + nbk : m, c) is written to vn[c] for every processed column (exact norm recomputation).
template <typename T, int CONJT>
__device__ __forceinline__ void apply_panel(T* __restrict__ X, int ldx, int r0, int m, int c_begin, int c_end, const T* __restrict__ Vs, int ldv,
                                            const T* __restrict__ Ts, int ldw, int NB, int nbk, int TC, T* __restrict__ Cs, T* __restrict__ Ws,
                                            T* __restrict__ W2s, double* __restrict__ vn) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  const int mv = m - r0, mvp = (mv + 7) & ~7;
  for (int c0 = c_begin; c0 < c_end; c0 += TC) {
    const int nc = min(TC, c_end - c0);
    for (int e = tid; e < mvp * TC; e += nthr) {
      const int r = e % mvp, c = e / mvp;
      Cs[r + (long)c * ldv] = (r < mv && c < nc) ? X[(r0 + r) + (long)(c0 + c) * ldx] : zero_<T>();
    }
    __syncthreads();
    blk_gemm<1, 0>(NB, TC, mvp, Vs, ldv, Cs, ldv, Ws, ldw, 1.0);                  // W  = V^H C
    __syncthreads();
    if (CONJT) blk_gemm<1, 0>(NB, TC, NB, Ts, ldw, Ws, ldw, W2s, ldw, 1.0);       // W2 = T^H W
    else blk_gemm<0, 0>(NB, TC, NB, Ts, ldw, Ws, ldw, W2s, ldw, 1.0);             // W2 = T W
    __syncthreads();
    blk_gemm<0, 1>(mvp, TC, NB, Vs, ldv, W2s, ldw, Cs, ldv, -1.0);                // C -= V W2
    __syncthreads();
    for (int e = tid; e < mv * nc; e += nthr) { const int r = e % mv, c = e / mv; X[(r0 + r) + (long)(c0 + c) * ldx] = Cs[r + (long)c * ldv]; }
    if (vn) {
      for (int c = warp; c < nc; c += nw) {
        double s = 0.0;
        for (int r = nbk + lane; r < mv; r += 32) s += abs2_(Cs[r + (long)c * ldv]);
        s = warp_sum(s);
        if (lane == 0) vn[c0 + c] = sqrt(s);
      }
    }
    __syncthreads();
  }
}

// Load the reflector panel (columns k0 .. k0+nbk-1 of a QR'd matrix, rows k0 .. m-1) as an explicit unit-lower trapezoid, zero padded
template <typename T>
__device__ __forceinline__ void load_vpanel(const T* __restrict__ A, int ld, int m, int k0, int nbk, int NB, T* __restrict__ Vs, int ldv) {
  const int mv = m - k0, mvp = (mv + 7) & ~7;
  for (int e = threadIdx.x; e < mvp * NB; e += blockDim.x) {
    const int r = e % mvp, c = e / mvp;
    T v = zero_<T>();
    if (c < nbk && r < mv) { if (r > c) v = A[(k0 + r) + (long)(k0 + c) * ld]; else if (r == c) v = one_<T>(); }
    Vs[r + (long)c * ldv] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_qrp_blk: windowed column-pivoted blocked Householder QR, in place (reflectors below the diagonal, R on and above),
// then D(i) = |R(i,i)|, R(i, i:) /= D(i).  One CTA per matrix.  Tbuf receives the NB x NB triangular factors of all panels.
// ------------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512) k_qrp_blk(T* __restrict__ A, int m, int n, int ld, long sA, T* __restrict__ tau, long sTau, int* __restrict__ jpvt,
                                                 long sP, double* __restrict__ D, long sD, QrOut* __restrict__ out, T* __restrict__ Tbuf, long sT,
                                                 int NB, int TC) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  A += (long)b * sA; tau += (long)b * sTau; jpvt += (long)b * sP; D += (long)b * sD; Tbuf += (long)b * sT;
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp), ldw = ld_pad(NB);
  T* Vs = reinterpret_cast<T*>(smem_raw);
  T* Cs = Vs + (long)ldv * NB;
  T* Ws = Cs + (long)ldv * TC;
  T* W2s = Ws + (long)ldw * TC;
  T* Ts = W2s + (long)ldw * TC;
  T* v_s = Ts + (long)ldw * NB;
  T* tau_s = v_s + mp;
  T* gram = tau_s + NB;                          // NB scratch for one column of V^H v
  double* vn = reinterpret_cast<double*>(gram + NB);
  double* pn = vn + n;                           // panel norms (NB used) + scratch
  int* ipv = reinterpret_cast<int*>(pn + n + 64);
  int* rank_s = ipv + n;
  int* lista = rank_s + n;
  int* listb = lista + n;
  __shared__ int s_na, s_nb2, s_piv;
  __shared__ double s_detq[2];

  for (int c = warp; c < n; c += nw) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) s += abs2_(A[i + (long)c * ld]);
    s = warp_sum(s);
    if (lane == 0) { vn[c] = sqrt(s); ipv[c] = c; }
  }
  if (tid == 0) { s_detq[0] = 1.0; s_detq[1] = 0.0; }
  __syncthreads();
  const int kmax = (m < n) ? m : n;
  for (int k0 = 0; k0 < kmax; k0 += NB) {
    const int nbk = min(NB, kmax - k0), mv = m - k0, mvp = (mv + 7) & ~7;
    // ---- (1) the nbk remaining columns of largest norm form the panel: rank by counting (ties -> lower index first)
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
      if (lane == 0) { s_na = na; s_nb2 = nb2; }
    }
    __syncthreads();
    const int na = s_na;
    // ---- (2) stage the panel's FULL columns (rows 0..m-1: the R rows above travel with in-panel swaps) in shared memory; the
    // unselected columns of the panel range move to the vacated positions in global memory
    T* Vp = Vs + k0;                                  // the (m - k0)-row panel proper
    for (int e = tid; e < mp * NB; e += nthr) { const int r = e % mp, c = e / mp; if (c >= nbk || r >= m) Vs[r + (long)c * ldv] = zero_<T>(); }
    for (int pi = warp; pi < na; pi += nw) {
      const int ca = lista[pi], cb = listb[pi];
      for (int r = lane; r < m; r += 32) {
        const T x = A[r + (long)ca * ld], y = A[r + (long)cb * ld];
        A[r + (long)ca * ld] = y;
        Vs[r + (long)(cb - k0) * ldv] = x;
      }
      if (lane == 0) { const int t = ipv[ca]; ipv[ca] = ipv[cb]; ipv[cb] = t; vn[ca] = vn[cb]; }
    }
    for (int c = k0 + warp; c < k0 + nbk; c += nw) {
      if (rank_s[c] < nbk) for (int r = lane; r < m; r += 32) Vs[r + (long)(c - k0) * ldv] = A[r + (long)c * ld];
    }
    __syncthreads();
    // exact norms of the panel columns (rows k0..), local pivot bookkeeping
    for (int c = warp; c < nbk; c += nw) {
      double s = 0.0;
      for (int r = lane; r < mv; r += 32) s += abs2_(Vp[r + (long)c * ldv]);
      s = warp_sum(s);
      if (lane == 0) { pn[c] = sqrt(s); rank_s[k0 + c] = c; }      // rank_s reused: in-panel permutation (position -> original panel slot)
    }
    __syncthreads();
    // ---- (3) exact column-pivoted Householder QR inside the panel (shared memory)
    const int jmax = min(nbk, mv);
    for (int j = 0; j < jmax; ++j) {
      if (warp == 0) {
        double best = -1.0; int bi = j;
        for (int c = j + lane; c < nbk; c += 32) { const double x = pn[c]; if (x > best) { best = x; bi = c; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) s_piv = bi;
      }
      __syncthreads();
      const int p = s_piv;
      if (p != j) {
        for (int r = tid; r < m; r += nthr) { const T t = Vs[r + (long)p * ldv]; Vs[r + (long)p * ldv] = Vs[r + (long)j * ldv]; Vs[r + (long)j * ldv] = t; }
        if (tid == 0) { const int t = rank_s[k0 + p]; rank_s[k0 + p] = rank_s[k0 + j]; rank_s[k0 + j] = t; pn[p] = pn[j]; }
        __syncthreads();
      }
      // reflector for column j (ZLARFG convention, Libraries/libqrref/zlarfg.f:107-203)
      if (warp == 0) {
        T* col = Vp + (long)j * ldv + j;
        const int len = mv - j;
        double xn2 = 0.0;
        for (int i = 1 + lane; i < len; i += 32) xn2 += abs2_(col[i]);
        xn2 = warp_sum(xn2);
        const T alpha = col[0];
        __syncwarp();                    // every lane has read the pivot entry before lane 0 overwrites it with beta
        T tj, scal; double beta;
        if (xn2 == 0.0 && imag_(alpha) == 0.0) { tj = zero_<T>(); scal = zero_<T>(); beta = real_(alpha); }
        else {
          beta = -copysign(sqrt(abs2_(alpha) + xn2), real_(alpha));
          tj = make_<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
          scal = one_<T>() / (alpha - make_<T>(beta, 0.0));
        }
        for (int i = 1 + lane; i < len; i += 32) { const T y = col[i] * scal; col[i] = y; v_s[i] = y; }
        if (lane == 0) {
          v_s[0] = one_<T>(); col[0] = make_<T>(beta, 0.0); tau_s[j] = tj; tau[k0 + j] = tj;
          if (abs2_(tj) != 0.0) {   // det(H_j) = 1 - 2 (tau/|tau|) Re(tau/|tau|)   (Prog/cgr1_mod.F90:338-347)
            const double X = abs_(tj); const cplx z = cplx(real_(tj) / X, imag_(tj) / X);
            const cplx d = cplx(1.0, 0.0) - 2.0 * (real_(tj) / X) * z; const double ad = abs_(d);
            const cplx dq = cplx(s_detq[0], s_detq[1]) * cplx(d.x / ad, d.y / ad); s_detq[0] = dq.x; s_detq[1] = dq.y;
          }
        }
      }
      __syncthreads();
      {   // apply H_j^H to the remaining panel columns, recompute their norms below row j
        const T tauc = conj_(tau_s[j]);
        const int len = mv - j;
        for (int c = j + 1 + warp; c < nbk; c += nw) {
          T* col = Vp + (long)c * ldv + j;
          T w = zero_<T>();
          for (int i = lane; i < len; i += 32) fmac_(w, v_s[i], col[i]);
          w = warp_sum(w);
          const T s = tauc * w;
          double nrm = 0.0;
          for (int i = lane; i < len; i += 32) { const T y = col[i] - v_s[i] * s; col[i] = y; if (i > 0) nrm += abs2_(y); }
          nrm = warp_sum(nrm);
          if (lane == 0) pn[c] = sqrt(nrm);
        }
      }
      __syncthreads();
    }
    for (int j = jmax + tid; j < nbk; j += nthr) { tau_s[j] = zero_<T>(); tau[k0 + j] = zero_<T>(); }
    // ---- write the factored panel back (R on/above the diagonal, reflectors below), apply the in-panel permutation to the R rows above
    for (int e = tid; e < m * nbk; e += nthr) { const int r = e % m, c = e / m; A[r + (long)(k0 + c) * ld] = Vs[r + (long)c * ldv]; }
    __syncthreads();
    if (tid < nbk) pn[n + tid] = (double)ipv[k0 + rank_s[k0 + tid]];     // permuted pivot labels
    __syncthreads();
    if (tid < nbk) ipv[k0 + tid] = (int)pn[n + tid];
    // ---- T factor of the panel (ZLARFT forward/columnwise): T(j,j) = tau_j, T(0:j,j) = -tau_j T(0:j,0:j) V(:,0:j)^H v_j
    for (int e = tid; e < mvp * NB; e += nthr) { const int r = e % mvp, c = e / mvp; if (c < nbk && r <= c && r < mv) Vp[r + (long)c * ldv] = (r == c) ? one_<T>() : zero_<T>(); }
    for (int e = tid; e < ldw * NB; e += nthr) Ts[e] = zero_<T>();
    __syncthreads();
    for (int j = 0; j < nbk; ++j) {
      for (int i = warp; i < j; i += nw) {       // gram[i] = V(:,i)^H V(:,j)
        T s = zero_<T>();
        for (int r = j + lane; r < mv; r += 32) fmac_(s, Vp[r + (long)i * ldv], Vp[r + (long)j * ldv]);
        s = warp_sum(s);
        if (lane == 0) gram[i] = s;
      }
      __syncthreads();
      const T tj = tau_s[j];
      for (int i = tid; i < j; i += nthr) {
        T s = zero_<T>();
        for (int l = i; l < j; ++l) fma_(s, Ts[i + (long)l * ldw], gram[l]);
        Ts[i + (long)j * ldw] = -(tj * s);
      }
      if (tid == 0) Ts[j + (long)j * ldw] = tj;
      __syncthreads();
    }
    for (int e = tid; e < NB * NB; e += nthr) Tbuf[(long)(k0 / NB) * NB * NB + e] = Ts[(e % NB) + (long)(e / NB) * ldw];
    // ---- (4) trailing update with exact recomputation of the remaining column norms
    apply_panel<T, 1>(A, ld, k0, m, k0 + nbk, n, Vp, ldv, Ts, ldw, NB, nbk, TC, Cs, Ws, W2s, vn);
    __syncthreads();
  }
  // ---- D(i) = |R(i,i)|, R(i, i:) /= D(i); phases (QDRP_decompose_mod.F90:86-100, Pivot_phase :103-126)
  for (int i = tid; i < kmax; i += nthr) { const double x = abs_(A[i + (long)i * ld]); D[i] = x; pn[i] = x; }
  __syncthreads();
  for (long e = tid; e < (long)kmax * n; e += nthr) {
    const int i = (int)(e % kmax), c = (int)(e / kmax);
    if (c >= i) A[i + (long)c * ld] = A[i + (long)c * ld] * (1.0 / pn[i]);
  }
  for (int c = tid; c < n; c += nthr) jpvt[c] = ipv[c];
  __syncthreads();
  if (tid == 0) {
    cplx ph = cplx(1.0, 0.0);
    for (int i = 0; i < kmax; ++i) { const T r = A[i + (long)i * ld]; ph = ph * cplx(real_(r), imag_(r)); }
    double sg = 1.0;
    for (int i = 0; i < n; ++i) vn[i] = 0.0;
    for (int i = 0; i < n; ++i) if (vn[i] == 0.0) {
      int next = i, L = 0;
      while (vn[next] == 0.0) { ++L; vn[next] = 1.0; next = ipv[next]; }
      if ((L & 1) == 0) sg = -sg;
    }
    out[b].perm_sign = sg; out[b].diag_phase = ph; out[b].detq = cplx(s_detq[0], s_detq[1]);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_apply_q: X <- Q^H X (MODE 0: panels forward with T^H; ZUNMQR 'L','C') or X <- Q X (MODE 1: panels backward with T; ZUNMQR
// 'L','N'; with X = 1 this is ZUNGQR).  QR holds the reflectors, Tbuf the panel factors.  grid = (column groups, batch).
// IDENT = 1 (MODE 1 only): X is known to be the identity on entry, so columns left of the current panel are still untouched zeros.
// ------------------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, int IDENT>
__global__ void __launch_bounds__(512) k_apply_q(const T* __restrict__ QR, int m, int n, int ld, long sQ, const T* __restrict__ Tbuf, long sT,
                                                 T* __restrict__ X, int ldx, long sX, int ncols, int NB, int TC, int cols_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
  QR += (long)b * sQ; Tbuf += (long)b * sT; X += (long)b * sX;
  const int mp = (m + 7) & ~7, ldv = ld_pad(mp), ldw = ld_pad(NB);
  T* Vs = reinterpret_cast<T*>(smem_raw);
  T* Cs = Vs + (long)ldv * NB;
  T* Ws = Cs + (long)ldv * TC;
  T* W2s = Ws + (long)ldw * TC;
  T* Ts = W2s + (long)ldw * TC;
  const int cb = blockIdx.x * cols_per_cta, ce = min(ncols, cb + cols_per_cta);
  const int kmax = (m < n) ? m : n, npan = (kmax + NB - 1) / NB;
  for (int pp = 0; pp < npan; ++pp) {
    const int pi = MODE ? (npan - 1 - pp) : pp, k0 = pi * NB, nbk = min(NB, kmax - k0);
    int c_lo = cb;
    if (IDENT) c_lo = max(cb, k0);
    if (c_lo >= ce) continue;
    load_vpanel<T>(QR, ld, m, k0, nbk, NB, Vs, ldv);
    for (int e = tid; e < ldw * NB; e += nthr) { const int r = e % ldw, c = e / ldw; Ts[e] = (r < NB) ? Tbuf[(long)pi * NB * NB + r + (long)c * NB] : zero_<T>(); }
    __syncthreads();
    apply_panel<T, MODE ? 0 : 1>(X, ldx, k0, m, c_lo, ce, Vs, ldv, Ts, ldw, NB, nbk, TC, Cs, Ws, W2s, nullptr);
    __syncthreads();
  }
}
