// Batched dense linear algebra for the stabilisation path, hand-written for sm_100a (no cuBLAS/cuSOLVER):
//   k_gemm      -- C = op(A) op(B), register-tiled FP64 (replaces ZGEMM/ZTRMM/MMULT call sites,
//                  Prog/cgr1_mod.F90:223-225,396,445; Prog/udv_state_mod.F90:569-572)
//   k_qrp       -- column-pivoted Householder QR + split-off of D (replaces ZGEQP3 + QDRP_decompose,
//                  Prog/QDRP_decompose_mod.F90:60-101; Pivot_phase :103-126)
//   k_formq     -- explicit Q from the reflectors (replaces ZUNGQR, Prog/udv_state_mod.F90:576)
//   k_trsm_lun  -- X = R^-1 B for upper-triangular R (replaces ZTRSM, Prog/cgr1_mod.F90:381)
// One CTA works on one matrix (QR/formQ) or one tile / column panel of one matrix (GEMM/TRSM);
// the batch (chains x flavors) is the grid's second dimension.
#pragma once
#include <type_traits>
#include "alf_types.cuh"

// ------------------------------------------------------------------------------------------------
// GEMM: C[b] = op(A[b]) * op(B[b]);  TA/TB: 0 = as stored, 1 = conjugate transpose.
// MASK: 1 -> A is upper triangular as stored (strict lower part read as 0), 2 -> same for B.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
// ------------------------------------------------------------------------------------------------
#define GEMM_BM 64
#define GEMM_BN 64
#define GEMM_BK 16
#define GEMM_PAD 4

template <typename T, int TA, int TB, int MASK>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const T* __restrict__ A, int lda, long sA,
                                              const T* __restrict__ B, int ldb, long sB, T* __restrict__ C, int ldc, long sC) {
  __shared__ T As[GEMM_BK][GEMM_BM + GEMM_PAD];
  __shared__ T Bs[GEMM_BK][GEMM_BN + GEMM_PAD];
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
  const int b = blockIdx.y;
  A += (long)b * sA; B += (long)b * sB; C += (long)b * sC;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = tm * GEMM_BM, n0 = tn * GEMM_BN;
  constexpr bool kDmma = std::is_same<T, double>::value;     // real: FP64 tensor cores, warp tile 16 x 32 (2 x 4 fragments)
  constexpr bool kCplxDmma = !kDmma;                          // complex: the same tiling on split real / imaginary accumulators
  double cre[2][4][2], cim[2][4][2];
#pragma unroll
  for (int ia = 0; ia < 2; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) { cre[ia][ib][0] = cre[ia][ib][1] = 0.0; cim[ia][ib][0] = cim[ia][ib][1] = 0.0; }
  const int lane = tid & 31, warp = tid >> 5, fg = lane >> 2, fq = lane & 3;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = zero_<T>();

  for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
    // ---- load A tile: As[k][m] = op(A)(m0+m, k0+k)
    if (TA == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int m = tid & 63, k = (tid >> 6) + 4 * r;
        int gm = m0 + m, gk = k0 + k;
        T v = zero_<T>();
        if (gm < M && gk < K && !(MASK == 1 && gm > gk)) v = A[gm + (long)gk * lda];
        As[k][m] = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int k = tid & 15, m = (tid >> 4) + 16 * r;
        int gm = m0 + m, gk = k0 + k;
        T v = zero_<T>();
        if (gm < M && gk < K && !(MASK == 1 && gk > gm)) v = conj_(A[gk + (long)gm * lda]);
        As[k][m] = v;
      }
    }
    // ---- load B tile: Bs[k][n] = op(B)(k0+k, n0+n)
    if (TB == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int k = tid & 15, n = (tid >> 4) + 16 * r;
        int gk = k0 + k, gn = n0 + n;
        T v = zero_<T>();
        if (gk < K && gn < N && !(MASK == 2 && gk > gn)) v = B[gk + (long)gn * ldb];
        Bs[k][n] = v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int n = tid & 63, k = (tid >> 6) + 4 * r;
        int gk = k0 + k, gn = n0 + n;
        T v = zero_<T>();
        if (gk < K && gn < N && !(MASK == 2 && gn > gk)) v = conj_(B[gn + (long)gk * ldb]);
        Bs[k][n] = v;
      }
    }
    __syncthreads();
    if constexpr (kCplxDmma) {
      // complex product as four real DMMA products per fragment pair: re += ar br - ai bi ; im += ar bi + ai br
#pragma unroll
      for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
        cplx a[2], bb[4];
#pragma unroll
        for (int ia = 0; ia < 2; ++ia) a[ia] = As[4 * ks + fq][wm + 8 * ia + fg];
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) bb[ib] = Bs[4 * ks + fq][wn + 8 * ib + fg];
#pragma unroll
        for (int ia = 0; ia < 2; ++ia)
#pragma unroll
          for (int ib = 0; ib < 4; ++ib) {
            dmma884(cre[ia][ib][0], cre[ia][ib][1], a[ia].x, bb[ib].x); dmma884(cre[ia][ib][0], cre[ia][ib][1], -a[ia].y, bb[ib].y);
            dmma884(cim[ia][ib][0], cim[ia][ib][1], a[ia].x, bb[ib].y); dmma884(cim[ia][ib][0], cim[ia][ib][1], a[ia].y, bb[ib].x);
          }
      }
    } else if constexpr (kDmma) {
      // acc[2*ia + (ib >> 1)][2 * (ib & 1) + h] <-> C(wm + 8 ia + fg, wn + 8 ib + 2 fq + h)
#pragma unroll
      for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
        double a[2], bb[4];
#pragma unroll
        for (int ia = 0; ia < 2; ++ia) a[ia] = As[4 * ks + fq][wm + 8 * ia + fg];
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) bb[ib] = Bs[4 * ks + fq][wn + 8 * ib + fg];
#pragma unroll
        for (int ia = 0; ia < 2; ++ia)
#pragma unroll
          for (int ib = 0; ib < 4; ++ib) dmma884(acc[2 * ia + (ib >> 1)][2 * (ib & 1)], acc[2 * ia + (ib >> 1)][2 * (ib & 1) + 1], a[ia], bb[ib]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < GEMM_BK; ++kk) {
        T a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][ty * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) fma_(acc[i][j], a[i], bb[j]);
      }
    }
    __syncthreads();
  }
  if constexpr (kCplxDmma) {
#pragma unroll
    for (int ia = 0; ia < 2; ++ia)
#pragma unroll
      for (int ib = 0; ib < 4; ++ib)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int gm = m0 + wm + 8 * ia + fg, gn = n0 + wn + 8 * ib + 2 * fq + h;
          if (gm < M && gn < N) C[gm + (long)gn * ldc] = make_<T>(cre[ia][ib][h], cim[ia][ib][h]);
        }
  } else if constexpr (kDmma) {
#pragma unroll
    for (int ia = 0; ia < 2; ++ia)
#pragma unroll
      for (int ib = 0; ib < 4; ++ib)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int gm = m0 + wm + 8 * ia + fg, gn = n0 + wn + 8 * ib + 2 * fq + h;
          if (gm < M && gn < N) C[gm + (long)gn * ldc] = acc[2 * ia + (ib >> 1)][2 * (ib & 1) + h];
        }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + ty * 4 + j;
      if (gn >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int gm = m0 + tx * 4 + i;
        if (gm < M) C[gm + (long)gn * ldc] = acc[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Householder helpers.  A warp owns whole columns: lane l holds rows r0+l, r0+l+32, ... in registers,
// so one reflector application reads and writes each trailing column exactly once.
// C(r0:m, c) -= tauc * v * (v^H C(r0:m, c)),  v in shared memory with v[0] = 1.
// If vn != nullptr the exact 2-norm of C(r0+1:m, c) is written to vn[c] (partial column norm for
// the next pivot search; LAPACK down-dates and occasionally recomputes, we always recompute because
// the column is already in registers).
// ------------------------------------------------------------------------------------------------
template <typename T, int MAXR>
__device__ __forceinline__ void reflect_cols(T* __restrict__ A, int ld, int r0, int m, int c_begin, int c_end,
                                             const T* __restrict__ v, T tauc, double* __restrict__ vn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int len = m - r0;
  T vv[MAXR];
#pragma unroll
  for (int q = 0; q < MAXR; ++q) { int i = lane + 32 * q; vv[q] = (i < len) ? v[i] : zero_<T>(); }
  for (int c = c_begin + warp; c < c_end; c += nwarps) {
    T* col = A + (long)c * ld + r0;
    T x[MAXR];
    T w = zero_<T>();
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
      int i = lane + 32 * q;
      x[q] = (i < len) ? col[i] : zero_<T>();
      fmac_(w, vv[q], x[q]);
    }
    w = warp_sum(w);
    T s = tauc * w;
    double nrm = 0.0;
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
      int i = lane + 32 * q;
      if (i < len) {
        T y = x[q] - vv[q] * s;
        col[i] = y;
        if (i > 0) nrm += abs2_(y);
      }
    }
    if (vn) {
      nrm = warp_sum(nrm);
      if (lane == 0) vn[c] = sqrt(nrm);
    }
  }
}

struct QrOut {           // per-matrix scalars produced by k_qrp
  double perm_sign;      // parity of the pivot permutation (Pivot_phase)
  cplx diag_phase;       // prod_i R_ii/|R_ii|
  cplx detq;             // prod_i det(H_i) = prod_i (1 - tau_i v_i^H v_i)
};

// Column-pivoted Householder QR of an m x n matrix (m >= n), in place: reflectors below the diagonal,
// R on and above it, then D(i) = |R(i,i)| and R(i,i:) /= D(i)  (QDRP_decompose_mod.F90:86-100).
// jpvt[j] = original index of the column now at position j (0-based; LAPACK's JPVT minus 1).
// If PIVOT == 0 the columns are taken in order (IPVT /= 0 "fixed" columns of ZGEQP3).
template <typename T, int MAXR, int PIVOT>
__device__ void qrp_device(T* __restrict__ A, int m, int n, int ld, T* __restrict__ tau, int* __restrict__ jpvt,
                           double* __restrict__ D, QrOut* out, T* __restrict__ v_s, double* __restrict__ vn_s,
                           int* __restrict__ ipv_s, double* __restrict__ red_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, nthr = blockDim.x;
  // initial column norms, warp per column
  for (int c = warp; c < n; c += nwarps) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) s += abs2_(A[i + (long)c * ld]);
    s = warp_sum(s);
    if (lane == 0) { vn_s[c] = sqrt(s); ipv_s[c] = c; }
  }
  __syncthreads();
  const int kmax = (m < n) ? m : n;
  cplx detq = cplx(1.0, 0.0);
  for (int j = 0; j < kmax; ++j) {
    // ---- pivot: first index of the maximal partial norm in [j, n)
    if (PIVOT) {
      if (warp == 0) {
        double best = -1.0; int bi = j;
        for (int c = j + lane; c < n; c += 32) { double x = vn_s[c]; if (x > best) { best = x; bi = c; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) red_s[0] = (double)bi;
      }
      __syncthreads();
      const int p = (int)red_s[0];
      if (p != j) {
        for (int i = tid; i < m; i += nthr) { T t = A[i + (long)p * ld]; A[i + (long)p * ld] = A[i + (long)j * ld]; A[i + (long)j * ld] = t; }
        if (tid == 0) { int t = ipv_s[p]; ipv_s[p] = ipv_s[j]; ipv_s[j] = t; vn_s[p] = vn_s[j]; }
      }
      __syncthreads();
    }
    // ---- generate the reflector for column j (ZLARFG, Libraries/libqrref/zlarfg.f:107-203; no safmin rescaling loop)
    if (warp == 0) {
      T* col = A + (long)j * ld + j;
      const int len = m - j;
      double xn2 = 0.0;
      for (int i = 1 + lane; i < len; i += 32) xn2 += abs2_(col[i]);
      xn2 = warp_sum(xn2);
      T alpha = col[0];
      __syncwarp();                      // every lane has read the pivot entry before lane 0 overwrites it with beta (racecheck)
      T tj, scal; double beta;
      if (xn2 == 0.0 && imag_(alpha) == 0.0) { tj = zero_<T>(); scal = zero_<T>(); beta = real_(alpha); }
      else {
        beta = -copysign(sqrt(abs2_(alpha) + xn2), real_(alpha));
        tj = make_<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta);
        scal = one_<T>() / (alpha - make_<T>(beta, 0.0));
      }
      for (int i = 1 + lane; i < len; i += 32) { T y = col[i] * scal; col[i] = y; v_s[i] = y; }
      if (lane == 0) { v_s[0] = one_<T>(); col[0] = make_<T>(beta, 0.0); tau[j] = tj; }
      if (abs2_(tj) != 0.0) {  // det(H_j) = 1 - 2 (tau/|tau|) (Re tau/|tau|)   (Prog/cgr1_mod.F90:338-347)
        double X = abs_(tj); cplx z = cplx(real_(tj) / X, imag_(tj) / X);
        cplx d = cplx(1.0, 0.0) - 2.0 * (real_(tj) / X) * z;
        double ad = abs_(d); detq = detq * cplx(d.x / ad, d.y / ad);
      }
    }
    __syncthreads();
    // ---- apply H_j^H to the trailing columns and refresh their partial norms
    {
      T tauc = conj_(tau[j]);
      if (abs2_(tauc) != 0.0) reflect_cols<T, MAXR>(A, ld, j, m, j + 1, n, v_s, tauc, PIVOT ? vn_s : nullptr);
      else if (PIVOT) {  // H = I: norms of rows below j only
        for (int c = j + 1 + warp; c < n; c += nwarps) {
          double s = 0.0;
          for (int i = j + 1 + lane; i < m; i += 32) s += abs2_(A[i + (long)c * ld]);
          s = warp_sum(s);
          if (lane == 0) vn_s[c] = sqrt(s);
        }
      }
    }
    __syncthreads();
  }
  // ---- D(i) = |R(i,i)|, R(i, i:) /= D(i); phases
  for (int i = tid; i < kmax; i += nthr) { double x = abs_(A[i + (long)i * ld]); D[i] = x; red_s[i] = x; }
  __syncthreads();
  for (long e = tid; e < (long)kmax * n; e += nthr) {
    int c = (int)((unsigned)e / (unsigned)kmax), i = (int)((unsigned)e - (unsigned)c * (unsigned)kmax);
    if (c >= i) A[i + (long)c * ld] = A[i + (long)c * ld] * (1.0 / red_s[i]);
  }
  for (int c = tid; c < n; c += nthr) jpvt[c] = ipv_s[c];
  __syncthreads();
  if (tid == 0) {
    cplx ph = cplx(1.0, 0.0);
    for (int i = 0; i < kmax; ++i) { T r = A[i + (long)i * ld]; ph = ph * cplx(real_(r), imag_(r)); }
    // permutation parity: cycles of even length flip the sign (QDRP_decompose_mod.F90:103-126)
    double sg = 1.0;
    for (int i = 0; i < n; ++i) vn_s[i] = 0.0;
    for (int i = 0; i < n; ++i) if (vn_s[i] == 0.0) {
      int next = i, L = 0;
      while (vn_s[next] == 0.0) { ++L; vn_s[next] = 1.0; next = ipv_s[next]; }
      if ((L & 1) == 0) sg = -sg;
    }
    out->perm_sign = sg; out->diag_phase = ph; out->detq = detq;
  }
  __syncthreads();
}

// dynamic smem layout: [optional matrix m*ld_s T] [v: m T] [vn: n double] [red: max(n,32) double] [ipv: n int]
template <typename T, int MAXR, int PIVOT, int STAGE>
__global__ void __launch_bounds__(512) k_qrp(T* __restrict__ A, int m, int n, int ld, long sA, T* __restrict__ tau, long sTau,
                                             int* __restrict__ jpvt, long sP, double* __restrict__ D, long sD, QrOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  A += (long)b * sA; tau += (long)b * sTau; jpvt += (long)b * sP; D += (long)b * sD;
  T* p = reinterpret_cast<T*>(smem_raw);
  T* As = nullptr; int lds = ld;
  if (STAGE) { As = p; lds = m; p += (long)m * n; }
  T* v_s = p; p += m;
  double* vn_s = reinterpret_cast<double*>(p);
  double* red_s = vn_s + n;
  int* ipv_s = reinterpret_cast<int*>(red_s + (n > 32 ? n : 32));
  if (STAGE) {
    for (long e = threadIdx.x; e < (long)m * n; e += blockDim.x) { int c = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)c * (unsigned)m); As[i + (long)c * m] = A[i + (long)c * ld]; }
    __syncthreads();
    qrp_device<T, MAXR, PIVOT>(As, m, n, lds, tau, jpvt, D, out + b, v_s, vn_s, ipv_s, red_s);
    for (long e = threadIdx.x; e < (long)m * n; e += blockDim.x) { int c = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)c * (unsigned)m); A[i + (long)c * ld] = As[i + (long)c * m]; }
  } else {
    qrp_device<T, MAXR, PIVOT>(A, m, n, ld, tau, jpvt, D, out + b, v_s, vn_s, ipv_s, red_s);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Small matrices (m, n <= 64; BASELINE config 2: N_dim = 64): the warp-per-column kernels above spend their time in shuffle reductions
// over two rows per lane.  Here ONE THREAD OWNS ONE COLUMN (staged in shared memory with an odd leading dimension, so the lanes of a
// warp hit different banks and the reflector is a broadcast): dot products and norms are serial per thread, no shuffles, 64 threads
// per matrix, several matrices per SM.  Same outputs as k_qrp.  Columns never move: the pivoting is a permutation table.  (The same layout was
// tried for the explicit Q and measured slower than k_formq, whose reflector applications have no pivot search in between.)
// ------------------------------------------------------------------------------------------------------------------------
template <typename T, int PIVOT>
__global__ void __launch_bounds__(64) k_qrp_small(T* __restrict__ A, int m, int n, int ld, long sA, T* __restrict__ tau, long sTau,
                                                  int* __restrict__ jpvt, long sP, double* __restrict__ D, long sD, QrOut* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, c = threadIdx.x, lane = c & 31;
  A += (long)b * sA; tau += (long)b * sTau; jpvt += (long)b * sP; D += (long)b * sD;
  const int lds = m | 1;
  T* As = reinterpret_cast<T*>(smem_raw);          // column c at As + c * lds
  T* v_s = As + (long)lds * 64;                    // reflector
  double* key_s = reinterpret_cast<double*>(v_s + 64);     // [2] per-warp best norm
  int* pos_s = reinterpret_cast<int*>(key_s + 2);  // [2] per-warp best position; then sto_at[64] (storage column at position), then scalars
  int* sto_at = pos_s + 2;
  __shared__ T s_tau; __shared__ int s_piv_sto;
  for (int e = c; e < m * n; e += 64) { const int i = e % m, cc = e / m; As[i + (long)cc * lds] = A[i + (long)cc * ld]; }
  if (c < n) sto_at[c] = c;
  __syncthreads();
  T* col = As + (long)c * lds;
  int mypos = c;                                   // position of my column in the permuted matrix
  bool done = c >= n;                              // columns already turned into reflectors
  double vn = 0.0;
  if (c < n) { for (int i = 0; i < m; ++i) vn += abs2_(col[i]); vn = sqrt(vn); }
  cplx detq = cplx(1.0, 0.0);
  const int kmax = (m < n) ? m : n;
  for (int j = 0; j < kmax; ++j) {
    // ---- pivot: largest partial norm among the remaining columns, lowest position on ties (ZGEQP3's IDAMAX over positions j..n-1)
    if (PIVOT) {
      double best = done ? -1.0 : vn; int bp = done ? (1 << 30) : mypos;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
      }
      if (lane == 0) { key_s[c >> 5] = best; pos_s[c >> 5] = bp; }
      __syncthreads();
      const double b0 = key_s[0], b1 = key_s[1]; const int p0 = pos_s[0], p1 = pos_s[1];
      const int pp = (b1 > b0 || (b1 == b0 && p1 < p0)) ? p1 : p0;       // position of the pivot column
      // swap positions j and pp
      const int sj = sto_at[j], sp = sto_at[pp];
      __syncthreads();
      if (c == sp) mypos = j; else if (c == sj) mypos = pp;
      if (c == 0) { sto_at[j] = sp; sto_at[pp] = sj; }
      if (c == sp) s_piv_sto = sp;
    } else if (c == 0) s_piv_sto = j;
    __syncthreads();
    const int ps = s_piv_sto;
    // ---- reflector from the pivot column (ZLARFG, no safmin rescaling loop): its owner works alone
    if (c == ps) {
      double xn2 = 0.0; for (int i = j + 1; i < m; ++i) xn2 += abs2_(col[i]);
      const T alpha = col[j]; T tj, scal; double beta;
      if (xn2 == 0.0 && imag_(alpha) == 0.0) { tj = zero_<T>(); scal = zero_<T>(); beta = real_(alpha); }
      else { beta = -copysign(sqrt(abs2_(alpha) + xn2), real_(alpha)); tj = make_<T>((beta - real_(alpha)) / beta, -imag_(alpha) / beta); scal = one_<T>() / (alpha - make_<T>(beta, 0.0)); }
      v_s[j] = one_<T>();
      for (int i = j + 1; i < m; ++i) { const T y = col[i] * scal; col[i] = y; v_s[i] = y; }
      col[j] = make_<T>(beta, 0.0); s_tau = tj; tau[j] = tj;
      done = true;
    }
    __syncthreads();
    const T tj = s_tau;
    if (abs2_(tj) != 0.0) {      // det(H_j) = 1 - 2 (tau/|tau|) (Re tau/|tau|)   (Prog/cgr1_mod.F90:338-347); every thread keeps the same product
      const double X = abs_(tj); const cplx z = cplx(real_(tj) / X, imag_(tj) / X);
      const cplx d = cplx(1.0, 0.0) - 2.0 * (real_(tj) / X) * z; const double ad = abs_(d); detq = detq * cplx(d.x / ad, d.y / ad);
    }
    // ---- apply H_j^H to my column and refresh its partial norm
    if (!done) {
      if (abs2_(tj) != 0.0) {
        const T tauc = conj_(tj);
        T w = zero_<T>(); for (int i = j; i < m; ++i) fmac_(w, v_s[i], col[i]);
        const T sc = tauc * w; double nrm = 0.0;
        for (int i = j; i < m; ++i) { const T y = col[i] - v_s[i] * sc; col[i] = y; if (i > j) nrm += abs2_(y); }
        vn = sqrt(nrm);
      } else if (PIVOT) { double nrm = 0.0; for (int i = j + 1; i < m; ++i) nrm += abs2_(col[i]); vn = sqrt(nrm); }
    }
    __syncthreads();
  }
  // ---- outputs: columns in position order, D(i) = |R(i,i)|, R(i, i:) /= D(i), jpvt, phases
  if (c < kmax) { const double x = abs_(As[c + (long)sto_at[c] * lds]); D[c] = x; key_s[0] = 0.0; reinterpret_cast<double*>(v_s)[c] = x; }
  __syncthreads();
  const double* dd = reinterpret_cast<const double*>(v_s);
  if (c < n) {
    T* dst = A + (long)mypos * ld;
    for (int i = 0; i < m; ++i) { T y = col[i]; if (i <= mypos && i < kmax) y = y * (1.0 / dd[i]); dst[i] = y; }
    jpvt[mypos] = c;
  }
  __syncthreads();
  if (c == 0) {
    cplx ph = cplx(1.0, 0.0);
    for (int i = 0; i < kmax; ++i) { const T r = As[i + (long)sto_at[i] * lds] * (1.0 / dd[i]); ph = ph * cplx(real_(r), imag_(r)); }
    double sg = 1.0; unsigned long long seen = 0ull;        // permutation parity: cycles of even length flip the sign (n <= 64)
    for (int i = 0; i < n; ++i) if (!((seen >> i) & 1ull)) { int next = i, L = 0; while (!((seen >> next) & 1ull)) { ++L; seen |= 1ull << next; next = sto_at[next]; } if ((L & 1) == 0) sg = -sg; }
    out[b].perm_sign = sg; out[b].diag_phase = ph; out[b].detq = detq;
  }
}
static size_t qrp_small_smem(size_t elem, int m) { return elem * ((size_t)(m | 1) * 64 + 64) + 2 * sizeof(double) + (2 + 64) * sizeof(int) + 64; }

// Explicit Q (m x n, n reflectors) in place of the reflectors, backward accumulation (ZUNG2R order).
// Optionally scales column 0 by `colscale[b]` afterwards (udv_state_mod.F90:578).
template <typename T, int MAXR, int STAGE>
__global__ void __launch_bounds__(512) k_formq(T* __restrict__ A, int m, int n, int ld, long sA, const T* __restrict__ tau, long sTau,
                                               const cplx* __restrict__ colscale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  A += (long)b * sA; tau += (long)b * sTau;
  T* p = reinterpret_cast<T*>(smem_raw);
  T* W = A; int lw = ld;
  if (STAGE) {
    W = p; lw = m; p += (long)m * n;
    for (long e = tid; e < (long)m * n; e += nthr) { int c = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)c * (unsigned)m); W[i + (long)c * m] = A[i + (long)c * ld]; }
  }
  T* v_s = p;
  __syncthreads();
  for (int j = n - 1; j >= 0; --j) {
    const int len = m - j;
    T* col = W + (long)j * lw + j;
    for (int i = tid; i < len; i += nthr) v_s[i] = (i == 0) ? one_<T>() : col[i];
    __syncthreads();
    T tj = tau[j];
    if (j < n - 1 && abs2_(tj) != 0.0) reflect_cols<T, MAXR>(W, lw, j, m, j + 1, n, v_s, tj, nullptr);
    // column j of Q: e_j - tau_j v
    for (int i = tid; i < m; i += nthr) {
      T y;
      if (i < j) y = zero_<T>();
      else if (i == j) y = one_<T>() - tj;
      else y = -(tj * v_s[i - j]);
      W[i + (long)j * lw] = y;
    }
    __syncthreads();
  }
  if (colscale) {
    cplx s = colscale[b]; T st = make_<T>(s.x, s.y);
    for (int i = tid; i < m; i += nthr) W[i] = W[i] * st;
    __syncthreads();
  }
  if (STAGE) {
    for (long e = tid; e < (long)m * n; e += nthr) { int c = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)c * (unsigned)m); A[i + (long)c * ld] = W[i + (long)c * m]; }
  }
}

// X = R^-1 B, R upper triangular n x n (as stored in a QR'd matrix, leading dimension ldr), B n x nrhs in place.
// A CTA owns TRSM_COLS right-hand sides; each warp keeps its columns in registers (rows over lanes);
// column i of R is staged through shared memory once per elimination step.
// If dinv != nullptr, B(i,:) is first multiplied by 1/dinv[i]  (the "apply inverse of D" loop, cgr1_mod.F90:368-377).
#define TRSM_WARPS 8
#define TRSM_CPW 4
#define TRSM_COLS (TRSM_WARPS * TRSM_CPW)
template <typename T, int MAXR, int LOWER = 0>
__global__ void __launch_bounds__(TRSM_WARPS * 32) k_trsm_lun(const T* __restrict__ R, int ldr, long sR, T* __restrict__ B, int ldb, long sB,
                                                             int n, int nrhs, const double* __restrict__ dinv, long sD) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* rcol = reinterpret_cast<T*>(smem_raw);        // n entries
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  R += (long)b * sR; B += (long)b * sB;
  if (dinv) dinv += (long)b * sD;
  const int c0 = blockIdx.x * TRSM_COLS + warp * TRSM_CPW;
  T x[TRSM_CPW][MAXR];
#pragma unroll
  for (int cc = 0; cc < TRSM_CPW; ++cc)
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
      int i = lane + 32 * q, c = c0 + cc;
      T v = zero_<T>();
      if (i < n && c < nrhs) { v = B[i + (long)c * ldb]; if (dinv) v = v * (1.0 / dinv[i]); }
      x[cc][q] = v;
    }
  // LOWER = 0: back substitution with the upper triangle; LOWER = 1: forward substitution with the lower triangle (R holds L)
  for (int ii = 0; ii < n; ++ii) {
    const int i = LOWER ? ii : n - 1 - ii;
    __syncthreads();
    if (LOWER) { for (int k = i + tid; k < n; k += blockDim.x) rcol[k] = R[k + (long)i * ldr]; }
    else { for (int k = tid; k <= i; k += blockDim.x) rcol[k] = R[k + (long)i * ldr]; }
    __syncthreads();
    const T rinv = one_<T>() / rcol[i];
    const int qi = i >> 5, li = i & 31;
#pragma unroll
    for (int cc = 0; cc < TRSM_CPW; ++cc) {
      T xi = zero_<T>();
#pragma unroll
      for (int q = 0; q < MAXR; ++q) if (q == qi) xi = x[cc][q];
      xi = shfl_(xi, li) * rinv;
#pragma unroll
      for (int q = 0; q < MAXR; ++q) {
        int k = lane + 32 * q;
        if (LOWER ? (k > i && k < n) : (k < i)) x[cc][q] = x[cc][q] - rcol[k] * xi;
        else if (k == i) x[cc][q] = xi;
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < TRSM_CPW; ++cc)
#pragma unroll
    for (int q = 0; q < MAXR; ++q) {
      int i = lane + 32 * q, c = c0 + cc;
      if (i < n && c < nrhs) B[i + (long)c * ldb] = x[cc][q];
    }
}

// ------------------------------------------------------------------------------------------------
// Blocked triangular solve on the FP64 tensor cores (real matrices): X = R^-1 diag(dinv)^-1 B (LOWER = 0, back substitution,
// strict lower part of the stored matrix ignored) or X = L^-1 B (LOWER = 1, forward substitution, strict upper part ignored).
// Step 1 (k_tri_inv_blocks): the 32 x 32 diagonal blocks are inverted, one warp per block (lane j solves for column j).
// Step 2 (k_trsm_blk): a CTA owns TRSMB_CW right-hand sides of one matrix in shared memory and sweeps the block rows:
//   X_k = inv(R_kk) B_k ;  B_i -= R_ik X_k for the remaining block rows i     -- both as DMMA m8n8k4 products whose A operand
//   (R, inv(R_kk)) is read from global memory / L2 directly in fragment layout and whose B operand is the X panel.
// n is zero-padded to a multiple of 32 (identity on the padded diagonal).
// ------------------------------------------------------------------------------------------------
#define TRSMB_BS 32
#define TRSMB_CW 32
template <int LOWER>
__global__ void __launch_bounds__(32) k_tri_inv_blocks(const double* __restrict__ R, int ldr, long sR, int n, double* __restrict__ Rinv, long sI) {
  __shared__ double Rs[32][33];
  __shared__ double Xs[32][33];      // Xs[k][j]: entry k of column j of the inverse
  const int kb = blockIdx.x, b = blockIdx.y, lane = threadIdx.x, k0 = kb * 32;
  R += (long)b * sR; Rinv += (long)b * sI + (long)kb * 1024;
  for (int c = 0; c < 32; ++c) {
    const int i = k0 + lane, j = k0 + c;
    double v = (lane == c) ? 1.0 : 0.0;
    if (i < n && j < n && (LOWER ? (i >= j) : (i <= j))) v = R[i + (long)j * ldr];
    Rs[lane][c] = v;
  }
  for (int k = 0; k < 32; ++k) Xs[k][lane] = 0.0;
  __syncwarp();
  const int j = lane;
  if (!LOWER) {
    for (int i = 31; i >= 0; --i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = i + 1; k < 32; ++k) s = fma(-Rs[i][k], Xs[k][j], s);
      Xs[i][j] = (i <= j) ? s / Rs[i][i] : 0.0;
    }
  } else {
    for (int i = 0; i < 32; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s = fma(-Rs[i][k], Xs[k][j], s);
      Xs[i][j] = (i >= j) ? s / Rs[i][i] : 0.0;
    }
  }
  __syncwarp();
  for (int c = 0; c < 32; ++c) Rinv[lane + c * 32] = Xs[lane][c];      // column-major 32 x 32
}

// zero_cols_from / zero_rows (LOWER only): right-hand sides c >= zero_cols_from are known to vanish in rows < zero_rows (block-diagonal
// right-hand side of solve_extended_System), so their forward substitution starts at block row zero_rows / 32.
// dout != nullptr: X(i,:) is divided by dout[i] on the way out (the "apply inverse of D" loop of cgr2_2_mod.F90:183-188).
template <int LOWER>
__global__ void __launch_bounds__(512) k_trsm_blk(const double* __restrict__ R, int ldr, long sR, const double* __restrict__ Rinv, long sI,
                                                  double* __restrict__ B, int ldb, long sB, int n, int nrhs, const double* __restrict__ dinv, long sD,
                                                  int zero_cols_from, int zero_rows, const double* __restrict__ dout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Xs = reinterpret_cast<double*>(smem_raw);          // [TRSMB_CW][ldx], rows contiguous
  const int b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  const int g = lane >> 2, q = lane & 3;
  R += (long)b * sR; Rinv += (long)b * sI; B += (long)b * sB;
  if (dinv) dinv += (long)b * sD;
  if (dout) dout += (long)b * sD;
  const int np = (n + 31) & ~31, nb = np >> 5, ldx = ld_pad(np);
  const int c0 = blockIdx.x * TRSMB_CW;
  const int s_begin = (LOWER && zero_rows > 0 && c0 >= zero_cols_from) ? (zero_rows >> 5) : 0;
  for (int e = tid; e < np * TRSMB_CW; e += nthr) {
    const int i = e % np, c = e / np;
    double v = 0.0;
    if (i < n && c0 + c < nrhs) { v = B[i + (long)(c0 + c) * ldb]; if (dinv) v = v * (1.0 / dinv[i]); }
    Xs[i + c * ldx] = v;
  }
  __syncthreads();
  for (int s = s_begin; s < nb; ++s) {
    const int kb = LOWER ? s : nb - 1 - s, k0 = kb * 32;
    // B fragments of the current block row of X: bf[ks][cb] = X(k0 + 4 ks + q, 8 cb + g)
    double bf[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) bf[ks][cb] = Xs[(k0 + 4 * ks + q) + (8 * cb + g) * ldx];
    // ---- X_k = inv(R_kk) B_k : 4 row tiles of 8 rows, warps 0..3
    double acc[4][2];
    if (warp < 4) {
      double af[8];
      const double* ri = Rinv + (long)kb * 1024 + (8 * warp + g);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) af[ks] = ri[(4 * ks + q) * 32];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) { acc[cb][0] = 0.0; acc[cb][1] = 0.0; }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma884(acc[cb][0], acc[cb][1], af[ks], bf[ks][cb]);
    }
    __syncthreads();                     // every warp holds the old block row in bf
    if (warp < 4) {
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) { Xs[(k0 + 8 * warp + g) + (8 * cb + 2 * q) * ldx] = acc[cb][0]; Xs[(k0 + 8 * warp + g) + (8 * cb + 2 * q + 1) * ldx] = acc[cb][1]; }
    }
    __syncthreads();
    if (s == nb - 1) break;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) bf[ks][cb] = Xs[(k0 + 4 * ks + q) + (8 * cb + g) * ldx];
    // ---- B_i -= R_ik X_k for the block rows not yet solved: 8-row tiles over the warps
    const int r_lo = LOWER ? (k0 + 32) : 0, r_hi = LOWER ? np : k0;
    for (int i0 = r_lo + 8 * warp; i0 < r_hi; i0 += 8 * nw) {
      double af[8];
      const int i = i0 + g;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) { const int k = k0 + 4 * ks + q; af[ks] = (i < n && k < n) ? -R[i + (long)k * ldr] : 0.0; }
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) { acc[cb][0] = Xs[i + (8 * cb + 2 * q) * ldx]; acc[cb][1] = Xs[i + (8 * cb + 2 * q + 1) * ldx]; }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) dmma884(acc[cb][0], acc[cb][1], af[ks], bf[ks][cb]);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) { Xs[i + (8 * cb + 2 * q) * ldx] = acc[cb][0]; Xs[i + (8 * cb + 2 * q + 1) * ldx] = acc[cb][1]; }
    }
    __syncthreads();
  }
  for (int e = tid; e < n * TRSMB_CW; e += nthr) {
    const int i = e % n, c = e / n;
    if (c0 + c < nrhs) B[i + (long)(c0 + c) * ldb] = dout ? Xs[i + c * ldx] * (1.0 / dout[i]) : Xs[i + c * ldx];
  }
}

// ---- complex variant of the blocked solve: 16 right-hand sides per CTA, complex products as four real DMMA products
#define TRSMB_CWC 16
template <int LOWER>
__global__ void __launch_bounds__(32) k_tri_inv_blocks_c(const cplx* __restrict__ R, int ldr, long sR, int n, cplx* __restrict__ Rinv, long sI) {
  __shared__ cplx Rs[32][33];
  __shared__ cplx Xs[32][33];
  const int kb = blockIdx.x, b = blockIdx.y, lane = threadIdx.x, k0 = kb * 32;
  R += (long)b * sR; Rinv += (long)b * sI + (long)kb * 1024;
  for (int c = 0; c < 32; ++c) {
    const int i = k0 + lane, j = k0 + c;
    cplx v = (lane == c) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
    if (i < n && j < n && (LOWER ? (i >= j) : (i <= j))) v = R[i + (long)j * ldr];
    Rs[lane][c] = v;
  }
  for (int k = 0; k < 32; ++k) Xs[k][lane] = cplx(0.0, 0.0);
  __syncwarp();
  const int j = lane;
  if (!LOWER) {
    for (int i = 31; i >= 0; --i) {
      cplx s = (i == j) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
      for (int k = i + 1; k < 32; ++k) s = s - Rs[i][k] * Xs[k][j];
      Xs[i][j] = (i <= j) ? s / Rs[i][i] : cplx(0.0, 0.0);
    }
  } else {
    for (int i = 0; i < 32; ++i) {
      cplx s = (i == j) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
      for (int k = 0; k < i; ++k) s = s - Rs[i][k] * Xs[k][j];
      Xs[i][j] = (i >= j) ? s / Rs[i][i] : cplx(0.0, 0.0);
    }
  }
  __syncwarp();
  for (int c = 0; c < 32; ++c) Rinv[lane + c * 32] = Xs[lane][c];
}

__device__ __forceinline__ void cmma(double (&re)[2], double (&im)[2], cplx a, cplx b) {
  dmma884(re[0], re[1], a.x, b.x); dmma884(re[0], re[1], -a.y, b.y);
  dmma884(im[0], im[1], a.x, b.y); dmma884(im[0], im[1], a.y, b.x);
}

template <int LOWER>
__global__ void __launch_bounds__(512) k_trsm_blk_c(const cplx* __restrict__ R, int ldr, long sR, const cplx* __restrict__ Rinv, long sI,
                                                    cplx* __restrict__ B, int ldb, long sB, int n, int nrhs, const double* __restrict__ dinv, long sD) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Xs = reinterpret_cast<cplx*>(smem_raw);          // [TRSMB_CWC][ldx]
  const int b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  const int g = lane >> 2, q = lane & 3;
  R += (long)b * sR; Rinv += (long)b * sI; B += (long)b * sB;
  if (dinv) dinv += (long)b * sD;
  const int np = (n + 31) & ~31, nb = np >> 5, ldx = np + 1;
  const int c0 = blockIdx.x * TRSMB_CWC;
  for (int e = tid; e < np * TRSMB_CWC; e += nthr) {
    const int i = e % np, c = e / np;
    cplx v = cplx(0.0, 0.0);
    if (i < n && c0 + c < nrhs) { v = B[i + (long)(c0 + c) * ldb]; if (dinv) v = v * (1.0 / dinv[i]); }
    Xs[i + c * ldx] = v;
  }
  __syncthreads();
  for (int s = 0; s < nb; ++s) {
    const int kb = LOWER ? s : nb - 1 - s, k0 = kb * 32;
    cplx bf[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) bf[ks][cb] = Xs[(k0 + 4 * ks + q) + (8 * cb + g) * ldx];
    double are[2][2], aim[2][2];
    if (warp < 4) {
      cplx af[8];
      const cplx* ri = Rinv + (long)kb * 1024 + (8 * warp + g);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) af[ks] = ri[(4 * ks + q) * 32];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) { are[cb][0] = are[cb][1] = 0.0; aim[cb][0] = aim[cb][1] = 0.0; }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) cmma(are[cb], aim[cb], af[ks], bf[ks][cb]);
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        Xs[(k0 + 8 * warp + g) + (8 * cb + 2 * q) * ldx] = cplx(are[cb][0], aim[cb][0]);
        Xs[(k0 + 8 * warp + g) + (8 * cb + 2 * q + 1) * ldx] = cplx(are[cb][1], aim[cb][1]);
      }
    }
    __syncthreads();
    if (s == nb - 1) break;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) bf[ks][cb] = Xs[(k0 + 4 * ks + q) + (8 * cb + g) * ldx];
    const int r_lo = LOWER ? (k0 + 32) : 0, r_hi = LOWER ? np : k0;
    for (int i0 = r_lo + 8 * warp; i0 < r_hi; i0 += 8 * nw) {
      cplx af[8];
      const int i = i0 + g;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) { const int k = k0 + 4 * ks + q; af[ks] = (i < n && k < n) ? -R[i + (long)k * ldr] : cplx(0.0, 0.0); }
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const cplx x0 = Xs[i + (8 * cb + 2 * q) * ldx], x1 = Xs[i + (8 * cb + 2 * q + 1) * ldx];
        are[cb][0] = x0.x; aim[cb][0] = x0.y; are[cb][1] = x1.x; aim[cb][1] = x1.y;
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) cmma(are[cb], aim[cb], af[ks], bf[ks][cb]);
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        Xs[i + (8 * cb + 2 * q) * ldx] = cplx(are[cb][0], aim[cb][0]); Xs[i + (8 * cb + 2 * q + 1) * ldx] = cplx(are[cb][1], aim[cb][1]);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < n * TRSMB_CWC; e += nthr) {
    const int i = e % n, c = e / n;
    if (c0 + c < nrhs) B[i + (long)(c0 + c) * ldb] = Xs[i + c * ldx];
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise helpers (all batched over blockIdx.y)
// ------------------------------------------------------------------------------------------------
// dst(i,j) = src(perm? ...) generic gather: MODE 0 copy, 1 rows gathered dst(i,:) = src(p[i],:), 2 cols gathered dst(:,j) = src(:,p[j]),
// 3 rows scattered dst(p[i],:) = src(i,:), 4 conj-transpose dst(i,j) = conj(src(j,i)),
// 5 rows scattered + conj-transpose: dst = (scatter_rows(src))^H  i.e. dst(j, p[i]) = conj(src(i,j))
template <typename T, int MODE>
__global__ void k_permcopy(T* __restrict__ dst, int ldd, long sDst, const T* __restrict__ src, int lds, long sSrc, int m, int n,
                           const int* __restrict__ perm, long sP) {
  const int b = blockIdx.y;
  dst += (long)b * sDst; src += (long)b * sSrc;
  if (perm) perm += (long)b * sP;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)m * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)j * (unsigned)m);
    T v = src[i + (long)j * lds];
    if (MODE == 0) dst[i + (long)j * ldd] = v;
    else if (MODE == 1) dst[i + (long)j * ldd] = src[perm[i] + (long)j * lds];
    else if (MODE == 2) dst[i + (long)j * ldd] = src[i + (long)perm[j] * lds];
    else if (MODE == 3) dst[perm[i] + (long)j * ldd] = v;
    else if (MODE == 4) dst[j + (long)i * ldd] = conj_(v);
    else if (MODE == 5) dst[j + (long)perm[i] * ldd] = conj_(v);
  }
}

// dst = src^H for n x n matrices through 32 x 32 shared-memory tiles (both sides coalesced); grid = (tiles^2, batch), 256 threads
template <typename T>
__global__ void __launch_bounds__(256) k_transpose_conj(T* __restrict__ dst, int ldd, long sDst, const T* __restrict__ src, int lds, long sSrc, int n) {
  __shared__ T tile[32][33];
  const int b = blockIdx.y, nt = (n + 31) / 32, ti = blockIdx.x % nt, tj = blockIdx.x / nt;
  dst += (long)b * sDst; src += (long)b * sSrc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int c = ty; c < 32; c += 8) { const int i = ti * 32 + tx, j = tj * 32 + c; if (i < n && j < n) tile[c][tx] = src[i + (long)j * lds]; }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) { const int i = tj * 32 + tx, j = ti * 32 + c; if (i < n && j < n) dst[i + (long)j * ldd] = conj_(tile[tx][c]); }
}

// A(:, j) *= d[j]   (udv_state_mod.F90:473-477)
template <typename T>
__global__ void k_colscale(T* __restrict__ A, int ld, long sA, int m, int n, const double* __restrict__ d, long sD) {
  const int b = blockIdx.y; A += (long)b * sA; d += (long)b * sD;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)m * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)j * (unsigned)m);
    A[i + (long)j * ld] = A[i + (long)j * ld] * d[j];
  }
}
// row 0 of A (first n entries) *= conj?(1/phase[b])   (udv_state_mod.F90:489-492)
template <typename T>
__global__ void k_row0scale(T* __restrict__ A, int ld, long sA, int n, const cplx* __restrict__ s) {
  const int b = blockIdx.y; A += (long)b * sA;
  T st = make_<T>(s[b].x, s[b].y);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) A[(long)j * ld] = A[(long)j * ld] * st;
}
// dst = src with column 0 multiplied by s[b]  (explicit Q of the blocked path -> U, udv_state_mod.F90:578)
template <typename T>
__global__ void k_copy_col0scale(T* __restrict__ dst, const T* __restrict__ src, long sM, int n, const cplx* __restrict__ s, long count) {
  const int b = blockIdx.y; dst += (long)b * sM; src += (long)b * sM;
  const T st = make_<T>(s[b].x, s[b].y);
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) dst[e] = (e < n) ? src[e] * st : src[e];
}
// identity / zero fill
template <typename T>
__global__ void k_set_identity(T* __restrict__ A, int ld, long sA, int m, int n) {
  const int b = blockIdx.y; A += (long)b * sA;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)m * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)j * (unsigned)m);
    A[i + (long)j * ld] = (i == j) ? one_<T>() : zero_<T>();
  }
}
static __global__ void k_fill_double(double* p, long n, double v) {
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) p[e] = v;
}
static __global__ void k_fill_cplx(cplx* p, long n, cplx v) {
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) p[e] = v;
}

// TPUP(i,j) = DR(i) * TPUP(i,j) * DL(j) + RHS(i,j)      (cgr1_mod.F90:233-236)
// SEP = 1: scale-separated form of the STAB3 branch (cgr1_mod.F90:241-268)
// CT = 1: the result is written conjugate-transposed (NVAR /= 1, cgr1_mod.F90:303-308)
template <typename T, int SEP, int CT>
__global__ void k_cgr_tpup(T* __restrict__ OUT, const T* __restrict__ TP, const T* __restrict__ RHS, long sM, int n,
                           const double* __restrict__ DR, const double* __restrict__ DL, long sD) {
  const int b = blockIdx.y; OUT += (long)b * sM; TP += (long)b * sM; RHS += (long)b * sM; DR += (long)b * sD; DL += (long)b * sD;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)n), i = (int)((unsigned)e - (unsigned)j * (unsigned)n);
    double dr = DR[i], dl = DL[j];
    T t = TP[e], r = RHS[e], o;
    if (!SEP) o = (dr * t) * dl + r;
    else {
      if (dl <= 1.0) { if (dr <= 1.0) o = r + (dr * dl) * t; else o = (1.0 / dr) * r + dl * t; }
      else { if (dr <= 1.0) o = (1.0 / dl) * r + dr * t; else o = (r * (1.0 / dr)) * (1.0 / dl) + t; }
    }
    if (CT) OUT[j + (long)i * n] = conj_(o); else OUT[e] = o;
  }
}
// rows i with d[i] > 1 scaled by 1/d[i] (ROWS=1) or columns (ROWS=0): the D_+^-1 scalings of the STAB3 branch
template <typename T, int ROWS>
__global__ void k_sep_scale(T* __restrict__ A, long sA, int n, const double* __restrict__ d, long sD) {
  const int b = blockIdx.y; A += (long)b * sA; d += (long)b * sD;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)n), i = (int)((unsigned)e - (unsigned)j * (unsigned)n);
    double x = d[ROWS ? i : j];
    if (x > 1.0) A[e] = A[e] * (1.0 / x);
  }
}

// A(i, :) *= 1/d[i]
template <typename T>
__global__ void k_rowscale_inv(T* __restrict__ A, int ld, long sA, int m, int n, const double* __restrict__ d, long sD) {
  const int b = blockIdx.y; A += (long)b * sA; d += (long)b * sD;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)m * n; e += (long)gridDim.x * blockDim.x) {
    int j = (int)((unsigned)e / (unsigned)m), i = (int)((unsigned)e - (unsigned)j * (unsigned)m);
    A[i + (long)j * ld] = A[i + (long)j * ld] * (1.0 / d[i]);
  }
}

// CGR2_2 (Prog/cgr2_2_mod.F90:318-425): builds HLPB1 = HLPB2^H (input of the 2N x 2N pivoted QR) and the block-diagonal
// right-hand side HLP = diag(UCT, VINV) of solve_extended_System (:155-191).  Per matrix the ordering of the blocks depends
// on D1(1) > D2(1) (:371, :398); the choice is recorded in first[b] for get_blocks.  STAB3: D_+ scalings (:352-366).
template <typename T, int STAB3>
__global__ void k_cgr22_build(T* __restrict__ HLPB1, T* __restrict__ HLP, long s22, const T* __restrict__ V1INV, const T* __restrict__ U1,
                              const double* __restrict__ D1, const T* __restrict__ U2, const T* __restrict__ V2, const double* __restrict__ D2,
                              long sM, long sD, int N, int* __restrict__ first) {
  const int b = blockIdx.y; const int N2 = 2 * N;
  HLPB1 += (long)b * s22; HLP += (long)b * s22; V1INV += (long)b * sM; U1 += (long)b * sM; U2 += (long)b * sM; V2 += (long)b * sM; D1 += (long)b * sD; D2 += (long)b * sD;
  const bool fst = D1[0] > D2[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) first[b] = fst ? 1 : 0;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)N2 * N2; e += (long)gridDim.x * blockDim.x) {
    const int J2 = (int)((unsigned)e / (unsigned)N2), I2 = (int)((unsigned)e - (unsigned)J2 * (unsigned)N2);
    const int bi = I2 >= N, bj = J2 >= N, I = I2 - bi * N, J = J2 - bj * N;
    // blocks of HLPB2 in the "first" ordering: (0,0) V1INV, (0,1) D1 U1^H, (1,0) -D2 V2, (1,1) U2^H ; otherwise both block rows and columns swap roles
    const int kind = fst ? (bi * 2 + bj) : ((1 - bi) * 2 + (1 - bj));
    T v, hl = zero_<T>();
    const double d1 = D1[I], d2 = D2[I];
    const double s1 = (STAB3 && d1 > 1.0) ? 1.0 / d1 : 1.0, m1 = (STAB3 && d1 > 1.0) ? 1.0 : d1;
    const double s2 = (STAB3 && d2 > 1.0) ? 1.0 / d2 : 1.0, m2 = (STAB3 && d2 > 1.0) ? 1.0 : d2;
    if (kind == 0) { v = V1INV[I + (long)J * N] * s1; hl = v; }
    else if (kind == 1) v = m1 * conj_(U1[J + (long)I * N]);
    else if (kind == 2) v = -(m2 * V2[I + (long)J * N]);
    else { v = conj_(U2[J + (long)I * N]) * s2; hl = v; }
    HLPB1[J2 + (long)I2 * N2] = conj_(v);
    HLP[I2 + (long)J2 * N2] = hl;
  }
}
// The same for N a multiple of 32: one CTA per 32 x 32 tile (inside one quadrant), every global access coalesced -- the transposed operands
// (U1^H, U2^H) and the transposed output HLPB1 go through a shared-memory tile.  grid = ((2N/32)^2, batch), 256 threads.
template <typename T, int STAB3>
__global__ void __launch_bounds__(256) k_cgr22_build_tiled(T* __restrict__ HLPB1, T* __restrict__ HLP, long s22, const T* __restrict__ V1INV, const T* __restrict__ U1,
                                                           const double* __restrict__ D1, const T* __restrict__ U2, const T* __restrict__ V2, const double* __restrict__ D2,
                                                           long sM, long sD, int N, int* __restrict__ first) {
  __shared__ T tile[32][33];       // tile[j][i] = v(I2 = i0 + i, J2 = j0 + j)
  const int b = blockIdx.y, N2 = 2 * N, nt2 = N2 >> 5, ti = blockIdx.x % nt2, tj = blockIdx.x / nt2, i0 = ti * 32, j0 = tj * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  HLPB1 += (long)b * s22; HLP += (long)b * s22; V1INV += (long)b * sM; U1 += (long)b * sM; U2 += (long)b * sM; V2 += (long)b * sM; D1 += (long)b * sD; D2 += (long)b * sD;
  const bool fst = D1[0] > D2[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) first[b] = fst ? 1 : 0;
  const int bi = i0 >= N, bj = j0 >= N, I0 = i0 - bi * N, J0 = j0 - bj * N;
  const int kind = fst ? (bi * 2 + bj) : ((1 - bi) * 2 + (1 - bj));
  const bool keep = (kind == 0 || kind == 3);                  // the blocks that also go into HLP
  if (kind == 0 || kind == 2) {                                // direct operands: lanes along I
    for (int r = ty; r < 32; r += 8) {
      const int I = I0 + tx, J = J0 + r;
      const double d1 = D1[I], d2 = D2[I];
      T v;
      if (kind == 0) v = V1INV[I + (long)J * N] * ((STAB3 && d1 > 1.0) ? 1.0 / d1 : 1.0);
      else v = -(((STAB3 && d2 > 1.0) ? 1.0 : d2) * V2[I + (long)J * N]);
      tile[r][tx] = v;
    }
  } else {                                                     // transposed operands U^H(I, J) = conj(U(J, I)): lanes along J
    for (int r = ty; r < 32; r += 8) {
      const int I = I0 + r, J = J0 + tx;
      const double d1 = D1[I], d2 = D2[I];
      T v;
      if (kind == 1) v = ((STAB3 && d1 > 1.0) ? 1.0 : d1) * conj_(U1[J + (long)I * N]);
      else v = conj_(U2[J + (long)I * N]) * ((STAB3 && d2 > 1.0) ? 1.0 / d2 : 1.0);
      tile[tx][r] = v;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    HLP[(i0 + tx) + (long)(j0 + r) * N2] = keep ? tile[r][tx] : zero_<T>();          // lanes along I2
    HLPB1[(j0 + tx) + (long)(i0 + r) * N2] = conj_(tile[tx][r]);                      // lanes along J2
  }
}
// get_blocks (cgr2_2_mod.F90:55-72) with the ordering flag: first: (A,B,C,D) = (G00,G0T,GT0,GTT), else (GTT,GT0,G0T,G00)
template <typename T>
__global__ void k_cgr22_blocks(const T* __restrict__ INP, long s22, T* __restrict__ GT0, T* __restrict__ G00, T* __restrict__ GTT, T* __restrict__ G0T,
                               long sM, int N, const int* __restrict__ first) {
  const int b = blockIdx.y; const int N2 = 2 * N;
  INP += (long)b * s22; GT0 += (long)b * sM; G00 += (long)b * sM; GTT += (long)b * sM; G0T += (long)b * sM;
  const bool fst = first[b] != 0;
#pragma unroll 4
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)N * N; e += (long)gridDim.x * blockDim.x) {
    const int J = (int)((unsigned)e / (unsigned)N), I = (int)((unsigned)e - (unsigned)J * (unsigned)N);
    const T a = INP[I + (long)J * N2], d = INP[(I + N) + (long)(J + N) * N2], c = INP[(I + N) + (long)J * N2], bb = INP[I + (long)(J + N) * N2];
    if (fst) { G00[e] = a; G0T[e] = bb; GT0[e] = c; GTT[e] = d; }
    else { GTT[e] = a; GT0[e] = bb; G0T[e] = c; G00[e] = d; }
  }
}

// Control_PrecisionG / COMPARE (control_mod.F90:207-298, mymats_mod.F90:683-704): per matrix max and mean |A-B|, NaN flag.
// The matrix is split over CMP_SPLIT CTAs (grid = (matrices, CMP_SPLIT)); slot (b, part): out[(b*CMP_SPLIT+part)*3 + 0] = xmax of the
// part, +1 = its sum of |A-B| divided by nelem (the parts add up to xmean), +2 = nan flag.  k_ctl_accum folds the parts.
#define CMP_SPLIT 8
template <typename T>
__global__ void __launch_bounds__(256) k_compare(const T* __restrict__ A, const T* __restrict__ B, long sM, long nelem, double* __restrict__ out) {
  __shared__ double smax[8], ssum[8]; __shared__ int snan[8];
  const int b = blockIdx.x; A += (long)b * sM; B += (long)b * sM;
  double mx = 0.0, sm = 0.0; int nn = 0;
  for (long e = (long)blockIdx.y * blockDim.x + threadIdx.x; e < nelem; e += (long)blockDim.x * gridDim.y) {
    T a = A[e], c = B[e];
    if (isnan_(a) || isnan_(c)) nn = 1;
    double d = abs_(a - c);
    mx = fmax(mx, d); sm += d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); sm += __shfl_xor_sync(0xffffffffu, sm, o); nn |= __shfl_xor_sync(0xffffffffu, nn, o); }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sm; snan[threadIdx.x >> 5] = nn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mx = fmax(mx, smax[w]); sm += ssum[w]; nn |= snan[w]; }
    double* o = out + ((long)b * gridDim.y + blockIdx.y) * 3;
    o[0] = mx; o[1] = sm / (double)nelem; o[2] = (double)nn;
  }
}
