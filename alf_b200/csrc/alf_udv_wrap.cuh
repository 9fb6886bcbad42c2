// UDV_Wrap_Pivot (Prog/UDV_WRAP_mod.F90:125-208, the default non-STAB1 variant) on the device: the stabilisation of the legacy
// STAB1 / STAB2 builds (wrapur_mod.F90:96, wrapul_mod.F90:98-102, cgr1_mod.F90:115,137).  A (N1 x N2, N2 <= N1) -> U D V with
//   * columns sorted by decreasing squared norm XNORM (first index wins ties, :141-153) and divided by XNORM (:154-157),
//   * UDV_C = unpivoted Householder QR (Libraries/Modules/mymats_mod.F90:933-1039): sign of det R moved into U(:,1) / V(1,:),
//     D = |R_ii|, V unit upper triangular,
//   * det V = 1 through Pivot_Phase (:159-168), scaling and permutation undone in D and V (:170-185).
// Pieces: k_uwp_sort (norms, rank sort, parity) -> k_uwp_scale_div (permuted + scaled copy) -> k_qrp<PIVOT = 0> -> k_uwp_finish (V, D, sign) -> k_formq.
#pragma once
#include "alf_la_host.cuh"

// one CTA per matrix.  xnorm: N2 doubles, ivpt: N2 ints (0-based: original index of the column at sorted position r), psign: parity
template <typename T>
__global__ void __launch_bounds__(256) k_uwp_sort(const T* __restrict__ A, int n1, int n2, long sA, double* __restrict__ xnorm, int* __restrict__ ivpt, double* __restrict__ psign) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* xs = reinterpret_cast<double*>(smem_raw); int* ip = reinterpret_cast<int*>(xs + n2); int* vis = ip + n2;
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  A += (long)b * sA; xnorm += (long)b * n2; ivpt += (long)b * n2;
  for (int c = tid; c < n2; c += nthr) {            // rows in order, as the reference's double loop (:134-139)
    double x = 0.0; const T* col = A + (long)c * n1;
    for (int i = 0; i < n1; ++i) x += abs2_(col[i]);
    xs[c] = x; xnorm[c] = x;
  }
  __syncthreads();
  for (int c = tid; c < n2; c += nthr) {            // selection sort of :141-153 as a rank: larger first, lower index first on ties
    const double x = xs[c]; int r = 0;
    for (int j = 0; j < n2; ++j) { const double y = xs[j]; r += (y > x || (y == x && j < c)) ? 1 : 0; }
    ip[r] = c;
  }
  __syncthreads();
  for (int c = tid; c < n2; c += nthr) { ivpt[c] = ip[c]; vis[c] = 0; }
  __syncthreads();
  if (tid == 0) {                                    // Pivot_Phase: every cycle of even length flips the sign (QDRP_decompose_mod.F90:103-126)
    double sg = 1.0;
    for (int i = 0; i < n2; ++i) if (!vis[i]) { int next = i, len = 0; while (!vis[next]) { ++len; vis[next] = 1; next = ip[next]; } if ((len & 1) == 0) sg = -sg; }
    psign[b] = sg;
  }
}
// permuted and scaled copy A1(:, r) = A(:, IVPT(r)) / XNORM(IVPT(r)).  The reference divides (A/XNORM), it does not multiply by the reciprocal: keep the division so that the scaled columns are bit-identical.
template <typename T>
__global__ void __launch_bounds__(256) k_uwp_scale_div(const T* __restrict__ A, int n1, int n2, long sA, T* __restrict__ A1, long sA1,
                                                       const double* __restrict__ xnorm, const int* __restrict__ ivpt) {
  const int b = blockIdx.y; A += (long)b * sA; A1 += (long)b * sA1; xnorm += (long)b * n2; ivpt += (long)b * n2;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n1 * n2; e += (long)gridDim.x * blockDim.x) {
    const int i = (int)(e % n1), r = (int)(e / n1), k = ivpt[r]; const double x = xnorm[k]; const T a = A[i + (long)k * n1];
    A1[e] = make_<T>(real_(a) / x, imag_(a) / x);
  }
}

// QR holds R (unit-scaled by k_qrp: R(i, i:) / |R_ii|) on and above the diagonal, Dq = |R_ii|.
// V(I, J) = V1(I, IVPTM1(J)),  V1(I, Jp) = s^[I == 0] * R(I, Jp) * XNORM(IVPT(Jp)) / XNORM(IVPT(I)) (Jp > I), R(I, I) (Jp == I), 0 below;
// D(I) = Dq(I) * XNORM(IVPT(I));  s = sign(prod R_ii) * parity is also the scale of U(:, 1) (colscale of k_formq).
template <typename T>
__global__ void __launch_bounds__(256) k_uwp_finish(const T* __restrict__ QR, int n1, int n2, long sQ, const double* __restrict__ Dq, const QrOut* __restrict__ qo,
                                                    const double* __restrict__ xnorm, const int* __restrict__ ivpt, const double* __restrict__ psign,
                                                    T* __restrict__ V, long sV, double* __restrict__ D, cplx* __restrict__ colscale) {
  const int b = blockIdx.y; QR += (long)b * sQ; V += (long)b * sV; xnorm += (long)b * n2; ivpt += (long)b * n2; Dq += (long)b * n2; D += (long)b * n2;
  const double s = ((qo[b].diag_phase.x < 0.0) ? -1.0 : 1.0) * psign[b];
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n2 * n2; e += (long)gridDim.x * blockDim.x) {
    const int I = (int)(e % n2), Jp = (int)(e / n2), J = ivpt[Jp];          // column Jp of V1 lands in column IVPT(Jp) of V
    T v = zero_<T>();
    if (Jp >= I) {
      v = QR[I + (long)Jp * n1];
      if (I == 0) v = v * s;
      if (Jp > I) v = (v * xnorm[J]) * (1.0 / xnorm[ivpt[I]]);
    }
    V[I + (long)J * n2] = v;
    if (Jp == 0) D[I] = Dq[I] * xnorm[ivpt[I]];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) colscale[b] = cplx(s, 0.0);
}

// A, U: n1 x n2 (U receives the result, A is not modified), V: n2 x n2, D: n2 per matrix; all contiguous batches
template <typename T>
static void la_udv_wrap_pivot(cudaStream_t st, const T* A, T* U, double* D, T* V, int n1, int n2, int batch) {
  if (n2 > n1) throw CudaError("UDV_Wrap_Pivot: N2 > N1");
  double *xnorm = nullptr, *psign = nullptr, *Dq = nullptr; int *ivpt = nullptr, *jp = nullptr; T* tau = nullptr; QrOut* qo = nullptr; cplx* cs = nullptr;
  CK(cudaMallocAsync(&xnorm, sizeof(double) * n2 * batch, st)); CK(cudaMallocAsync(&psign, sizeof(double) * batch, st)); CK(cudaMallocAsync(&Dq, sizeof(double) * n2 * batch, st));
  CK(cudaMallocAsync(&ivpt, sizeof(int) * n2 * batch, st)); CK(cudaMallocAsync(&jp, sizeof(int) * n2 * batch, st)); CK(cudaMallocAsync(&tau, sizeof(T) * n2 * batch, st));
  CK(cudaMallocAsync(&qo, sizeof(QrOut) * batch, st)); CK(cudaMallocAsync(&cs, sizeof(cplx) * batch, st));
  const long sA = (long)n1 * n2, sV = (long)n2 * n2;
  const size_t smem = (sizeof(double) + 2 * sizeof(int)) * (size_t)n2;
  KL(KC_EW, st, k_uwp_sort<T><<<batch, 256, smem, st>>>(A, n1, n2, sA, xnorm, ivpt, psign));
  KL(KC_EW, st, k_uwp_scale_div<T><<<dim3(ew_blocks(sA), batch), 256, 0, st>>>(A, n1, n2, sA, U, sA, xnorm, ivpt));
  launch_qrp<T, 0>(st, U, n1, n2, n1, sA, tau, n2, jp, n2, Dq, n2, qo, batch);
  KL(KC_EW, st, k_uwp_finish<T><<<dim3(ew_blocks(sV), batch), 256, 0, st>>>(U, n1, n2, sA, Dq, qo, xnorm, ivpt, psign, V, sV, D, cs));
  launch_formq<T>(st, U, n1, n2, n1, sA, tau, n2, cs, batch);
  CK(cudaFreeAsync(xnorm, st)); CK(cudaFreeAsync(psign, st)); CK(cudaFreeAsync(Dq, st)); CK(cudaFreeAsync(ivpt, st)); CK(cudaFreeAsync(jp, st));
  CK(cudaFreeAsync(tau, st)); CK(cudaFreeAsync(qo, st)); CK(cudaFreeAsync(cs, st));
}
