// Application of sequences of small operators  exp(-dtau T_bond), exp(g phi(s) O_n)  to N x N matrices.
// Replaces Hop_mod_mmthr/_m1/mmthl/_m1/mmthlc/Symm (Prog/Hop_mod.F90:143-299, ZDSLSYMM/ZSLHEMM of
// Libraries/Modules/Mat_subroutines_mod.F90:518,967) and Op_mmultR/Op_mmultL (Prog/Operator_mod.F90:555-717),
// hence the bodies of WRAPUR/WRAPUL (Prog/wrapur_mod.F90:110-123, wrapul_mod.F90:117-129) and
// PROPR/PROPRM1 (Prog/tau_m_mod.F90:215-263).
//
// B200 mapping: a left multiplication acts on every column independently, a right multiplication on every
// row independently, so a CTA stages a panel of PW columns (or rows) in shared memory, applies ALL operators of
// ALL requested time slices to it, and writes it back: one HBM read + one write of the matrix per wrap
// (2*w*N^2 bytes) instead of one per operator.  Operators with disjoint support are grouped in
// "levels" on the host (checkerboard families are levels by construction); inside a level every (operator, lane)
// pair is an independent work item, levels are separated by __syncthreads().
#pragma once
#include "alf_types.cuh"

#define ALF_KMAX 4           // largest small-operator dimension handled by the op lists (bond ops: 2)
#define ALF_NVAR 5           // field values sp = -2..2 -> table index sp+2 (types 1 and 2)
#define ALF_FMAX 4           // max number of computed flavors

struct OpListDev {
  int n_ops, n_levels, nvar;        // nvar = 1: fixed matrices (hopping); nvar = 5: field dependent (vertices)
  const int* level_start;           // n_levels + 1
  const int* k;                     // n_ops
  const int* P;                     // n_ops * ALF_KMAX   (0-based)
  const int* fidx;                  // n_ops : field index n (0-based) or -1
  const void* mat;                  // n_ops * nvar * KMAX*KMAX entries of T, column-major a + b*KMAX
};

enum {  // which operator lists a launch applies (per slice nt in [nt_a, nt_b])
  MODE_WRAPUR = 0,    // left:  for nt ascending : e^{-dtau T} (mmthr) then e^{V_n} n = 1..M        (also PROPR)
  MODE_WRAPUL = 1,    // left:  for nt descending: (e^{V_n})^H n = M..1 then mmthlc
  MODE_TL_FWD = 2,    // left:  mmthr
  MODE_TL_INV = 3,    // left:  mmthr_m1
  MODE_TL_HALF = 4,   // left:  Symm half step  (mat_1D2, nc = Ncheck..1)
  MODE_TR_FWD = 5,    // right: mmthl
  MODE_TR_INV = 6,    // right: mmthl_m1
  MODE_TR_HALFINV = 7,// right: Symm half step inverse (invmat_1D2)
  MODE_PROPRM1 = 8,   // right: mmthl_m1 then e^{-V_n} n = 1..M
  MODE_TL_C = 9       // left:  mmthlc
};

enum { L_TL_FWD = 0, L_TL_INV, L_TL_C, L_TL_HALF, L_TR_FWD, L_TR_INV, L_TR_HALFINV, L_VL_N, L_VL_C, L_VR_INV, L_COUNT };

struct ModelDev {
  OpListDev lists[L_COUNT][ALF_FMAX];
};

template <typename T>
__device__ __forceinline__ void apply_list(T* __restrict__ S, int ldp, int pw, const OpListDev& L, const int8_t* __restrict__ fld) {
  const T* mats = reinterpret_cast<const T*>(L.mat);
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int lv = 0; lv < L.n_levels; ++lv) {
    const int a0 = L.level_start[lv], a1 = L.level_start[lv + 1];
    const int work = (a1 - a0) * pw;
    for (int idx = tid; idx < work; idx += nthr) {
      const int o = a0 + idx / pw, lane = idx % pw;
      const int k = L.k[o];
      int var = 0;
      if (L.nvar > 1) var = (int)fld[L.fidx[o]] + 2;
      const T* A = mats + ((long)o * L.nvar + var) * (ALF_KMAX * ALF_KMAX);
      const int* P = L.P + o * ALF_KMAX;
      if (k == 1) {
        T* p = S + (long)P[0] * ldp + lane;
        *p = A[0] * (*p);
      } else if (k == 2) {
        T* p0 = S + (long)P[0] * ldp + lane; T* p1 = S + (long)P[1] * ldp + lane;
        T v0 = *p0, v1 = *p1;
        *p0 = A[0] * v0 + A[ALF_KMAX] * v1;
        *p1 = A[1] * v0 + A[ALF_KMAX + 1] * v1;
      } else {
        T v[ALF_KMAX], r[ALF_KMAX];
#pragma unroll
        for (int a = 0; a < ALF_KMAX; ++a) v[a] = (a < k) ? S[(long)P[a] * ldp + lane] : zero_<T>();
#pragma unroll
        for (int a = 0; a < ALF_KMAX; ++a) {
          T s = zero_<T>();
#pragma unroll
          for (int b = 0; b < ALF_KMAX; ++b) if (b < k) fma_(s, A[a + b * ALF_KMAX], v[b]);
          r[a] = s;
        }
#pragma unroll
        for (int a = 0; a < ALF_KMAX; ++a) if (a < k) S[(long)P[a] * ldp + lane] = r[a];
      }
    }
    __syncthreads();
  }
}

// SIDE 0: panel of PW columns [c0, c0+PW) of M;  S[i][j] = M(i, c0+j)
// SIDE 1: panel of PW rows    [r0, r0+PW) of M;  S[i][j] = M(r0+j, i)   (right multiplication = left on the transpose)
// grid = (ceil(nvec/PW), n_matrices); matrix b belongs to chain b / F, flavor b % F.
template <typename T, int SIDE>
__global__ void __launch_bounds__(256) k_apply_ops(T* __restrict__ M, long sM, int N, int nvec, int pw_max, ModelDev md, int F, int mode,
                                                   int nt_a, int nt_b, const int8_t* __restrict__ fields, int Ltrot, int n_opv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);
  const int b = blockIdx.y, chain = b / F, f = b % F;
  M += (long)b * sM;
  const int v0 = blockIdx.x * pw_max;
  const int pw = min(pw_max, nvec - v0);
  const int ldp = pw_max + 1;
  const int tid = threadIdx.x, nthr = blockDim.x;
  // ---- stage
  if (SIDE == 0) {
    for (int e = tid; e < N * pw; e += nthr) { int i = e % N, j = e / N; S[(long)i * ldp + j] = M[i + (long)(v0 + j) * N]; }
  } else {
    for (int e = tid; e < N * pw; e += nthr) { int j = e % pw, i = e / pw; S[(long)i * ldp + j] = M[(v0 + j) + (long)i * N]; }
  }
  __syncthreads();
  const int8_t* fbase = fields ? fields + (long)chain * Ltrot * n_opv : nullptr;
  switch (mode) {
    case MODE_WRAPUR:
      for (int nt = nt_a; nt <= nt_b; ++nt) {
        apply_list<T>(S, ldp, pw, md.lists[L_TL_FWD][f], nullptr);
        apply_list<T>(S, ldp, pw, md.lists[L_VL_N][f], fbase + (long)(nt - 1) * n_opv);
      }
      break;
    case MODE_WRAPUL:
      for (int nt = nt_b; nt >= nt_a; --nt) {
        apply_list<T>(S, ldp, pw, md.lists[L_VL_C][f], fbase + (long)(nt - 1) * n_opv);
        apply_list<T>(S, ldp, pw, md.lists[L_TL_C][f], nullptr);
      }
      break;
    case MODE_TL_FWD: apply_list<T>(S, ldp, pw, md.lists[L_TL_FWD][f], nullptr); break;
    case MODE_TL_INV: apply_list<T>(S, ldp, pw, md.lists[L_TL_INV][f], nullptr); break;
    case MODE_TL_C: apply_list<T>(S, ldp, pw, md.lists[L_TL_C][f], nullptr); break;
    case MODE_TL_HALF: apply_list<T>(S, ldp, pw, md.lists[L_TL_HALF][f], nullptr); break;
    case MODE_TR_FWD: apply_list<T>(S, ldp, pw, md.lists[L_TR_FWD][f], nullptr); break;
    case MODE_TR_INV: apply_list<T>(S, ldp, pw, md.lists[L_TR_INV][f], nullptr); break;
    case MODE_TR_HALFINV: apply_list<T>(S, ldp, pw, md.lists[L_TR_HALFINV][f], nullptr); break;
    case MODE_PROPRM1:
      for (int nt = nt_a; nt <= nt_b; ++nt) {
        apply_list<T>(S, ldp, pw, md.lists[L_TR_INV][f], nullptr);
        apply_list<T>(S, ldp, pw, md.lists[L_VR_INV][f], fbase + (long)(nt - 1) * n_opv);
      }
      break;
  }
  // ---- write back
  if (SIDE == 0) {
    for (int e = tid; e < N * pw; e += nthr) { int i = e % N, j = e / N; M[i + (long)(v0 + j) * N] = S[(long)i * ldp + j]; }
  } else {
    for (int e = tid; e < N * pw; e += nthr) { int j = e % pw, i = e / pw; M[(v0 + j) + (long)i * N] = S[(long)i * ldp + j]; }
  }
}
