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
// pair is an independent work item, levels are separated by __syncthreads().  See k_apply_ops below for the mapping.
#pragma once
#include "alf_types.cuh"

#define ALF_KMAX 4           // largest small-operator dimension handled by the op lists (bond ops: 2)
#define ALF_NVAR 5           // field values sp = -2..2 -> table index sp+2 (types 1 and 2)
#define ALF_FMAX 4           // max number of computed flavors
#define ALF_GM_MAXLEN 16      // longest Flip_list of one global-in-slice move (Wrapgr_Random_update)

struct OpListDev {
  int n_ops, n_levels, nvar;        // nvar = 1: fixed matrices (hopping); nvar = 5: field dependent (vertices)
  const int* level_start;           // n_levels + 1  (levels are cut into chunks of <= OPS_CH operators)
  const int* k;                     // n_ops
  const int* P;                     // n_ops * ALF_KMAX   (0-based)
  const int* fidx;                  // n_ops : field index n (0-based) or -1
  const void* mat;                  // n_ops * nvar * KMAX*KMAX entries of T, column-major a + b*KMAX
  const unsigned char* uniform;     // n_levels: 1 if every operator of the chunk is a k = 2 operator with one and the same matrix
  const unsigned char* cont;        // n_ops (vertex lists only): 1 = continuous field (type 3, k = 1): mat slot 0 holds the coefficient c, the factor is exp(c phi)
  long mat_nt_stride;               // time-dependent couplings g_t (Operator_mod.F90:66): one table per time slice, mat + (nt - 1) * mat_nt_stride; else 0
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
  const double* fields_c;           // continuous fields [chain][nt][n] (type 3 vertices), or nullptr
};

// Every launch executes a "program": for each requested time slice, up to two operator lists in a fixed order.  The lists are
// cut on the host into CHUNKS of at most OPS_CH operators with pairwise disjoint support (checkerboard families are such
// sets by construction; longer ones are split), so inside a chunk every (operator, panel lane) pair is independent work.
//  * lanes of a warp = the 32 columns (rows) of the panel, a warp walks over the operators of the chunk: the operator
//    descriptor is warp-uniform and comes from shared memory (broadcast), the data accesses S[P * ldp + lane] are conflict free;
//  * the descriptors of chunk t+1 (support, k, and the k x k matrix of the field value found in nsigma for vertex lists) are
//    fetched from global memory into registers while chunk t is processed and land in the other half of a double buffer:
//    one block-wide barrier per chunk.
#define OPS_CH 64            // operators per chunk
#define OPS_PW 32            // panel width = warp size
#define OPS_NT 256           // threads per CTA (8 warps, three CTAs per SM at N = 256; measured: 512 threads and 128-operator chunks are both slower)

// Panel layout: element (row i, lane j) at S[i * 32 + (j ^ (i & 31))].  Rows are aligned 256-byte (512-byte) lines, so a warp's access to one
// row costs the minimum number of shared-memory wavefronts (with the former odd row stride of 33 every row straddled three 128-byte bank rows
// instead of two), and the XOR keeps the transposing accesses of the staging (32 consecutive rows, one lane) conflict free as well.
__device__ __forceinline__ int ops_sw(int off, int lane) { return off + (lane ^ ((off >> 5) & 31)); }

// Shared-memory descriptor of one operator: x = (k << 28) | element offset of row P[0] in the panel, y, z, w = offsets of P[1..3];
// its k x k matrix sits in dM[o << 2 LK] with leading dimension 1 << LK.
template <typename T, int LK>
__device__ __forceinline__ void ops_process_chunk(T* __restrict__ S, bool lane_ok, int cnt, const int4* __restrict__ dP, const T* __restrict__ dM, bool uniform) {
  constexpr int kk = 1 << LK, ms = 1 << (2 * LK);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // element (row offset off = P * 32, lane): rows are aligned 32-element lines, the lane index is XOR-swizzled with the row (ops_sw)
  #define Sl(off) S[ops_sw(off, lane)]
  if (LK >= 1 && uniform) {                   // translation-invariant checkerboard family: one matrix for the whole chunk
    const T a00 = dM[0], a10 = dM[1], a01 = dM[kk], a11 = dM[kk + 1];
    for (int o = warp; o < cnt; o += 2 * nw) {
      const int o2 = (o + nw < cnt) ? o + nw : o;
      const int4 P = dP[o], Q = dP[o2];
      const int px = P.x & 0x0fffffff, qx = Q.x & 0x0fffffff;
      if (lane_ok) {
        const T v0 = Sl(px), v1 = Sl(P.y), w0 = Sl(qx), w1 = Sl(Q.y);
        Sl(px) = a00 * v0 + a01 * v1; Sl(P.y) = a10 * v0 + a11 * v1;
        if (o2 != o) { Sl(qx) = a00 * w0 + a01 * w1; Sl(Q.y) = a10 * w0 + a11 * w1; }
      }
    }
    return;
  }
  for (int o = warp; o < cnt; o += 2 * nw) {
    const int o2 = (o + nw < cnt) ? o + nw : o;          // second operator of this iteration (same as the first if there is none)
    const bool two = o2 != o;
    const int4 P = dP[o], Q = dP[o2];
    const int k = ((unsigned)P.x) >> 28, k2 = ((unsigned)Q.x) >> 28;
    const int px = P.x & 0x0fffffff, qx = Q.x & 0x0fffffff;
    const T* A = dM + o * ms; const T* Bm = dM + o2 * ms;
    if (LK >= 1 && k == 2 && k2 == 2) {       // common case (bond operators): two independent operators in flight
      const T a00 = A[0], a10 = A[1], a01 = A[kk], a11 = A[kk + 1];
      const T b00 = Bm[0], b10 = Bm[1], b01 = Bm[kk], b11 = Bm[kk + 1];
      if (lane_ok) {
        const T v0 = Sl(px), v1 = Sl(P.y), w0 = Sl(qx), w1 = Sl(Q.y);
        Sl(px) = a00 * v0 + a01 * v1; Sl(P.y) = a10 * v0 + a11 * v1;
        if (two) { Sl(qx) = b00 * w0 + b01 * w1; Sl(Q.y) = b10 * w0 + b11 * w1; }
      }
      continue;
    }
    if (k == 1 && k2 == 1) {                  // diagonal single-site vertices
      const T a = A[0], bq = Bm[0];
      if (lane_ok) { const T v0 = Sl(px), w0 = Sl(qx); Sl(px) = a * v0; if (two) Sl(qx) = bq * w0; }
      continue;
    }
    for (int h = 0; h < (two ? 2 : 1); ++h) {
      const int4 PP = h ? Q : P; const int kc = h ? k2 : k; const T* AA = h ? Bm : A;
      if (!lane_ok) continue;
      const int Pa[ALF_KMAX] = {PP.x & 0x0fffffff, PP.y, PP.z, PP.w};
      T v[kk], r[kk];
#pragma unroll
      for (int a = 0; a < kk; ++a) v[a] = (a < kc) ? Sl(Pa[a]) : zero_<T>();
#pragma unroll
      for (int a = 0; a < kk; ++a) {
        T sacc = zero_<T>();
#pragma unroll
        for (int bq = 0; bq < kk; ++bq) if (a < kc && bq < kc) fma_(sacc, AA[a + bq * kk], v[bq]);
        r[a] = sacc;
      }
#pragma unroll
      for (int a = 0; a < kk; ++a) if (a < kc) Sl(Pa[a]) = r[a];
    }
  }
}

// SIDE 0: panel of OPS_PW columns [v0, v0+pw) of M;  S[i][j] = M(i, v0+j)
// SIDE 1: panel of OPS_PW rows    [v0, v0+pw) of M;  S[i][j] = M(v0+j, i)   (right multiplication = left on the transpose)
// grid = (ceil(nvec/OPS_PW), n_matrices); matrix b belongs to chain b / F, flavor b % F.  LK = log2 of the largest operator size.
template <typename T, int SIDE, int LK>
__global__ void __launch_bounds__(OPS_NT) k_apply_ops(T* M, long sM, int N, int nvec, ModelDev md, int F, int mode,
                                                   int nt_a, int nt_b, const int8_t* __restrict__ fields, int Ltrot, int n_opv, T* Mout) {   // Mout: result buffer (may be M)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int ldp = OPS_PW, ms = 1 << (2 * LK), kk = 1 << LK;
  constexpr int MPT = (OPS_CH * ms + OPS_NT - 1) / OPS_NT;       // descriptor-matrix entries prefetched per thread
  T* S = reinterpret_cast<T*>(smem_raw);
  T* dMb = S + ((N * ldp + 1) & ~1);                  // keeps the int4 descriptor arrays 16-byte aligned
  int4* dPb = reinterpret_cast<int4*>(dMb + 2 * OPS_CH * ms);
  const int b = blockIdx.y, chain = b / F, f = b % F;
  M += (long)b * sM; Mout += (long)b * sM;
  const int v0 = blockIdx.x * OPS_PW;
  const int pw = min(OPS_PW, nvec - v0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // ---- the program of this launch: per slice, list a then list b
  int li0 = 0, li1 = -1, uf0 = 0, uf1 = 0, dir = 1;
  switch (mode) {
    case MODE_WRAPUR: li0 = L_TL_FWD; li1 = L_VL_N; uf1 = 1; break;
    case MODE_WRAPUL: li0 = L_VL_C; uf0 = 1; li1 = L_TL_C; dir = -1; break;
    case MODE_TL_FWD: li0 = L_TL_FWD; break;
    case MODE_TL_INV: li0 = L_TL_INV; break;
    case MODE_TL_C: li0 = L_TL_C; break;
    case MODE_TL_HALF: li0 = L_TL_HALF; break;
    case MODE_TR_FWD: li0 = L_TR_FWD; break;
    case MODE_TR_INV: li0 = L_TR_INV; break;
    case MODE_TR_HALFINV: li0 = L_TR_HALFINV; break;
    case MODE_PROPRM1: li0 = L_TR_INV; li1 = L_VR_INV; uf1 = 1; break;
  }
  const OpListDev La = md.lists[li0][f];
  const OpListDev Lb = md.lists[li1 >= 0 ? li1 : li0][f];
  const int nch0 = La.n_levels, nch1 = (li1 >= 0) ? Lb.n_levels : 0, per = nch0 + nch1;
  const int ns = (uf0 || uf1) ? (nt_b - nt_a + 1) : 1, total = ns * per;
  const int8_t* fbase = fields ? fields + (long)chain * Ltrot * n_opv : nullptr;

  // Descriptor prefetch, two chunks deep: the chunk boundaries ("meta": first operator, count, uniform flag, field slice) of chunk
  // t + 3 and the operator data of chunk t + 2 are requested from global memory while chunk t is processed, and the data of chunk
  // t + 1 (requested one step earlier) is committed to the other half of the shared-memory double buffer afterwards.
  struct Pref { int4 rP; T rM[MPT]; int rcnt; bool runi; };
  Pref pf0, pf1; pf0.rcnt = pf1.rcnt = 0; pf0.runi = pf1.runi = false;
  int f_sl = 0, f_r = 0, f_n = 0;                     // cursor of the next meta load
  bool m_second = false; int m_a0 = 0, m_cnt = 0, m_nt = 1; bool m_uni = false; const int8_t* m_fld = nullptr; const double* m_fc = nullptr;
  auto meta_load = [&]() {
    if (f_n >= total) { m_cnt = 0; return; }
    m_second = f_r >= nch0;
    const OpListDev& L = m_second ? Lb : La;
    const int c = m_second ? f_r - nch0 : f_r;
    m_a0 = L.level_start[c]; m_cnt = L.level_start[c + 1] - m_a0; m_uni = L.uniform[c] != 0;
    const int nt = (dir > 0) ? nt_a + f_sl : nt_b - f_sl;
    m_fld = ((m_second ? uf1 : uf0) && fbase) ? fbase + (nt - 1) * n_opv : nullptr;
    m_fc = (m_fld && md.fields_c) ? md.fields_c + ((long)chain * Ltrot + (nt - 1)) * n_opv : nullptr;
    m_nt = nt;
    ++f_n; if (++f_r == per) { f_r = 0; ++f_sl; }
  };
  auto data_issue = [&](Pref& pf) {                   // uses the meta registers loaded one step earlier
    const OpListDev& L = m_second ? Lb : La;
    pf.rcnt = m_cnt; pf.runi = m_uni;
    if (tid < m_cnt) {
      const int4 p = reinterpret_cast<const int4*>(L.P)[m_a0 + tid];
      pf.rP = make_int4((p.x * ldp) | (L.k[m_a0 + tid] << 28), p.y * ldp, p.z * ldp, p.w * ldp);
    }
    const T* mats = reinterpret_cast<const T*>(L.mat) + (long)(m_nt - 1) * L.mat_nt_stride;
#pragma unroll
    for (int u = 0; u < MPT; ++u) {
      const int e = tid + u * OPS_NT;
      if (e < m_cnt * ms) {
        const int o = e >> (2 * LK), rr = e & (ms - 1), a = rr & (kk - 1), bb = rr >> LK, og = m_a0 + o;
        int var = 0;
        if (L.nvar > 1) {
          if (m_fc && L.cont[og]) {        // continuous field: exp(c phi) on the fly (Op_exp for types 3, 4: Operator_mod.F90:585-600)
            pf.rM[u] = (rr == 0) ? exp_(mats[(og * L.nvar) * (ALF_KMAX * ALF_KMAX)] * m_fc[L.fidx[og]]) : zero_<T>();
            continue;
          }
          var = (int)m_fld[L.fidx[og]] + 2;
        }
        pf.rM[u] = mats[(og * L.nvar + var) * (ALF_KMAX * ALF_KMAX) + a + bb * ALF_KMAX];
      }
    }
  };
  auto commit = [&](const Pref& pf, int buf) {
    if (tid < pf.rcnt) dPb[buf * OPS_CH + tid] = pf.rP;
#pragma unroll
    for (int u = 0; u < MPT; ++u) { const int e = tid + u * OPS_NT; if (e < pf.rcnt * ms) dMb[buf * OPS_CH * ms + e] = pf.rM[u]; }
  };
  meta_load();
  if (total > 0) { data_issue(pf0); meta_load(); }
  if (total > 1) { data_issue(pf1); meta_load(); }
  // ---- stage the panel
  if (SIDE == 0) {
    for (int j = warp; j < pw; j += nw) { const T* col = M + (long)(v0 + j) * N; for (int i = lane; i < N; i += 32) S[ops_sw(i * ldp, j)] = col[i]; }
  } else {
    if (lane < pw) { const T* src = M + v0 + lane; for (int i = warp; i < N; i += nw) S[ops_sw(i * ldp, lane)] = src[(long)i * N]; }
  }
  int cnt_cur = pf0.rcnt; bool uni_cur = pf0.runi;
  if (total > 0) commit(pf0, 0);
  __syncthreads();
  for (int t = 0; t < total; t += 2) {
    // even step: chunk t sits in buffer 0; pf0 is free, pf1 carries chunk t + 1, the meta registers describe chunk t + 2
    if (t + 2 < total) { data_issue(pf0); meta_load(); }
    ops_process_chunk<T, LK>(S, lane < pw, cnt_cur, dPb, dMb, uni_cur);
    if (t + 1 < total) { commit(pf1, 1); cnt_cur = pf1.rcnt; uni_cur = pf1.runi; }
    __syncthreads();
    if (t + 1 >= total) break;
    // odd step: chunk t + 1 in buffer 1; pf1 is free, pf0 carries chunk t + 2, the meta registers describe chunk t + 3
    if (t + 3 < total) { data_issue(pf1); meta_load(); }
    ops_process_chunk<T, LK>(S, lane < pw, cnt_cur, dPb + OPS_CH, dMb + OPS_CH * ms, uni_cur);
    if (t + 2 < total) { commit(pf0, 0); cnt_cur = pf0.rcnt; uni_cur = pf0.runi; }
    __syncthreads();
  }
  // ---- write back
  if (SIDE == 0) {
    for (int j = warp; j < pw; j += nw) { T* col = Mout + (long)(v0 + j) * N; for (int i = lane; i < N; i += 32) col[i] = S[ops_sw(i * ldp, j)]; }
  } else {
    if (lane < pw) { T* dst = Mout + v0 + lane; for (int i = warp; i < N; i += nw) dst[(long)i * N] = S[ops_sw(i * ldp, lane)]; }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_apply_ops_fixed: the same programs for the common case that the program's hopping list consists of k = 2 bond operators only
// (checkerboard decomposition) and its vertex list, if any, of diagonal single-site factors (every Hubbard variant).
// k_apply_ops is instruction bound (ncu: 55 % issue-slot utilisation; about 45 % of its instructions are the per-chunk descriptor
// pipeline that resolves field-dependent matrices).  Here the descriptors of the WHOLE hopping list are resident in shared memory
// in their final form, and e^{V(s)} of a slice is a row scaling built once per slice from the field values.
//
// RING GROUPS.  A checkerboard family is a perfect matching of the sites; the union of two matchings A, B decomposes into even
// rings r_0 r_1 ... r_{n-1} whose bonds alternate: A = (r_2i, r_2i+1), B = (r_2i+1, r_2i+2 mod n) (square lattice, ALF's families:
// staircases of 2 L sites).  Consecutive families that use only these two matchings (1 2 | 3 4 3 | 2 1 of the symmetric Trotter
// order) form a GROUP: one thread takes one ring of one panel lane into registers (n <= 32 values), applies every family of the
// group as register rotations and stores the ring: one shared-memory pass per group instead of one per family (7 -> 3 passes per
// e^{-dtau T}), and the diagonal vertices of a slice are fused into the first / last pass of the slice as a scaling on load / store.
// Families that do not pair up (ring too long, different supports) keep the one-pass-per-family form (kind 0).
// ------------------------------------------------------------------------------------------------------------------------
struct FixFamDev { int type, uniform, mat_off, pad; };    // type 0: bonds (r_2i, r_2i+1), 1: (r_2i+1, r_2i+2 mod n); mat_off into rmat: 4 entries (uniform) or nrings * 16 * 4
struct FixGroupDev { int kind, fam0, nfam, n, nrings, nmax, ring_off, cover; };   // kind 1: ring group; ring_off: first entry of its table in `ring`; cover: the rings contain every row
// The descriptors of one list live in ONE blob that every CTA copies to shared memory together with its panel (cp.async), so the
// family loop never waits for global memory (ncu on the first version: 30 % of the stall samples were dependent descriptor loads).
// Sections (byte offsets, 16-byte aligned): groups, ring-family records, ring matrices, ring tables (32-bit entries row * 32 + (row & 31),
// OPS_RSTR per ring), and for the families outside ring groups: operator offsets (P0 * 32) | (P1 * 32) << 16,
// family starts, uniform flags, one 2 x 2 matrix per family (a00, a10, a01, a11).
struct FixListDev {
  int n_fam, n_ops, n_grp;          // n_fam = 0: the list has no fixed form
  int blob_bytes, o_grp, o_rfam, o_rmat, o_ring, o_offs, o_fs, o_uni, o_m4;
  const unsigned char* blob;
  const void* mat;                  // global: n_ops * 4 entries of T (only read for families whose operators carry different matrices)
};
struct VDiagDev { const void* tab; const int* fidx; const unsigned char* cont; long tab_nt_stride; };   // diagonal vertex list by site: tab[i * 5 + s + 2] (continuous: slot 0 = coefficient), field index or -1
struct ModelFixDev {
  FixListDev fix[L_COUNT][ALF_FMAX];
  unsigned char diag_ok[L_COUNT][ALF_FMAX];   // vertex lists: only k = 1 factors, every site at most once
  VDiagDev vd[L_COUNT][ALF_FMAX];
};
static inline size_t ops_fixed_smem(size_t sizeof_T, int N, int max_blob) { return sizeof_T * ((size_t)N * OPS_PW + N) + (size_t)max_blob + 64; }

__device__ __forceinline__ void ops_cp_async(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ops_cp_async(cplx* smem_dst, const cplx* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ops_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gsrc) : "memory");
}
#define OPS_RSTR 32          // stride of the ring tables (entries per ring, padded)

// One ring of one panel lane: n values to registers, every family of the group as register rotations, back.  FULL: n == NMAX (no predicates).
// Ring table entry t = row * 32 + (row & 31): the element of panel lane l sits at S[t ^ l] (the XOR touches the low five bits only).
template <typename T, int NMAX, bool FULL>
__device__ __forceinline__ void ops_ring_task(T* __restrict__ S, const unsigned* __restrict__ ro, int ring, int n, int nfam, const FixFamDev* __restrict__ rfam,
                                              const T* __restrict__ rmat, const T* __restrict__ pre, const T* __restrict__ post, unsigned lane) {
  const uint4* ro4 = reinterpret_cast<const uint4*>(ro);
  unsigned ad[NMAX];
#pragma unroll
  for (int q = 0; q < NMAX / 4; ++q) { const uint4 w = ro4[q]; ad[4 * q] = w.x ^ lane; ad[4 * q + 1] = w.y ^ lane; ad[4 * q + 2] = w.z ^ lane; ad[4 * q + 3] = w.w ^ lane; }
  T r[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) if (FULL || i < n) r[i] = S[ad[i]];
  if (pre) {
#pragma unroll
    for (int i = 0; i < NMAX; ++i) if (FULL || i < n) r[i] = pre[ad[i] >> 5] * r[i];
  }
  for (int fi = 0; fi < nfam; ++fi) {
    const FixFamDev fd = rfam[fi];
    const T* m = rmat + fd.mat_off;
    if (fd.uniform) {
      const T a00 = m[0], a10 = m[1], a01 = m[2], a11 = m[3];
#define ALF_ROT(x, y) { const T t0_ = a00 * (x) + a01 * (y); (y) = a10 * (x) + a11 * (y); (x) = t0_; }
      if (fd.type == 0) {
#pragma unroll
        for (int i = 0; i < NMAX / 2; ++i) if (FULL || 2 * i + 1 < n) ALF_ROT(r[2 * i], r[2 * i + 1])
      } else {
#pragma unroll
        for (int i = 0; i < NMAX / 2; ++i) {
          if (2 * i + 2 < NMAX && (FULL || 2 * i + 2 < n)) ALF_ROT(r[2 * i + 1], r[(2 * i + 2) % NMAX])
          else if (FULL ? (2 * i + 2 == NMAX) : (2 * i + 2 == n)) ALF_ROT(r[2 * i + 1], r[0])
        }
      }
#undef ALF_ROT
    } else {
      m += (long)ring * (OPS_RSTR / 2) * 4;
#define ALF_ROT(x, y, i) { const T a00 = m[4 * (i)], a10 = m[4 * (i) + 1], a01 = m[4 * (i) + 2], a11 = m[4 * (i) + 3]; const T t0_ = a00 * (x) + a01 * (y); (y) = a10 * (x) + a11 * (y); (x) = t0_; }
      if (fd.type == 0) {
#pragma unroll
        for (int i = 0; i < NMAX / 2; ++i) if (FULL || 2 * i + 1 < n) ALF_ROT(r[2 * i], r[2 * i + 1], i)
      } else {
#pragma unroll
        for (int i = 0; i < NMAX / 2; ++i) {
          if (2 * i + 2 < NMAX && (FULL || 2 * i + 2 < n)) ALF_ROT(r[2 * i + 1], r[(2 * i + 2) % NMAX], i)
          else if (FULL ? (2 * i + 2 == NMAX) : (2 * i + 2 == n)) ALF_ROT(r[2 * i + 1], r[0], i)
        }
      }
#undef ALF_ROT
    }
  }
  if (post) {
#pragma unroll
    for (int i = 0; i < NMAX; ++i) if (FULL || i < n) r[i] = post[ad[i] >> 5] * r[i];
  }
#pragma unroll
  for (int i = 0; i < NMAX; ++i) if (FULL || i < n) S[ad[i]] = r[i];
}

#define OPSF_NT 512          // threads of the persistent kernel at large N (one CTA per SM, 128 registers per thread)
static inline size_t ops_fixed_smem2(size_t sizeof_T, int N, int max_blob, int F, int nbuf) {
  return sizeof_T * ((size_t)N * OPS_PW * nbuf + N) + (size_t)(max_blob + 16) * F + 64;
}

// PERSISTENT form: the CTAs walk over the panels of the whole batch (panel q = blockIdx.x + k gridDim.x: matrix q / npan, panel q % npan);
// with nbuf = 2 the panel of step k + 1 is in flight (cp.async) while the families run on panel k and panel k - 1 drains to HBM, so the
// load -> compute -> store phases of the resident CTAs no longer run in lockstep (ncu on the one-panel-per-CTA form: 2 TB/s of DRAM
// traffic, shared-memory pipe 43 % busy, i.e. neither resource saturated).  The descriptor blobs of all flavors are staged once per CTA.
template <typename T, int SIDE, int NMAX>
__global__ void __launch_bounds__((NMAX > 16 || sizeof(T) > 8) ? 256 : OPSF_NT, 1) k_apply_ops_fixed(
    T* M, long sM, int N, int nvec, ModelDev md, ModelFixDev mf, int F, int mode, int nt_a, int nt_b, const int8_t* __restrict__ fields, int Ltrot, int n_opv, T* Mout,
    int npan, int total, int nbuf, int blob_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int ldp = OPS_PW;
  T* Sb = reinterpret_cast<T*>(smem_raw);
  T* dsc = Sb + (long)N * ldp * nbuf;
  unsigned char* blob0 = reinterpret_cast<unsigned char*>(dsc + N);
  blob0 = smem_raw + (((size_t)(blob0 - smem_raw) + 15) & ~(size_t)15);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nthr = blockDim.x;
  int li0 = 0, li1 = -1, uf0 = 0, uf1 = 0, dir = 1;
  switch (mode) {
    case MODE_WRAPUR: li0 = L_TL_FWD; li1 = L_VL_N; uf1 = 1; break;
    case MODE_WRAPUL: li0 = L_VL_C; uf0 = 1; li1 = L_TL_C; dir = -1; break;
    case MODE_TL_FWD: li0 = L_TL_FWD; break;
    case MODE_TL_INV: li0 = L_TL_INV; break;
    case MODE_TL_C: li0 = L_TL_C; break;
    case MODE_TL_HALF: li0 = L_TL_HALF; break;
    case MODE_TR_FWD: li0 = L_TR_FWD; break;
    case MODE_TR_INV: li0 = L_TR_INV; break;
    case MODE_TR_HALFINV: li0 = L_TR_HALFINV; break;
    case MODE_PROPRM1: li0 = L_TR_INV; li1 = L_VR_INV; uf1 = 1; break;
  }
  const int lt = uf0 ? li1 : li0;                              // the program's hopping list (exactly one per mode)
  const bool has_v = uf0 || uf1;
  const int ns = has_v ? (nt_b - nt_a + 1) : 1;
  const int lv = uf0 ? li0 : li1;                              // the vertex list, if any
  // ---- panel q -> buffer: LDGSTS, every thread has all of its copies in flight; the XOR-swizzled rows keep the transposing writes of the
  // left-multiplication panel conflict free
  auto stage = [&](int q, T* S) {
    const int b = q / npan, v0 = (q - b * npan) * OPS_PW, pw = min(OPS_PW, nvec - v0);
    const T* Mb = M + (long)b * sM;
    if (SIDE == 0) {      // lane = row within a block of 32 rows: S[(l + 32 k) * 32 + (j ^ l)] = base + 1024 k
      for (int j = warp; j < pw; j += nw) {
        const T* col = Mb + (long)(v0 + j) * N + lane; T* sp = S + (lane * ldp + (j ^ lane));
        for (int i = lane; i < N; i += 32, col += 32, sp += 32 * ldp) ops_cp_async(sp, col);
      }
    } else {
      if (lane < pw) { const T* src = Mb + v0 + lane + (long)warp * N; const long st = (long)nw * N; for (int i = warp; i < N; i += nw, src += st) ops_cp_async(&S[i * ldp + (lane ^ (i & 31))], src); }
    }
  };
  const int G = gridDim.x;
  int q = blockIdx.x;
  for (int f = 0; f < F; ++f) { const FixListDev& FLf = mf.fix[lt][f]; for (int e = tid * 16; e < FLf.blob_bytes; e += nthr * 16) ops_cp_async16(blob0 + (long)f * blob_stride + e, FLf.blob + e); }
  if (q < total) stage(q, Sb);
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (nbuf == 2) { if (q + G < total) stage(q + G, Sb + (long)N * ldp); asm volatile("cp.async.commit_group;" ::: "memory"); }
  for (int k = 0; q < total; q += G, ++k) {
    T* S = Sb + ((nbuf == 2 && (k & 1)) ? (long)N * ldp : 0);
    const int b = q / npan, chain = b / F, f = b - chain * F, v0 = (q - b * npan) * OPS_PW, pw = min(OPS_PW, nvec - v0);
    const FixListDev& FL = mf.fix[lt][f];
    const unsigned char* blob = blob0 + (long)f * blob_stride;
    // diagonal vertices: thread i owns site i (+ blockDim.x, ...): the field value of the NEXT slice of the thread's first site is fetched one slice ahead
    const VDiagDev VD = mf.vd[has_v ? lv : 0][f];
    const T* vtab = reinterpret_cast<const T*>(VD.tab);
    const int8_t* flc = fields + (long)chain * Ltrot * n_opv;
    const double* fcc = md.fields_c ? md.fields_c + (long)chain * Ltrot * n_opv : nullptr;
    int vn0 = -1; int8_t fnext = 0;
    if (has_v) {
      if (tid < N) vn0 = VD.fidx[tid];
      if (vn0 >= 0) fnext = flc[(long)((dir > 0 ? nt_a : nt_b) - 1) * n_opv + vn0];
    }
    if (nbuf == 2) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const FixGroupDev* grp = reinterpret_cast<const FixGroupDev*>(blob + FL.o_grp);
    const FixFamDev* rfam = reinterpret_cast<const FixFamDev*>(blob + FL.o_rfam);
    const T* rmat = reinterpret_cast<const T*>(blob + FL.o_rmat);
    const unsigned* rtab = reinterpret_cast<const unsigned*>(blob + FL.o_ring);
    const unsigned* offs = reinterpret_cast<const unsigned*>(blob + FL.o_offs);
    const int* fam_start = reinterpret_cast<const int*>(blob + FL.o_fs);
    const unsigned char* funi = blob + FL.o_uni;
    const T* fm4 = reinterpret_cast<const T*>(blob + FL.o_m4);
    const bool lane_ok = lane < pw;
    const T* mats = reinterpret_cast<const T*>(FL.mat);
    const int n_grp = FL.n_grp;
    // the scaling is fused into the first (V before T) / last (T before V) group if that group's rings contain every row
    const bool fuse = has_v && n_grp > 0 && grp[uf0 ? 0 : n_grp - 1].kind == 1 && grp[uf0 ? 0 : n_grp - 1].cover;
    for (int sl = 0; sl < ns; ++sl) {
      const int nt = (dir > 0) ? nt_a + sl : nt_b - sl;
      if (has_v) {
        // diagonal vertices of slice nt: row scaling d(P_n) = exp(+-g phi(s_n) E_n) (tabulated per field value; continuous fields on the fly)
        for (int i = tid; i < N; i += nthr) {
          const int n = (i == tid) ? vn0 : VD.fidx[i];
          T v = one_<T>();
          if (n >= 0) {
            const T* vt = vtab + (long)(nt - 1) * VD.tab_nt_stride;          // g_t: one table per time slice
            if (fcc && VD.cont[i]) v = exp_(vt[(long)i * ALF_NVAR] * fcc[(long)(nt - 1) * n_opv + n]);
            else v = vt[(long)i * ALF_NVAR + (int)((i == tid) ? fnext : flc[(long)(nt - 1) * n_opv + n]) + 2];
          }
          dsc[i] = v;
        }
        if (sl + 1 < ns && vn0 >= 0) fnext = flc[(long)(((dir > 0) ? nt + 1 : nt - 1) - 1) * n_opv + vn0];
        __syncthreads();
        if (uf0 && !fuse) {
          if (lane_ok) for (int i = warp; i < N; i += nw) { const int e = ops_sw(i * ldp, lane); S[e] = dsc[i] * S[e]; }
          __syncthreads();
        }
      }
      for (int gi = 0; gi < n_grp; ++gi) {
        const FixGroupDev g = grp[gi];
        if (g.kind == 1) {
          const T* pre = (uf0 && fuse && gi == 0) ? dsc : nullptr;
          const T* post = (uf1 && fuse && gi == n_grp - 1) ? dsc : nullptr;
          const unsigned* rt = rtab + g.ring_off;
          if (g.n == NMAX) { for (int ring = warp; ring < g.nrings; ring += nw) ops_ring_task<T, NMAX, true>(S, rt + ring * OPS_RSTR, ring, g.n, g.nfam, rfam + g.fam0, rmat, pre, post, lane); }
          else { for (int ring = warp; ring < g.nrings; ring += nw) ops_ring_task<T, NMAX, false>(S, rt + ring * OPS_RSTR, ring, g.n, g.nfam, rfam + g.fam0, rmat, pre, post, lane); }
          __syncthreads();
          continue;
        }
        for (int fa = g.fam0; fa < g.fam0 + g.nfam; ++fa) {
          const int o0 = fam_start[fa], o1 = fam_start[fa + 1];
          if (funi[fa]) {
            const T a00 = fm4[4 * fa], a10 = fm4[4 * fa + 1], a01 = fm4[4 * fa + 2], a11 = fm4[4 * fa + 3];
            for (int o = o0 + warp; o < o1; o += 2 * nw) {
              const bool two = o + nw < o1;
              const unsigned d0 = offs[o], d1 = offs[two ? o + nw : o];
              if (lane_ok) {
                const int p0 = ops_sw(d0 & 0xffff, lane), p1 = ops_sw(d0 >> 16, lane), q0 = ops_sw(d1 & 0xffff, lane), q1 = ops_sw(d1 >> 16, lane);
                const T x0 = S[p0], x1 = S[p1], y0 = S[q0], y1 = S[q1];
                S[p0] = a00 * x0 + a01 * x1; S[p1] = a10 * x0 + a11 * x1;
                if (two) { S[q0] = a00 * y0 + a01 * y1; S[q1] = a10 * y0 + a11 * y1; }
              }
            }
          } else {
            for (int o = o0 + warp; o < o1; o += nw) {
              const unsigned d0 = offs[o];
              const T a00 = mats[4 * o], a10 = mats[4 * o + 1], a01 = mats[4 * o + 2], a11 = mats[4 * o + 3];
              if (lane_ok) {
                const int p0 = ops_sw(d0 & 0xffff, lane), p1 = ops_sw(d0 >> 16, lane);
                const T x0 = S[p0], x1 = S[p1];
                S[p0] = a00 * x0 + a01 * x1; S[p1] = a10 * x0 + a11 * x1;
              }
            }
          }
          __syncthreads();
        }
      }
      if (uf1 && !fuse) {
        if (lane_ok) for (int i = warp; i < N; i += nw) { const int e = ops_sw(i * ldp, lane); S[e] = dsc[i] * S[e]; }
        __syncthreads();
      }
    }
    // ---- write back, then refill this buffer with the panel two steps ahead (one step ahead with a single buffer)
    T* Mo = Mout + (long)b * sM;
    if (SIDE == 0) {
      for (int j = warp; j < pw; j += nw) {
        T* col = Mo + (long)(v0 + j) * N + lane; const T* sp = S + (lane * ldp + (j ^ lane));
        for (int i = lane; i < N; i += 32, col += 32, sp += 32 * ldp) *col = *sp;
      }
    } else {
      if (lane < pw) { T* dst = Mo + v0 + lane + (long)warp * N; const long st = (long)nw * N; for (int i = warp; i < N; i += nw, dst += st) *dst = S[i * ldp + (lane ^ (i & 31))]; }
    }
    __syncthreads();
    const int qn = q + nbuf * G;
    if (qn < total) stage(qn, S);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
}
