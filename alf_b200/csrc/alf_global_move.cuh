// Global-in-slice moves on the device: Wrapgr_PlaceGR and Wrapgr_Random_update (Prog/Wrapgr_mod.F90:247-433) with
// Upgrade2 in its "Intermediate" / "Final" modes (Prog/upgrade_mod.F90:203-222) for every chain of the handle.
//
// ham%Global_move_tau is a plugin callback (it proposes Flip_list, Flip_value, T0_Proposal_ratio, S0_ratio); the host
// evaluates it and hands the proposals of all chains to this kernel, which performs everything the reference does with
// them: sort order is expected on input (the C-ABI sorts, Wrapgr_sort :437-480), then per flip PlaceGR to n - 1,
// Op_Wrapup N_type 1, Upgrade2, Op_Wrapup N_type 2, the Metropolis test on the accumulated ratio with the chain's own
// random stream, and on rejection of a multi-field move the rollback of G (GR_st) and of the fields.
//
// One CTA per chain.  Moves are rare and touch few fields, so G is updated in place in global memory (rank-1 updates per
// non-zero eigen-direction: the same Sherman-Morrison sequence as upgrade_mod.F90:225-285); no delayed factors here.
#pragma once
#include "alf_update.cuh"

template <typename T>
__device__ __forceinline__ void gm_similarity(T* __restrict__ Gf, int N, int ldg, const VopDev<T>* op, int s, int mode, T* ones, T* ones_r, T* AL_s, T* AR_s) {
  // mode 0: Op_Wrapup N_type 1 (AL = diag(e) U^H, AR = U diag(1/e)); 1: Op_Wrapup N_type 2 (AL = U, AR = U^H)
  // mode 2: Op_Wrapdo N_type 2 (AL = U^H, AR = U);                    3: Op_Wrapdo N_type 1 (AL = U diag(1/e), AR = diag(e) U^H)
  // mode 4 / 5: both steps of a whole vertex at once, AL = U diag(e^{+-1}) U^H, AR = U diag(e^{-+1}) U^H
  const int k = op->k;
  if (op->diag && (mode == 1 || mode == 2)) return;                 // U = 1: nothing to rotate
  if (threadIdx.x < ALF_KMAX * ALF_KMAX) {
    const int a = threadIdx.x % ALF_KMAX, b = threadIdx.x / ALF_KMAX;
    T al = zero_<T>(), ar = zero_<T>();
    if (a < k && b < k) {
      const T ea = op->E_exp[a][s + 2], eb = op->E_exp[b][s + 2];
      const T uab = op->U[a + b * ALF_KMAX], uba_c = conj_(op->U[b + a * ALF_KMAX]);
      if (mode == 0) { al = ea * uba_c; ar = uab * (one_<T>() / eb); }
      else if (mode == 1) { al = uab; ar = uba_c; }
      else if (mode == 2) { al = uba_c; ar = uab; }
      else if (mode == 3) { al = uab * (one_<T>() / eb); ar = ea * uba_c; }
      else {      // 4: e^{V} G e^{-V} = modes 0 then 1 in one step; 5: e^{-V} G e^{V} = modes 2 then 3 (Wrapgr_PlaceGR moves G by whole vertices)
        for (int c = 0; c < k; ++c) {
          const T ec = op->E_exp[c][s + 2], eci = one_<T>() / ec; const T uu = op->U[a + c * ALF_KMAX] * conj_(op->U[b + c * ALF_KMAX]);
          al = al + uu * (mode == 4 ? ec : eci); ar = ar + uu * (mode == 4 ? eci : ec);
        }
      }
    }
    AL_s[a + b * ALF_KMAX] = al; AR_s[a + b * ALF_KMAX] = ar;
  }
  __syncthreads();
  T AL[ALF_KMAX * ALF_KMAX], AR[ALF_KMAX * ALF_KMAX];
#pragma unroll
  for (int e = 0; e < ALF_KMAX * ALF_KMAX; ++e) { AL[e] = AL_s[e]; AR[e] = AR_s[e]; }
  similarity_immediate<T>(Gf, N, ldg, nullptr, nullptr, 0, 0, ones, ones_r, op->P, k, AL, AR);      // dl, dr: two arrays of ones (no pending diagonal factors here)
}

// ham%Global_move_tau for Ising star moves as tables (Hamiltonian_Z2_Matter_smod.F90:535-643), evaluated on the device: a site
// I = nranf(n_sites) is drawn from the chain's stream, the (ascending) fields move_fields[move_start[I] ..) are flipped, S0_Matter is the
// product of the site's coupling terms (same form as S0TabDev: products of current +-1 fields at time offsets -> flip-ratio table),
// T0_Proposal = 1 - 1/(1 + S0_Matter) is tested against one ranf() and T0_Proposal_ratio = 1/S0_Matter or 0 (:633-641).
struct GmtDev { int on, n_sites; const int* move_start; const int* move_fields; S0TabDev terms; };

// product over the owner's terms, evaluated by the whole CTA (one entry per thread and pass, parity through __syncthreads_count):
// every thread returns the same value.  Must be called by all threads of the block.
__device__ __forceinline__ double ising_terms_block(const S0TabDev& t, const int8_t* __restrict__ fchain, int owner, int nt, int Ltrot, int n_opv) {
  double S = 1.0;
  for (int q = t.op_start[owner]; q < t.op_start[owner + 1]; ++q) {
    const int e0 = t.term_start[q], e1 = t.term_start[q + 1];
    int neg = 0, out = 0;
    for (int e = e0 + (int)threadIdx.x; e < e1; e += (int)blockDim.x) {
      int nt1 = nt + t.e_dt[e];
      if (nt1 > Ltrot || nt1 < 1) { if (t.open_bc) { out = 1; continue; } nt1 = (nt1 > Ltrot) ? nt1 - Ltrot : nt1 + Ltrot; }
      neg ^= (fchain[(long)(nt1 - 1) * n_opv + t.e_op[e]] < 0) ? 1 : 0;
    }
    const int nneg = __syncthreads_count(neg), nout = __syncthreads_count(out);
    if (!nout) S *= t.w[2 * q + ((nneg & 1) ? 0 : 1)];
  }
  return S;
}

// One step of Wrapgr_PlaceGR: G <- A G A^-1 with A = e^{+-V_n(s)} restricted to the vertex' support P (k x k, column-major, ld ALF_KMAX; both
// matrices tabulated on the host per vertex and field value).  The rows P of the columns outside P, the columns P of the rows outside P and
// the P x P block are disjoint pieces handled by different threads, so a step costs ONE block barrier.
// what one step of Wrapgr_PlaceGR needs, fetched one step ahead of its use (the chain field value -> table row -> matrix is two dependent
// memory latencies, longer than the step itself)
template <typename T>
__device__ __forceinline__ T* Gc_end(T* AR_s, int stage_g, int F, int N) { T* g = AR_s + ALF_KMAX * ALF_KMAX; return stage_g ? g + (long)F * (N | 1) * N : g; }
template <typename T>
struct PlaceOp { int p0, p1, k; T a00, a10, a01, a11, i00, i10, i01, i11; const T* A; const T* Ai; const int* P; };     // scalars only: stays in registers
template <typename T>
__device__ __forceinline__ void gm_place_fetch(PlaceOp<T>& o, const int* __restrict__ place_pk, const T* __restrict__ place_tab, long nf, int s, bool fwd) {
  o.P = place_pk + nf * 8;
  const int4 pk = *reinterpret_cast<const int4*>(o.P); o.k = o.P[4]; o.p0 = pk.x; o.p1 = (o.k > 1) ? pk.y : -1;
  const T* tb = place_tab + (nf * ALF_NVAR + (s + 2)) * 2 * ALF_KMAX * ALF_KMAX;
  o.A = fwd ? tb : tb + ALF_KMAX * ALF_KMAX; o.Ai = fwd ? tb + ALF_KMAX * ALF_KMAX : tb;
  o.a00 = o.A[0]; o.a10 = o.A[1]; o.a01 = o.A[ALF_KMAX]; o.a11 = o.A[1 + ALF_KMAX];          // k <= 2: all that is needed (zero padded for k = 1)
  o.i00 = o.Ai[0]; o.i10 = o.Ai[1]; o.i01 = o.Ai[ALF_KMAX]; o.i11 = o.Ai[1 + ALF_KMAX];
}
// One step of Wrapgr_PlaceGR: G <- A G A^-1 with A = e^{+-V_n(s)} restricted to the vertex' support P (k x k, column-major, ld ALF_KMAX; both
// matrices tabulated on the host per vertex and field value).  The rows P of the columns outside P, the columns P of the rows outside P and
// the P x P block are disjoint pieces handled by different threads, so a step costs ONE block barrier.
template <typename T>
__device__ __forceinline__ void gm_place_step(T* __restrict__ Gf, int N, int ldg, const PlaceOp<T>& o) {
  const int k = o.k;
  if (k <= 2) {
    const int p0 = o.p0, p1 = o.p1;
    for (int t = threadIdx.x; t < 2 * N + 1; t += blockDim.x) {
      if (t < N) {                 // column t: rows P
        const int j = t; if (j == p0 || j == p1) continue;
        T* g0 = Gf + p0 + (long)j * ldg;
        if (k == 2) { T* g1 = Gf + p1 + (long)j * ldg; const T v0 = *g0, v1 = *g1; *g0 = o.a00 * v0 + o.a01 * v1; *g1 = o.a10 * v0 + o.a11 * v1; }
        else *g0 = o.a00 * *g0;
      } else if (t < 2 * N) {      // row t - N: columns P
        const int i = t - N; if (i == p0 || i == p1) continue;
        T* g0 = Gf + i + (long)p0 * ldg;
        if (k == 2) { T* g1 = Gf + i + (long)p1 * ldg; const T v0 = *g0, v1 = *g1; *g0 = v0 * o.i00 + v1 * o.i10; *g1 = v0 * o.i01 + v1 * o.i11; }
        else *g0 = *g0 * o.i00;
      } else {                     // the P x P block: A G(P,P) A^-1
        if (k == 2) {
          const T b00 = Gf[p0 + (long)p0 * ldg], b10 = Gf[p1 + (long)p0 * ldg], b01 = Gf[p0 + (long)p1 * ldg], b11 = Gf[p1 + (long)p1 * ldg];
          const T c00 = o.a00 * b00 + o.a01 * b10, c10 = o.a10 * b00 + o.a11 * b10, c01 = o.a00 * b01 + o.a01 * b11, c11 = o.a10 * b01 + o.a11 * b11;
          Gf[p0 + (long)p0 * ldg] = c00 * o.i00 + c01 * o.i10; Gf[p1 + (long)p0 * ldg] = c10 * o.i00 + c11 * o.i10;
          Gf[p0 + (long)p1 * ldg] = c00 * o.i01 + c01 * o.i11; Gf[p1 + (long)p1 * ldg] = c10 * o.i01 + c11 * o.i11;
        } else Gf[p0 + (long)p0 * ldg] = (o.a00 * Gf[p0 + (long)p0 * ldg]) * o.i00;
      }
    }
    __syncthreads();
    return;
  }
  const int* P = o.P; const T* A = o.A; const T* Ai = o.Ai;      // larger vertices: operands straight from the tables
  for (int t = threadIdx.x; t < 2 * N + 1; t += blockDim.x) {
    if (t < N) {
      const int j = t; bool in = false;
      for (int a = 0; a < k; ++a) in = in || (P[a] == j);
      if (in) continue;
      T v[ALF_KMAX];
      for (int b = 0; b < k; ++b) v[b] = Gf[P[b] + (long)j * ldg];
      for (int a = 0; a < k; ++a) { T x = zero_<T>(); for (int b = 0; b < k; ++b) fma_(x, A[a + b * ALF_KMAX], v[b]); Gf[P[a] + (long)j * ldg] = x; }
    } else if (t < 2 * N) {
      const int i = t - N; bool in = false;
      for (int a = 0; a < k; ++a) in = in || (P[a] == i);
      if (in) continue;
      T v[ALF_KMAX];
      for (int b = 0; b < k; ++b) v[b] = Gf[i + (long)P[b] * ldg];
      for (int a = 0; a < k; ++a) { T x = zero_<T>(); for (int b = 0; b < k; ++b) fma_(x, v[b], Ai[b + a * ALF_KMAX]); Gf[i + (long)P[a] * ldg] = x; }
    } else {
      T B[ALF_KMAX][ALF_KMAX], C[ALF_KMAX][ALF_KMAX];
      for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) B[a][b] = Gf[P[a] + (long)P[b] * ldg];
      for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) { T x = zero_<T>(); for (int c = 0; c < k; ++c) fma_(x, A[a + c * ALF_KMAX], B[c][b]); C[a][b] = x; }
      for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) { T x = zero_<T>(); for (int c = 0; c < k; ++c) fma_(x, C[a][c], Ai[c + b * ALF_KMAX]); Gf[P[a] + (long)P[b] * ldg] = x; }
    }
  }
  __syncthreads();
}

// proposals: [chain][move]: length, t0 ratio, s0 ratio; [chain][move][maxlen]: 0-based op index (ascending), new field value
template <typename T>
__global__ void __launch_bounds__(512, 1) k_random_update(T* __restrict__ G, T* __restrict__ Gst, int N, int F, int n_sun, int M, const VopDev<T>* __restrict__ vops,
                                                          FieldTabDev ft, int8_t* __restrict__ fields, int Ltrot, int nt, uint64_t* __restrict__ rng,
                                                          cplx* __restrict__ phase, unsigned long long* __restrict__ counters, int* __restrict__ mpos,
                                                          int n_moves, int maxlen, const int* __restrict__ flip_len, const int* __restrict__ flip_list,
                                                          const int8_t* __restrict__ flip_val, const double* __restrict__ t0r, const double* __restrict__ s0r,
                                                          uint8_t* __restrict__ acc_out, int place_to /* >= 0: final PlaceGR target, -1: none */, GmtDev gmt, int stage_g, const T* __restrict__ place_tab /* [n][f][var][2][KMAX*KMAX]: e^{V}, e^{-V} */,
                                                          const int* __restrict__ place_pk /* [n][f][8]: P[0..3], k */, int stage_f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ones = reinterpret_cast<T*>(smem_raw);        // N
  T* ones_r = ones + N;                            // N
  T* col = ones_r + N;                             // N
  T* row = col + N;                                // N
  T* AL_s = row + N; T* AR_s = AL_s + ALF_KMAX * ALF_KMAX;
  __shared__ int s_acc; __shared__ T s_xf; __shared__ double s_prev[2];
  const int chain = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  T* Gglob = G + (long)chain * F * N * N; T* Gs = Gst + (long)chain * F * N * N;
  // stage_g: G of the chain in shared memory (odd leading dimension) while the moves walk it through the operator positions of the slice
  const int ldg = stage_g ? (N | 1) : N; const long sG = (long)ldg * N;
  T* Gc = stage_g ? (AR_s + ALF_KMAX * ALF_KMAX) : Gglob;
  if (stage_g) for (long e = tid; e < (long)F * N * N; e += nthr) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gc[fq * sG + (q % N) + (q / N) * ldg] = Gglob[e]; }
  int8_t* fld_g = fields + ((long)chain * Ltrot + (nt - 1)) * M;
  // stage_f: the slice's fields in shared memory too (read at every PlaceGR step); writes go to both copies
  int8_t* fld = stage_f ? reinterpret_cast<int8_t*>(Gc_end(AR_s, stage_g, F, N)) : fld_g;
  if (stage_f) for (int i = tid; i < M; i += nthr) fld[i] = fld_g[i];
  for (int i = tid; i < N; i += nthr) { ones[i] = one_<T>(); ones_r[i] = one_<T>(); }
  __syncthreads();
  int m = mpos[chain];
  Xoshiro r; cplx ph = phase[chain];
  r.s0 = rng[chain * 4 + 0]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3];    // every thread keeps a copy; thread 0's is stored
  unsigned long long n_acc = 0, n_prop = 0;

  auto place = [&](int m1) {                       // Wrapgr_PlaceGR (:247-312), whole vertices per step, operands fetched one step ahead
    const bool fwd = m1 > m; const int steps = (fwd ? m1 - m : m - m1) * F;      // one step = (vertex, flavor)
    if (steps > 0) {
      PlaceOp<T> cur, nxt;
      const int n_first = fwd ? m : m - 1;
      gm_place_fetch<T>(cur, place_pk, place_tab, (long)n_first * F, fld[n_first], fwd);
      for (int st_ = 0; st_ < steps; ++st_) {
        const int f = st_ % F;
        if (st_ + 1 < steps) { const int q = (st_ + 1) / F, n2 = fwd ? m + q : m - 1 - q, f2 = (st_ + 1) % F;
          gm_place_fetch<T>(nxt, place_pk, place_tab, (long)n2 * F + f2, fld[n2], fwd); }
        gm_place_step<T>(Gc + f * sG, N, ldg, cur);
        cur = nxt;
      }
    }
    m = m1;
  };

  for (int mv = 0; mv < n_moves; ++mv) {
    const long pi = (long)chain * n_moves + mv;
    int len; double T0, S0; const int* fl; const int8_t* fv; int8_t fv_dev[ALF_GM_MAXLEN];
    if (gmt.on) {                                   // the proposal comes from the tables (every thread draws the same numbers)
      const int I = r.nranf(gmt.n_sites) - 1;
      fl = gmt.move_fields + gmt.move_start[I]; len = gmt.move_start[I + 1] - gmt.move_start[I];
      for (int c = 0; c < len && c < ALF_GM_MAXLEN; ++c) fv_dev[c] = (int8_t)(-fld[fl[c]]);      // nsigma%flip of an Ising field (Fields_mod.F90:190)
      fv = fv_dev;
      S0 = ising_terms_block(gmt.terms, fields + (long)chain * Ltrot * M, I, nt, Ltrot, M);
      const double T0_Proposal = 1.0 - 1.0 / (1.0 + S0);
      T0 = (T0_Proposal > r.ranf()) ? 1.0 / S0 : 0.0;
    } else { len = flip_len[pi]; T0 = t0r[pi]; S0 = s0r[pi]; fl = flip_list + pi * maxlen; fv = flip_val + pi * maxlen; }
    if (!(T0 > 10e-8) || len <= 0) { if (acc_out && tid == 0) acc_out[pi] = 2; continue; }
    cplx prev = cplx(1.0, 0.0);
    int acc = 0;
    int8_t old_vals[ALF_GM_MAXLEN];
    for (int c = 0; c < len; ++c) old_vals[c] = fld[fl[c]];
    for (int c = 0; c < len; ++c) {
      const int n = fl[c];
      place(n);                                    // reference: PlaceGR(m, n - 1) with 1-based n
      if (c == 0 && len > 1) { for (long e = tid; e < (long)F * N * N; e += nthr) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gs[e] = Gc[fq * sG + (q % N) + (q / N) * ldg]; } }
      const int s_old = fld[n], s_new = fv[c];
      const VopDev<T>* op0 = vops + (long)n * F;
      for (int f = 0; f < F; ++f) gm_similarity<T>(Gc + f * sG, N, ldg, op0 + f, s_old, 0, ones, ones_r, AL_s, AR_s);
      // ---- Upgrade2: ratio (every thread computes it redundantly from global G; k <= 4)
      cplx ratiotot = cplx(1.0, 0.0);
      for (int f = 0; f < F; ++f) {
        const VopDev<T>* op = op0 + f; const int nz = op->nnz;
        T Mat[ALF_KMAX][ALF_KMAX];
        for (int q = 0; q < ALF_KMAX; ++q) for (int w = 0; w < ALF_KMAX; ++w) Mat[q][w] = zero_<T>();
        for (int w = 0; w < nz; ++w) {
          const T d = op->delta[w][s_old + 2][s_new + 2];
          for (int q = 0; q < nz; ++q) Mat[q][w] = -(d * Gc[f * sG + op->P[q] + (long)op->P[w] * ldg]);
          Mat[w][w] = Mat[w][w] + (d + one_<T>());
        }
        T D;
        if (nz == 0) D = one_<T>();
        else if (nz == 1) D = Mat[0][0];
        else if (nz == 2) {
          T s1 = Mat[0][0] * Mat[1][1], s2 = Mat[1][0] * Mat[0][1];
          if (abs_(s1) > abs_(s2)) D = s1 * (one_<T>() - s2 / s1); else D = s2 * (s1 / s2 - one_<T>());
        } else {
          D = one_<T>();
          for (int cc = 0; cc < nz; ++cc) {
            int pv = cc; double best = abs_(Mat[cc][cc]);
            for (int q = cc + 1; q < nz; ++q) if (abs_(Mat[q][cc]) > best) { best = abs_(Mat[q][cc]); pv = q; }
            if (pv != cc) { for (int q = 0; q < nz; ++q) { T t = Mat[cc][q]; Mat[cc][q] = Mat[pv][q]; Mat[pv][q] = t; } D = -D; }
            D = D * Mat[cc][cc];
            for (int q = cc + 1; q < nz; ++q) { T l = Mat[q][cc] / Mat[cc][cc]; for (int w = cc; w < nz; ++w) Mat[q][w] = Mat[q][w] - l * Mat[cc][w]; }
          }
        }
        const T rf = D * op->expalpha[s_old + 2][s_new + 2];
        ratiotot = ratiotot * cplx(real_(rf), imag_(rf));
      }
      cplx rt = ratiotot;
      for (int q = 1; q < n_sun; ++q) rt = rt * ratiotot;
      const int type = op0->type;
      rt = rt * (ft.gama[type][s_new + 2] / ft.gama[type][s_old + 2]);
      const bool fin = (c == len - 1);
      double weight;
      if (fin) { rt = rt * prev; const cplx pr = ph * rt; weight = S0 * T0 * fabs(pr.x / ph.x); }
      else { weight = 1.5; prev = prev * rt; }
      const double u = r.ranf();
      const int toggle = (weight > u) ? 1 : 0;
      if (fin) { n_prop++; acc = toggle; if (toggle) { n_acc++; const double ar = abs_(rt); ph = ph * cplx(rt.x / ar, rt.y / ar); } }
      __syncthreads();
      if (toggle) {
        // rank-1 updates over the non-zero eigen-directions (upgrade_mod.F90:225-285 in sequential Sherman-Morrison form)
        const int nzmax = op0->nnz;
        for (int a = 0; a < nzmax; ++a) for (int f = 0; f < F; ++f) {
          const VopDev<T>* op = op0 + f;
          if (a >= op->nnz) continue;
          T* Gf = Gc + f * sG; const int p = op->P[a];
          for (int i = tid; i < N; i += nthr) { col[i] = Gf[i + (long)p * ldg]; row[i] = Gf[p + (long)i * ldg]; }
          __syncthreads();
          const T d = op->delta[a][s_old + 2][s_new + 2];
          const T xf = d / (one_<T>() + (one_<T>() - col[p]) * d);
          for (long e = tid; e < (long)N * N; e += nthr) {
            const int i = (int)(e % N), j = (int)(e / N);
            const T y = ((j == p) ? one_<T>() : zero_<T>()) - row[j];
            Gf[i + (long)j * ldg] = Gf[i + (long)j * ldg] - (xf * col[i]) * y;
          }
          __syncthreads();
        }
        if (tid == 0) { fld[n] = (int8_t)s_new; fld_g[n] = (int8_t)s_new; }
        __syncthreads();
      }
      for (int f = 0; f < F; ++f) gm_similarity<T>(Gc + f * sG, N, ldg, op0 + f, s_old, 1, ones, ones_r, AL_s, AR_s);
      m = n + 1;                                   // reference: m = n (1-based)
    }
    if (!acc && len > 1) {                         // rollback (:421-427)
      for (long e = tid; e < (long)F * N * N; e += nthr) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gc[fq * sG + (q % N) + (q / N) * ldg] = Gs[e]; }
      if (tid == 0) for (int c = 0; c + 1 < len; ++c) { fld[fl[c]] = old_vals[c]; fld_g[fl[c]] = old_vals[c]; }
      m = fl[0];
      __syncthreads();
    }
    if (acc_out && tid == 0) acc_out[pi] = (uint8_t)acc;
  }
  if (place_to >= 0) place(place_to);
  if (stage_g) { __syncthreads(); for (long e = tid; e < (long)F * N * N; e += nthr) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gglob[e] = Gc[fq * sG + (q % N) + (q / N) * ldg]; } }
  if (tid == 0) {
    mpos[chain] = m; phase[chain] = ph;
    rng[chain * 4 + 0] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
    counters[chain * 4 + 0] += n_prop; counters[chain * 4 + 1] += n_acc; counters[chain * 4 + 2] += n_prop; counters[chain * 4 + 3] += n_acc;
  }
}
