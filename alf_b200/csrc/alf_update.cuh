// One time slice of sequential single-field updates for every Markov chain: the body of the n-loop of
// WRAPGRUP / WRAPGRDO (Prog/Wrapgr_mod.F90:115-146, 191-237) = Op_Wrapup/Op_Wrapdo (Prog/Operator_mod.F90:743-951),
// Fields%flip (Prog/Fields_mod.F90:173-217), Upgrade2 (Prog/upgrade_mod.F90:105-302) and the acceptance
// counters (Prog/control_mod.F90:164-176).
//
// B200 mapping.  The reference applies every accepted flip as a rank-1 ZGERU on the N x N Green function
// (32 N^2 bytes of traffic per accept: HBM bound).  Here one CTA owns one chain (all flavors) and keeps the
// accepted updates as DELAYED factors in shared memory,
//        G_cur = DL * G0 * DR  -  X * Y^T ,
// where G0 stays in global memory (L2 resident), X, Y hold up to KD accepted rank-1 updates, and DL/DR are the pending
// diagonal similarity transformations e^{V_n} ... e^{-V_n} of diagonal vertices.  A field visit then only needs
// G_cur(P,P) (k^2 short dot products); an accepted flip needs one column and one row of G_cur (O(N*nd)); the O(N^2)
// work happens once per KD accepts as a register-tiled FP64 rank-KD update ("flush").  In exact arithmetic the Markov
// chain is identical to the reference's.  Non-diagonal vertices (k >= 2, e.g. Kondo J_K) apply their similarity
// transformations to the k rows / k columns of G0 immediately.
//
// Random numbers: one xoshiro256** stream per chain, consumed in the reference's order
// [flip draw(s)] -> [proposal draw] -> [acceptance draw]   (Wrapgr_mod.F90:125,133; upgrade_mod.F90:222).
#pragma once
#include "alf_types.cuh"
#include "alf_ops.cuh"

template <typename T>
struct VopDev {                       // one interaction vertex Op_V(n, nf) as the update kernel needs it
  int k, nnz, diag, type;
  int P[ALF_KMAX];
  T E_exp[ALF_KMAX][ALF_NVAR];        // exp(g phi(s) E_a), index s+2      (Operator_mod.F90:413-470)
  T delta[ALF_KMAX][ALF_NVAR][ALF_NVAR];   // exp(g (phi(s')-phi(s)) E_a) - 1, [a][s+2][s'+2]   (upgrade_mod.F90:168-175)
  T expalpha[ALF_NVAR][ALF_NVAR];     // exp(g (phi(s')-phi(s)) alpha)    (upgrade_mod.F90:193)
  T U[ALF_KMAX * ALF_KMAX];           // eigenvectors (non-diagonal vertices), column-major
  T gE[ALF_KMAX];                     // g E_a: continuous fields (type 3) evaluate exp(g phi E_a) on the fly (Operator_mod.F90:585-600)
  T galpha;                           // g alpha
};

struct FieldTabDev {                  // Prog/Fields_mod.F90:258-303
  double gama[3][ALF_NVAR];           // [type][s+2]
  int flip[ALF_NVAR][4];              // Flip_st(s, 1..3)
};

struct Xoshiro {
  uint64_t s0, s1, s2, s3;
  __device__ __forceinline__ static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  __device__ __forceinline__ double ranf() {
    const uint64_t result = rotl(s1 * 5, 7) * 9;
    const uint64_t t = s1 << 17;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 45);
    return (double)(result >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ __forceinline__ int nranf(int N) {   // random_wrap_mod.F90:159-168
    int r = (int)llrint(floor(ranf() * (double)N + 0.5 + 0.5)); // nint(x) = floor(x+0.5) for x >= 0
    if (r < 1) r = 1; if (r > N) r = N; return r;
  }
};

struct UpdCtl {                       // per-op broadcast block (double buffered by op parity)
  int accept; int s_new; double phi_new;
};

// ---- G0 <- DL G0 DR - X Y^T for one flavor, then DL = DR = 1.  128 x 128 output tiles, 8 x 4 register micro-tiles.
template <typename T>
__device__ __forceinline__ void flush_flavor(T* __restrict__ G0, int N, int ldg, const T* __restrict__ X, const T* __restrict__ Y, int ldx, int nd,
                                             T* __restrict__ dl, T* __restrict__ dr) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;      // 16 x 32 threads
  for (int j0 = 0; j0 < N; j0 += 128)
    for (int i0 = 0; i0 < N; i0 += 128) {
      T acc[8][4];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = zero_<T>();
      const int ib = i0 + tx * 8, jb = j0 + ty * 4;
      for (int k = 0; k < nd; ++k) {
        T xr[8], yr[4];
        const T* xk = X + (long)k * ldx; const T* yk = Y + (long)k * ldx;
#pragma unroll
        for (int r = 0; r < 8; ++r) xr[r] = (ib + r < N) ? xk[ib + r] : zero_<T>();
#pragma unroll
        for (int c = 0; c < 4; ++c) yr[c] = (jb + c < N) ? yk[jb + c] : zero_<T>();
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) fma_(acc[r][c], xr[r], yr[c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = jb + c;
        if (j >= N) continue;
        const T drj = dr[j];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = ib + r;
          if (i < N) { T g = G0[i + (long)j * ldg]; G0[i + (long)j * ldg] = (dl[i] * g) * drj - acc[r][c]; }
        }
      }
    }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) { dl[i] = one_<T>(); dr[i] = one_<T>(); }
}

// Apply  G_cur <- AL * G_cur * AR  restricted to rows / columns P (k x k matrices) IMMEDIATELY to G0 (global) and to the
// delayed factors: rows P of X by AL, rows P of Y by AR^T.  Pending diagonal factors on P are folded in first.
template <typename T>
__device__ __forceinline__ void similarity_immediate(T* __restrict__ G0, int N, int ldg, T* __restrict__ X, T* __restrict__ Y, int ldx, int nd,
                                                     T* __restrict__ dl, T* __restrict__ dr, const int* P, int k, const T* AL, const T* AR) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  // rows: G0(P,:) <- AL * diag(dl_P) * G0(P,:)
  for (int j = tid; j < N; j += nthr) {
    T v[ALF_KMAX];
    for (int a = 0; a < k; ++a) v[a] = dl[P[a]] * G0[P[a] + (long)j * ldg];
    for (int a = 0; a < k; ++a) { T s = zero_<T>(); for (int b = 0; b < k; ++b) fma_(s, AL[a + b * ALF_KMAX], v[b]); G0[P[a] + (long)j * ldg] = s; }
  }
  for (int kk = tid; kk < nd; kk += nthr) {
    T v[ALF_KMAX], w[ALF_KMAX];
    for (int a = 0; a < k; ++a) { v[a] = X[(long)kk * ldx + P[a]]; w[a] = Y[(long)kk * ldx + P[a]]; }
    for (int a = 0; a < k; ++a) {
      T s = zero_<T>(), t = zero_<T>();
      for (int b = 0; b < k; ++b) { fma_(s, AL[a + b * ALF_KMAX], v[b]); fma_(t, AR[b + a * ALF_KMAX], w[b]); }
      X[(long)kk * ldx + P[a]] = s; Y[(long)kk * ldx + P[a]] = t;
    }
  }
  __syncthreads();
  for (int a = tid; a < k; a += nthr) dl[P[a]] = one_<T>();
  // columns: G0(:,P) <- G0(:,P) * diag(dr_P) * AR
  for (int i = tid; i < N; i += nthr) {
    T v[ALF_KMAX];
    for (int a = 0; a < k; ++a) v[a] = G0[i + (long)P[a] * ldg] * dr[P[a]];
    for (int a = 0; a < k; ++a) { T s = zero_<T>(); for (int b = 0; b < k; ++b) fma_(s, v[b], AR[b + a * ALF_KMAX]); G0[i + (long)P[a] * ldg] = s; }
  }
  __syncthreads();
  for (int a = tid; a < k; a += nthr) dr[P[a]] = one_<T>();
  __syncthreads();
}

// ham%S0(n, nt, Hs_new) of Ising actions as tables (Hamiltonian_Z2_Matter_smod.F90:439-512): term t of field n lists the (field, time offset)
// pairs whose current product of +-1 fields indexes its flip ratio w[t][prod < 0 ? 0 : 1]; S0 is the product over the field's terms.  Time
// offsets wrap periodically, or (open_bc, projective algorithm :465-472) terms that leave 1..Ltrot are dropped.  on == 0: S0 = 1 (S0_base).
struct S0TabDev { int on, open_bc; const int* op_start; const int* term_start; const int* e_op; const int* e_dt; const double* w; };
__device__ __forceinline__ double s0_eval(const S0TabDev& t, const int8_t* __restrict__ fchain, int n, int nt, int Ltrot, int n_opv) {
  if (!t.on) return 1.0;
  double S = 1.0;
  for (int q = t.op_start[n]; q < t.op_start[n + 1]; ++q) {
    int prod = 1; bool skip = false;
    for (int e = t.term_start[q]; e < t.term_start[q + 1]; ++e) {
      int nt1 = nt + t.e_dt[e];
      if (nt1 > Ltrot || nt1 < 1) { if (t.open_bc) { skip = true; break; } nt1 = (nt1 > Ltrot) ? nt1 - Ltrot : nt1 + Ltrot; }
      prod *= (fchain[(long)(nt1 - 1) * n_opv + t.e_op[e]] < 0) ? -1 : 1;
    }
    if (!skip) S *= t.w[2 * q + (prod > 0 ? 1 : 0)];
  }
  return S;
}

// the same product evaluated by one warp: lane l takes entry l (mod 32) of the field's flattened entry list, the parity of every term comes
// from a ballot; all lanes return the value.  Terms of one field hold at most a few dozen entries.
__device__ __forceinline__ double s0_eval_warp(const S0TabDev& t, const int8_t* __restrict__ fchain, int n, int nt, int Ltrot, int n_opv) {
  if (!t.on) return 1.0;
  const int lane = threadIdx.x & 31;
  const int q0 = t.op_start[n], q1 = t.op_start[n + 1];
  double S = 1.0;
  for (int q = q0; q < q1; ++q) {
    const int e0 = t.term_start[q], e1 = t.term_start[q + 1];
    int neg = 0, out = 0;
    for (int e = e0 + lane; e < e1; e += 32) {
      int nt1 = nt + t.e_dt[e];
      if (nt1 > Ltrot || nt1 < 1) { if (t.open_bc) { out = 1; continue; } nt1 = (nt1 > Ltrot) ? nt1 - Ltrot : nt1 + Ltrot; }
      neg ^= (fchain[(long)(nt1 - 1) * n_opv + t.e_op[e]] < 0) ? 1 : 0;
    }
    const unsigned mneg = __ballot_sync(0xffffffffu, neg), mout = __ballot_sync(0xffffffffu, out);
    if (!mout) S *= t.w[2 * q + ((__popc(mneg) & 1) ? 0 : 1)];
  }
  return S;
}

// dynamic smem: X[F][KD][ldx], Y[F][KD][ldx], dl[F][N], dr[F][N], gdiag[F][N] (all T)
template <typename T, int UP>
__global__ void __launch_bounds__(512, 1) k_wrapgr(T* __restrict__ G, int N, int F, int n_sun, int n0, int cnt, int n_opv, int log_off,
                                                   const VopDev<T>* __restrict__ vops,
                                                   FieldTabDev ft, int8_t* __restrict__ fields, int Ltrot, int nt, uint64_t* __restrict__ rng,
                                                   cplx* __restrict__ phase, unsigned long long* __restrict__ counters, int KD,
                                                   uint8_t* __restrict__ acclog, int propose_s0, S0TabDev s0t, int stage_g,
                                                   double* __restrict__ fields_c, int s0_gaussian, double amplitude) {   // visits the vertices n0 .. n0 + cnt - 1 of the slice
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ UpdCtl ctl[2];
  __shared__ T gpp_s[ALF_FMAX][ALF_KMAX][ALF_KMAX];
  __shared__ T xfac_s[ALF_FMAX];
  __shared__ T bc_s[2];
  __shared__ double s0_s;
  const int chain = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int ldx = N + 2;
  T* Xs = reinterpret_cast<T*>(smem_raw);
  T* Ys = Xs + (long)F * KD * ldx;
  T* dl = Ys + (long)F * KD * ldx;
  T* dr = dl + (long)F * N;
  T* gdiag = dr + (long)F * N;
  T* Gglob = G + (long)chain * F * N * N;
  // stage_g: the chain's G lives in shared memory for the whole slice (small lattices: every non-diagonal vertex applies its similarity
  // transformation to G immediately, which is L2 latency per visit otherwise); odd leading dimension against bank conflicts of row accesses
  const int ldg = stage_g ? (N | 1) : N; const long sG = (long)ldg * N;
  T* Gc = stage_g ? (gdiag + (long)F * N) : Gglob;
  if (stage_g) { for (long e = threadIdx.x; e < (long)F * N * N; e += blockDim.x) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gc[fq * sG + (q % N) + (q / N) * ldg] = Gglob[e]; } __syncthreads(); }
  int8_t* fld = fields + ((long)chain * Ltrot + (nt - 1)) * n_opv;
  double* fc = fields_c ? fields_c + ((long)chain * Ltrot + (nt - 1)) * n_opv : nullptr;      // continuous fields (type 3) of this slice
  for (int e = tid; e < F * N; e += nthr) { dl[e] = one_<T>(); dr[e] = one_<T>(); int f = e / N, i = e % N; gdiag[e] = Gc[f * sG + i + (long)i * ldg]; }
  Xoshiro r; cplx ph; unsigned long long n_acc = 0, n_prop = 0, n_flush = 1;
  if (tid == 0) { r.s0 = rng[chain * 4 + 0]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3]; ph = phase[chain]; }
  int nd = 0;
  __syncthreads();

  for (int step = 0; step < cnt; ++step) {
    const int n = UP ? n0 + step : (n0 + cnt - 1 - step);
    const VopDev<T>* op0 = vops + (long)n * F;
    const int k = op0->k, isdiag = op0->diag, type = op0->type;
    const int s_old = (int)fld[n];
    const bool cont = (type == 3) && fc; const double phi_old = cont ? fc[n] : 0.0;
    UpdCtl* cb = &ctl[step & 1];
    // ham%S0 of this visit (fields only; they were last changed before the previous visit's closing barrier): warp 1 evaluates it while the
    // others run the similarity transformation; lane 0 of warp 0 reads it after the barrier in front of the decision
    // (diagonal vertices have no block barrier before the decision: warp 0 evaluates it itself)
    double s0_local = 1.0; const int s0_warp = (isdiag || nthr <= 32) ? 0 : 1;
    if (s0t.on && warp == s0_warp) { s0_local = s0_eval_warp(s0t, fields + (long)chain * Ltrot * n_opv, n, nt, Ltrot, n_opv); if (s0_warp == 1 && lane == 0) s0_s = s0_local; }

    // ---------- similarity before the update: UP = Op_Wrapup N_type 1 (old field); DOWN = Op_Wrapdo N_type 2
    if (isdiag) {
      if (UP && warp == 0) {
        for (int f = 0; f < F; ++f) {
          const VopDev<T>* op = op0 + f;
          for (int a = 0; a < k; ++a) {
            const T e = cont ? exp_(op->gE[a] * phi_old) : op->E_exp[a][s_old + 2]; const T ei = one_<T>() / e; const int p = op->P[a];
            for (int kk = lane; kk < nd; kk += 32) { Xs[((long)f * KD + kk) * ldx + p] = Xs[((long)f * KD + kk) * ldx + p] * e; Ys[((long)f * KD + kk) * ldx + p] = Ys[((long)f * KD + kk) * ldx + p] * ei; }
            if (lane == 0) { dl[f * N + p] = dl[f * N + p] * e; dr[f * N + p] = dr[f * N + p] * ei; }
          }
        }
        __syncwarp();
      }
    } else {
      for (int f = 0; f < F; ++f) {
        const VopDev<T>* op = op0 + f;
        T AL[ALF_KMAX * ALF_KMAX], AR[ALF_KMAX * ALF_KMAX];
        for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) {
          if (UP) {   // AL = diag(e) U^H ; AR = U diag(1/e)
            const T ea = op->E_exp[a][s_old + 2], eb = op->E_exp[b][s_old + 2];
            AL[a + b * ALF_KMAX] = ea * conj_(op->U[b + a * ALF_KMAX]);
            AR[a + b * ALF_KMAX] = op->U[a + b * ALF_KMAX] * (one_<T>() / eb);
          } else {    // AL = U^H ; AR = U
            AL[a + b * ALF_KMAX] = conj_(op->U[b + a * ALF_KMAX]);
            AR[a + b * ALF_KMAX] = op->U[a + b * ALF_KMAX];
          }
        }
        similarity_immediate<T>(Gc + f * sG, N, ldg, Xs + (long)f * KD * ldx, Ys + (long)f * KD * ldx, ldx, nd, dl + f * N, dr + f * N, op->P, k, AL, AR);
        for (int a = tid; a < k; a += nthr) gdiag[f * N + op->P[a]] = Gc[f * sG + op->P[a] + (long)op->P[a] * ldg];
      }
      __syncthreads();
    }

    // ---------- warp 0: G_cur(P,P), proposal, ratio, Metropolis decision
    if (warp == 0) {
      for (int f = 0; f < F; ++f) {
        const VopDev<T>* op = op0 + f; const int nz = op->nnz;
        for (int a = 0; a < nz; ++a) for (int b = 0; b < nz; ++b) {
          const int pa = op->P[a], pb = op->P[b];
          T s = zero_<T>();
          for (int kk = lane; kk < nd; kk += 32) fma_(s, Xs[((long)f * KD + kk) * ldx + pa], Ys[((long)f * KD + kk) * ldx + pb]);
          s = warp_sum(s);
          if (lane == 0) {
            T g0 = (a == b) ? gdiag[f * N + pa] : Gc[f * sG + pa + (long)pb * ldg];
            gpp_s[f][a][b] = (dl[f * N + pa] * g0) * dr[f * N + pb] - s;
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        // --- proposal: nsigma%flip (Fields_mod.F90:173-217), types 1 and 2
        int s_new; double phi_new = 0.0;
        if (cont) { s_new = s_old; phi_new = phi_old + amplitude * (r.ranf() - 0.5); }      // Fields_mod.F90:185-186
        else if (type == 1) s_new = -s_old; else s_new = ft.flip[s_old + 2][r.nranf(3)];
        double S0_ratio = s0t.on ? (s0_warp == 0 ? s0_local : s0_s) : 1.0, T0_proposal = 1.5, T0_Proposal_ratio = 1.0;   // Propose_S0 only for type 1 (Wrapgr_mod.F90:127-132)
        if (cont && s0_gaussian) S0_ratio = exp((-phi_new * phi_new + phi_old * phi_old) / 2.0);      // Hamiltonian_Hubbard_smod.F90:880-882
        if (propose_s0 && type == 1) { T0_proposal = 1.0 - 1.0 / (1.0 + S0_ratio); T0_Proposal_ratio = 1.0 / S0_ratio; }
        int acc = 0;
        if (T0_proposal > r.ranf()) {
          cplx ratiotot = cplx(1.0, 0.0);
          for (int f = 0; f < F; ++f) {
            const VopDev<T>* op = op0 + f; const int nz = op->nnz;
            T Mat[ALF_KMAX][ALF_KMAX]; T d0 = zero_<T>();
            for (int m = 0; m < nz; ++m) {
              const T d = cont ? exp_(op->gE[m] * (phi_new - phi_old)) - one_<T>() : op->delta[m][s_old + 2][s_new + 2];
              if (m == 0) d0 = d;
              for (int q = 0; q < nz; ++q) Mat[q][m] = -(d * gpp_s[f][q][m]);
              Mat[m][m] = Mat[m][m] + (d + one_<T>());
            }
            T D;
            if (nz == 0) D = one_<T>();
            else if (nz == 1) D = Mat[0][0];
            else if (nz == 2) {
              T s1 = Mat[0][0] * Mat[1][1], s2 = Mat[1][0] * Mat[0][1];
              if (abs_(s1) > abs_(s2)) D = s1 * (one_<T>() - s2 / s1); else D = s2 * (s1 / s2 - one_<T>());
            } else {   // LU without pivoting refinement is enough for k <= 4: partial pivoting
              D = one_<T>();
              for (int c = 0; c < nz; ++c) {
                int pv = c; double best = abs_(Mat[c][c]);
                for (int q = c + 1; q < nz; ++q) if (abs_(Mat[q][c]) > best) { best = abs_(Mat[q][c]); pv = q; }
                if (pv != c) { for (int q = 0; q < nz; ++q) { T t = Mat[c][q]; Mat[c][q] = Mat[pv][q]; Mat[pv][q] = t; } D = -D; }
                D = D * Mat[c][c];
                for (int q = c + 1; q < nz; ++q) { T l = Mat[q][c] / Mat[c][c]; for (int w = c; w < nz; ++w) Mat[q][w] = Mat[q][w] - l * Mat[c][w]; }
              }
            }
            const T rf = D * (cont ? exp_(op->galpha * (phi_new - phi_old)) : op->expalpha[s_old + 2][s_new + 2]);
            ratiotot = ratiotot * cplx(real_(rf), imag_(rf));
            if (nz >= 1) xfac_s[f] = d0 / Mat[0][0];   // only used for nnz == 1
          }
          cplx rt = ratiotot;
          for (int q = 1; q < n_sun; ++q) rt = rt * ratiotot;
          const double gr = cont ? 1.0 : ft.gama[type][s_new + 2] / ft.gama[type][s_old + 2];
          rt = rt * gr;
          const cplx pr = ph * rt;
          const double weight = S0_ratio * T0_Proposal_ratio * fabs(pr.x / ph.x);
          n_prop++;
          if (weight > r.ranf()) {
            acc = 1; n_acc++;
            const double ar = abs_(rt);
            ph = ph * cplx(rt.x / ar, rt.y / ar);
          }
          if (acclog) acclog[(long)chain * n_opv + log_off + step] = (uint8_t)acc;
        } else if (acclog) acclog[(long)chain * n_opv + log_off + step] = 2;
        cb->accept = acc; cb->s_new = s_new; cb->phi_new = phi_new;
        if (acc) { if (cont) fc[n] = phi_new; else fld[n] = (int8_t)s_new; }
      }
    }
    __syncthreads();
    const int accepted = cb->accept;
    const int s_cur = accepted ? cb->s_new : s_old;
    const double phi_cur = (cont && accepted) ? cb->phi_new : phi_old;

    // ---------- accepted: append the rank-1 factors (sequentially over the non-zero eigen-directions)
    if (accepted) {
      const int nzmax = op0->nnz;   // same for all flavors in the shipped models; flavors with fewer are skipped below
      for (int a = 0; a < nzmax; ++a) {
        for (int f = 0; f < F; ++f) {
          const VopDev<T>* op = op0 + f;
          if (a >= op->nnz) continue;
          const int p = op->P[a];
          T* X = Xs + (long)f * KD * ldx; T* Y = Ys + (long)f * KD * ldx;
          const T* G0 = Gc + f * sG;
          T xf;
          if (op->nnz == 1) xf = xfac_s[f];
          else {
            // current G(p,p) including the sub-updates already appended for this vertex
            if (warp == 0) {
              T s = zero_<T>();
              for (int kk = lane; kk < nd; kk += 32) fma_(s, X[(long)kk * ldx + p], Y[(long)kk * ldx + p]);
              s = warp_sum(s);
              if (lane == 0) {
                T gpp = (dl[f * N + p] * G0[p + (long)p * ldg]) * dr[f * N + p] - s;
                T d = op->delta[a][s_old + 2][s_cur + 2];
                bc_s[0] = d / (one_<T>() + (one_<T>() - gpp) * d);
              }
            }
            __syncthreads();
            xf = bc_s[0];
          }
          const T dlp = dl[f * N + p], drp = dr[f * N + p];
          for (int t = tid; t < 2 * N; t += nthr) {
            if (t < N) {          // column p of G_cur  ->  X(:, nd) = xfac * G_cur(:, p)
              T s = (dl[f * N + t] * G0[t + (long)p * ldg]) * drp;
              for (int kk = 0; kk < nd; ++kk) s = s - X[(long)kk * ldx + t] * Y[(long)kk * ldx + p];
              X[(long)nd * ldx + t] = xf * s;
            } else {              // row p of G_cur     ->  Y(:, nd) = e_p - G_cur(p, :)
              const int j = t - N;
              T s = (dlp * G0[p + (long)j * ldg]) * dr[f * N + j];
              for (int kk = 0; kk < nd; ++kk) s = s - X[(long)kk * ldx + p] * Y[(long)kk * ldx + j];
              Y[(long)nd * ldx + j] = ((j == p) ? one_<T>() : zero_<T>()) - s;
            }
          }
        }
        __syncthreads();
        nd += 1;
        if (nd == KD) {
          for (int f = 0; f < F; ++f)
            flush_flavor<T>(Gc + f * sG, N, ldg, Xs + (long)f * KD * ldx, Ys + (long)f * KD * ldx, ldx, nd, dl + f * N, dr + f * N);
          nd = 0; n_flush++;
          __syncthreads();
          for (int e = tid; e < F * N; e += nthr) { int f = e / N, i = e % N; gdiag[e] = Gc[f * sG + i + (long)i * ldg]; }
          __syncthreads();
        }
      }
    }

    // ---------- similarity after the update: UP = Op_Wrapup N_type 2; DOWN = Op_Wrapdo N_type 1 with the NEW field
    if (isdiag) {
      if (!UP && warp == 0) {
        for (int f = 0; f < F; ++f) {
          const VopDev<T>* op = op0 + f;
          for (int a = 0; a < k; ++a) {
            const T e = cont ? exp_(op->gE[a] * phi_cur) : op->E_exp[a][s_cur + 2]; const T ei = one_<T>() / e; const int p = op->P[a];
            for (int kk = lane; kk < nd; kk += 32) { Xs[((long)f * KD + kk) * ldx + p] = Xs[((long)f * KD + kk) * ldx + p] * ei; Ys[((long)f * KD + kk) * ldx + p] = Ys[((long)f * KD + kk) * ldx + p] * e; }
            if (lane == 0) { dl[f * N + p] = dl[f * N + p] * ei; dr[f * N + p] = dr[f * N + p] * e; }
          }
        }
        __syncwarp();
      }
    } else {
      for (int f = 0; f < F; ++f) {
        const VopDev<T>* op = op0 + f;
        T AL[ALF_KMAX * ALF_KMAX], AR[ALF_KMAX * ALF_KMAX];
        for (int a = 0; a < k; ++a) for (int b = 0; b < k; ++b) {
          if (UP) {   // AL = U ; AR = U^H
            AL[a + b * ALF_KMAX] = op->U[a + b * ALF_KMAX];
            AR[a + b * ALF_KMAX] = conj_(op->U[b + a * ALF_KMAX]);
          } else {    // AL = U diag(1/e) ; AR = diag(e) U^H   (new field)
            const T ea = op->E_exp[a][s_cur + 2], eb = op->E_exp[b][s_cur + 2];
            AL[a + b * ALF_KMAX] = op->U[a + b * ALF_KMAX] * (one_<T>() / eb);
            AR[a + b * ALF_KMAX] = ea * conj_(op->U[b + a * ALF_KMAX]);
          }
        }
        similarity_immediate<T>(Gc + f * sG, N, ldg, Xs + (long)f * KD * ldx, Ys + (long)f * KD * ldx, ldx, nd, dl + f * N, dr + f * N, op->P, k, AL, AR);
        for (int a = tid; a < k; a += nthr) gdiag[f * N + op->P[a]] = Gc[f * sG + op->P[a] + (long)op->P[a] * ldg];
      }
      __syncthreads();
    }
  }
  // ---------- end of slice: materialise G
  __syncthreads();
  for (int f = 0; f < F; ++f)
    flush_flavor<T>(Gc + f * sG, N, ldg, Xs + (long)f * KD * ldx, Ys + (long)f * KD * ldx, ldx, nd, dl + f * N, dr + f * N);
  if (stage_g) { __syncthreads(); for (long e = tid; e < (long)F * N * N; e += nthr) { const long fq = e / ((long)N * N), q = e - fq * N * N; Gglob[e] = Gc[fq * sG + (q % N) + (q / N) * ldg]; } }
  if (tid == 0) {
    rng[chain * 4 + 0] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
    phase[chain] = ph;
    counters[chain * 4 + 0] += n_prop;   // NC_up
    counters[chain * 4 + 1] += n_acc;    // ACC_up
    counters[chain * 4 + 2] += (unsigned long long)cnt;     // NC_eff_up
    counters[chain * 4 + 3] += n_acc;    // ACC_eff_up
    if (!stage_g) counters[4 * (long)gridDim.x + chain] += n_flush;   // rewrites of G0 in global memory (measurement support)
  }
}
