// Fast slice kernel for GROUPS of interaction vertices with pairwise disjoint supports that are diagonal single-site operators
// (k = 1: every Hubbard variant, Kondo U_f) or, in pair mode, k = 2 operators in their eigenbasis (Kondo J_K; see the kernel's header):
// same semantics as k_wrapgr (alf_update.cuh) = the n-loop of WRAPGRUP / WRAPGRDO (Prog/Wrapgr_mod.F90:115-146, 191-237) with
// Op_Wrapup/Op_Wrapdo (Prog/Operator_mod.F90:743-951), Fields%flip (Prog/Fields_mod.F90:173-217), Upgrade2
// (Prog/upgrade_mod.F90:105-302) and the counters of Prog/control_mod.F90:164-176 -- restructured around what is sequential.
//
//  * Every field of a slice is visited exactly once, so the old field value, the proposed value (type 2: one nranf(3) draw)
//    and all uniforms of the slice are known up front: a prologue draws the chain's xoshiro256** stream in the reference's
//    order ([flip] -> [proposal] -> [acceptance] per visit) and tabulates per visit p, exp(g dphi E) - 1, exp(g dphi alpha),
//    gamma ratio and the acceptance uniform in shared memory.
//  * Accepted flips are kept as delayed rank-1 factors  G_cur = DL G0 DR - X Y^T  (X, Y, DL, DR in shared memory, G0 in
//    global memory); every KD accepts G0 is updated by a rank-KD FP64 tensor-core (DMMA m8n8k4) product.
//  * The slice is processed in WINDOWS of ALF_WIN consecutive visits ("submatrix updates").  The Metropolis decisions of a
//    window need only the ALF_WIN x ALF_WIN block G_cur(P_w, P_w): it is gathered into shared memory, and ONE warp runs the
//    window's decisions with rank-1 updates of that small block -- no block-wide barrier and no global memory access per visit.
//    The warp records, per accepted flip, the new factor restricted to the window's sites.
//  * After the window all threads build the full-length factors of the accepted flips in one pass (thread i owns element i of
//    every factor; the only cross-thread data are the window-restricted values recorded by the decision warp), so the G0
//    columns / rows are fetched for ACCEPTED flips only, all at once (one memory latency per window, not per visit).
#pragma once
#include "alf_update.cuh"


// G0 <- DL G0 DR - X Y^T, then DL = DR = 1.  Real: warp tiles of 32 x 32; the accumulator fragments are INITIALISED with the
// scaled G0 tile (32 independent loads per thread issued up front, so the HBM/L2 latency is paid once per tile), then
// (-X) Y^T is accumulated with DMMA m8n8k4 and the tile is stored.  X, Y: [k][ldx] with rows nd .. nd4-1 zeroed by the caller
// (nd4 = nd rounded up to 4); ldx % 16 == 4 makes the fragment loads bank-conflict free.
// rev = 1 walks the tiles in the opposite order: consecutive flushes alternate, so the tiles written last by one flush (still
// in L2: the batch of Green functions is only slightly larger than the 126 MB L2) are the first ones the next flush reads.
// Warp tiles are 32 x 16 and software-pipelined two deep: the 16 loads of the next tile are in flight while the current one
// is scaled, updated and stored, so every warp keeps HBM requests outstanding for the whole flush.
static __device__ __noinline__ void flush_g0(double* __restrict__ G0, int N, const double* __restrict__ X, const double* __restrict__ Y, int ldx, int nd4,
                                      double* __restrict__ dl, double* __restrict__ dr, int rev) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int tm = (N + 31) / 32, tn = (N + 15) / 16, tiles = tm * tn;
  const int g = lane >> 2, q = lane & 3;
  auto load = [&](int tt, double (&c)[4][2][2]) {
    const int t = rev ? tiles - 1 - tt : tt;
    const int i0 = (t % tm) * 32, j0 = (t / tm) * 16;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + 8 * b + 2 * q + h;
#pragma unroll
        for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; c[a][b][h] = (i < N && j < N) ? G0[i + (long)j * N] : 0.0; }
      }
  };
  auto finish = [&](int tt, double (&c)[4][2][2]) {
    const int t = rev ? tiles - 1 - tt : tt;
    const int i0 = (t % tm) * 32, j0 = (t / tm) * 16;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + 8 * b + 2 * q + h; const double drj = (j < N) ? dr[j] : 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; c[a][b][h] = ((i < N ? dl[i] : 0.0) * c[a][b][h]) * drj; }
      }
    for (int k0 = 0; k0 < nd4; k0 += 4) {
      double av[4], bv[2];
      const double* xk = X + (long)(k0 + q) * ldx; const double* yk = Y + (long)(k0 + q) * ldx;
#pragma unroll
      for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; av[a] = (i < N) ? -xk[i] : 0.0; }
#pragma unroll
      for (int b = 0; b < 2; ++b) { const int j = j0 + 8 * b + g; bv[b] = (j < N) ? yk[j] : 0.0; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) dmma884(c[a][b][0], c[a][b][1], av[a], bv[b]);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + 8 * b + 2 * q + h;
#pragma unroll
        for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; if (i < N && j < N) G0[i + (long)j * N] = c[a][b][h]; }
      }
  };
  double cA[4][2][2], cB[4][2][2];
  int tt = warp;
  if (tt < tiles) {
    load(tt, cA);
    while (true) {
      const int t1 = tt + nw;
      if (t1 < tiles) load(t1, cB);
      finish(tt, cA);
      if (t1 >= tiles) break;
      const int t2 = t1 + nw;
      if (t2 < tiles) load(t2, cA);
      finish(t1, cB);
      if (t2 >= tiles) break;
      tt = t2;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) { dl[i] = 1.0; dr[i] = 1.0; }
}
// complex: register-tiled FMA version of alf_update.cuh
// complex: the same 32 x 16 warp tiles on split real / imaginary accumulators, every complex product as four real DMMA products
static __device__ __noinline__ void flush_g0(cplx* __restrict__ G0, int N, const cplx* __restrict__ X, const cplx* __restrict__ Y, int ldx, int nd4,
                                         cplx* __restrict__ dl, cplx* __restrict__ dr, int rev) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int tm = (N + 31) / 32, tn = (N + 15) / 16, tiles = tm * tn;
  const int g = lane >> 2, q = lane & 3;
  for (int tt = warp; tt < tiles; tt += nw) {
    const int t = rev ? tiles - 1 - tt : tt;
    const int i0 = (t % tm) * 32, j0 = (t / tm) * 16;
    double cr[4][2][2], ci[4][2][2];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + 8 * b + 2 * q + h; const cplx drj = (j < N) ? dr[j] : cplx(0.0, 0.0);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = i0 + 8 * a + g;
          cplx v = cplx(0.0, 0.0);
          if (i < N && j < N) v = (dl[i] * G0[i + (long)j * N]) * drj;
          cr[a][b][h] = v.x; ci[a][b][h] = v.y;
        }
      }
    for (int k0 = 0; k0 < nd4; k0 += 4) {
      cplx av[4], bv[2];
      const cplx* xk = X + (long)(k0 + q) * ldx; const cplx* yk = Y + (long)(k0 + q) * ldx;
#pragma unroll
      for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; av[a] = (i < N) ? -xk[i] : cplx(0.0, 0.0); }
#pragma unroll
      for (int b = 0; b < 2; ++b) { const int j = j0 + 8 * b + g; bv[b] = (j < N) ? yk[j] : cplx(0.0, 0.0); }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          dmma884(cr[a][b][0], cr[a][b][1], av[a].x, bv[b].x); dmma884(cr[a][b][0], cr[a][b][1], -av[a].y, bv[b].y);
          dmma884(ci[a][b][0], ci[a][b][1], av[a].x, bv[b].y); dmma884(ci[a][b][0], ci[a][b][1], av[a].y, bv[b].x);
        }
    }
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = j0 + 8 * b + 2 * q + h;
#pragma unroll
        for (int a = 0; a < 4; ++a) { const int i = i0 + 8 * a + g; if (i < N && j < N) G0[i + (long)j * N] = cplx(cr[a][b][h], ci[a][b][h]); }
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) { dl[i] = cplx(1.0, 0.0); dr[i] = cplx(1.0, 0.0); }
}

// Weight = |Re(Phase R) / Re(Phase)| (upgrade_mod.F90:210); for real arithmetic Phase = +-1 and the weight is |R|.
__device__ __forceinline__ double upd_weight(const cplx&, double rt) { return fabs(rt); }
__device__ __forceinline__ double upd_weight(const cplx& ph, cplx rt) { const cplx pr = ph * rt; return fabs(pr.x / ph.x); }
__device__ __forceinline__ void upd_phase(cplx& ph, double rt) { if (rt < 0.0) ph = cplx(-ph.x, -ph.y); }                       // Phase * R/|R|
__device__ __forceinline__ void upd_phase(cplx& ph, cplx rt) { const double ar = abs_(rt); ph = ph * cplx(rt.x / ar, rt.y / ar); }

#define ALF_WIN 16          // visits per window
#ifdef ALF_UPD_PROF      // experimental build only: clock64 accounting of the phases of k_wrapgr_fast (thread 0, summed over CTAs)
__device__ unsigned long long g_upd_prof[16];
#define UPP_T0 long long up_t = clock64();
#define UPP_ACC(i) if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_upd_prof[i], (unsigned long long)(t_ - up_t)); up_t = t_; }
#else
#define UPP_T0
#define UPP_ACC(i)
#endif
#define ALF_CH 8            // accepted flips whose G0 column/row are in flight at once in the bulk phase

// The kernel visits the vertices n0 .. n0 + cnt - 1 of the slice (a GROUP of vertices with pairwise disjoint supports; Mtot =
// size(Op_V,1) is the stride of the field array and of the accept log, log_off the number of vertices visited before the group).
// PAIR = 1: the vertices are k = 2 operators that are DIAGONAL in the basis the Green function has been rotated to (the caller
// applies U^dagger . U of the whole group before and U . U^dagger after, alf_engine.cuh); each vertex then is a pair of
// consecutive single-site "pseudo-visits": the Metropolis decision is taken once with the 2 x 2 determinant of
// upgrade_mod.F90:178-188, the two rank-1 updates follow (second one with the G already updated by the first).
template <typename T, int UP, int IPT, int PAIR>
__global__ void __launch_bounds__(512, 1) k_wrapgr_fast(T* __restrict__ G, int N, int F, int n_sun, int n0, int cnt, int Mtot, int log_off,
                                                        const VopDev<T>* __restrict__ vops, FieldTabDev ft,
                                                        int8_t* __restrict__ fields, int Ltrot, int nt, uint64_t* __restrict__ rng, cplx* __restrict__ phase,
                                                        unsigned long long* __restrict__ counters, int KD, int ldx, uint8_t* __restrict__ acclog) {
  const int M = PAIR ? 2 * cnt : cnt;                 // pseudo-visits (one site each)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int win_nvis, win_nacc;
  __shared__ int acc_vis[ALF_WIN];
  __shared__ int8_t acc_flag[ALF_WIN];
  constexpr int W = ALF_WIN, LW = ALF_WIN + 1;
  const int chain = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  // ---- shared memory carve-up
  T* Xs = reinterpret_cast<T*>(smem_raw);
  T* Ys = Xs + (long)F * KD * ldx;
  T* dl = Ys + (long)F * KD * ldx;
  T* dr = dl + (long)F * N;
  T* st_d = dr + (long)F * N;             // [f][s]  exp(g (phi(s')-phi(s)) E) - 1
  T* st_c1 = st_d + (long)F * M;          // [f][s]  (delta + 1) exp(g dphi alpha)       ratio_f = c1 - c2 G_pp
  T* st_c2 = st_c1 + (long)F * M;         // [f][s]  delta exp(g dphi alpha)
  T* st_eold = st_c2 + (long)F * M;       // [f][s]  exp(g phi(s) E)
  T* st_enew = st_eold + (long)F * M;     // [f][s]  exp(g phi(s') E)
  T* st_eoldi = st_enew + (long)F * M;    // reciprocal (no division on the per-visit path)
  T* Gw = st_eoldi + (long)F * M;         // [f][a][b], leading dimension LW: G_cur on the window's sites
  T* Xw = Gw + (long)F * W * LW;          // [k][f][a]: new factor k at the window's sites (as created)
  T* Yw = Xw + (long)W * F * W;
  T* win_xf = Yw + (long)W * F * W;       // [k][f]: delta / (1 + delta (1 - G_pp))   (upgrade_mod.F90:233-240)
  double* st_u = reinterpret_cast<double*>(win_xf + (long)W * F);   // acceptance uniform
  double* st_gr = st_u + M;               // gamma(s')/gamma(s)
  int* st_p = reinterpret_cast<int*>(st_gr + M);                      // [f][s]
  int8_t* st_sold = reinterpret_cast<int8_t*>(st_p + (long)F * M);
  int8_t* st_snew = st_sold + M;
  int8_t* st_type = st_snew + M;

  UPP_T0
  T* Gc = G + (long)chain * F * N * N;
  int8_t* fld = fields + ((long)chain * Ltrot + (nt - 1)) * Mtot;
  auto vertex_of = [&](int s) { const int pv = PAIR ? (s >> 1) : s; return UP ? n0 + pv : n0 + cnt - 1 - pv; };
  const int items = F * N;

  // ---- prologue A: field values, types, sites; DL = DR = 1
  for (int s = tid; s < M; s += nthr) {
    const int n = vertex_of(s), a = PAIR ? (s & 1) : 0;
    st_sold[s] = fld[n]; st_type[s] = (int8_t)vops[(long)n * F].type;
    for (int f = 0; f < F; ++f) st_p[f * M + s] = vops[(long)n * F + f].P[a];
  }
  for (int e = tid; e < items; e += nthr) { dl[e] = one_<T>(); dr[e] = one_<T>(); }
  __syncthreads();
  // ---- prologue B: the chain's random stream for this slice, in the reference's order
  if (tid == 0) {
    Xoshiro r; r.s0 = rng[chain * 4 + 0]; r.s1 = rng[chain * 4 + 1]; r.s2 = rng[chain * 4 + 2]; r.s3 = rng[chain * 4 + 3];
    for (int s = 0; s < M; s += (PAIR ? 2 : 1)) {
      const int so = st_sold[s];
      int sn;
      if (st_type[s] == 1) sn = -so; else sn = ft.flip[so + 2][r.nranf(3)];      // Fields_mod.F90:173-217
      (void)r.ranf();                                                              // proposal draw: T0_proposal = 1.5 > ranf() always (Wrapgr_mod.F90:133)
      st_u[s] = r.ranf();                                                          // acceptance draw (upgrade_mod.F90:222)
      st_snew[s] = (int8_t)sn;
      if (PAIR) { st_u[s + 1] = st_u[s]; st_snew[s + 1] = (int8_t)sn; }
    }
    rng[chain * 4 + 0] = r.s0; rng[chain * 4 + 1] = r.s1; rng[chain * 4 + 2] = r.s2; rng[chain * 4 + 3] = r.s3;
  }
  // window-block gather of G0 (raw, without the pending factors): element e = (f, b, a), a fastest
  constexpr int GIT = (ALF_FMAX * W * W + 511) / 512;
  T gwraw[GIT];
  auto gw_prefetch = [&](int v0) {
#pragma unroll
    for (int it = 0; it < GIT; ++it) {
      const int e = tid + it * nthr;
      gwraw[it] = zero_<T>();
      if (e < F * W * W) {
        const int a = e % W, b = (e / W) % W, f = e / (W * W);
        if (v0 + a < M && v0 + b < M) gwraw[it] = Gc[(long)f * N * N + st_p[f * M + v0 + a] + (long)st_p[f * M + v0 + b] * N];
      }
    }
  };
  gw_prefetch(0);
  __syncthreads();
  // ---- prologue C: per-visit tables
  for (int e = tid; e < F * M; e += nthr) {
    const int f = e / M, s = e % M; const int n = vertex_of(s), a = PAIR ? (s & 1) : 0;
    const VopDev<T>* op = vops + (long)n * F + f;
    const int so = st_sold[s] + 2, sn = st_snew[s] + 2;
    const T d = op->delta[a][so][sn], ea = op->expalpha[so][sn];
    st_d[e] = d; st_eold[e] = op->E_exp[a][so]; st_enew[e] = op->E_exp[a][sn];
    if (PAIR) { st_c1[e] = ea; st_c2[e] = zero_<T>(); }                           // the pair's ratio is a 2 x 2 determinant times exp(g dphi alpha)
    else { st_c1[e] = (d + one_<T>()) * ea; st_c2[e] = d * ea; }
    st_eoldi[e] = one_<T>() / st_eold[e];
    if (f == 0) { const int ty = st_type[s]; st_gr[s] = ft.gama[ty][sn] / ft.gama[ty][so]; }
  }
  __syncthreads();

  UPP_ACC(0)
  cplx ph = phase[chain];
  unsigned long long n_acc = 0, n_flush = 0;
  int nd = 0;
  int flush_rev = (nt + (UP ? 0 : 1)) & 1;      // alternate the tile order between consecutive flushes (also across slices)

  auto do_flush = [&]() {
    const int nd4 = (nd + 3) & ~3;
    for (int e = tid; e < F * (nd4 - nd) * ldx; e += nthr) {
      const int f = e / ((nd4 - nd) * ldx), r = e % ((nd4 - nd) * ldx);
      Xs[((long)f * KD + nd) * ldx + r] = zero_<T>(); Ys[((long)f * KD + nd) * ldx + r] = zero_<T>();
    }
    __syncthreads();
    for (int ff = 0; ff < F; ++ff) {
      const int f = flush_rev ? F - 1 - ff : ff;
      flush_g0(Gc + (long)f * N * N, N, Xs + (long)f * KD * ldx, Ys + (long)f * KD * ldx, ldx, nd4, dl + f * N, dr + f * N, flush_rev);
    }
    flush_rev ^= 1;
    nd = 0; n_flush++;
    __syncthreads();
  };

  int v0 = 0;
  bool raw_valid = true;
  while (v0 < M) {
    if (nd == KD) { UPP_ACC(1) do_flush(); raw_valid = false; UPP_ACC(2) }
    if (!raw_valid) { gw_prefetch(v0); raw_valid = true; }
    const int Wn = min(W, M - v0), cap = min(KD - nd, W);
    // ---- (a) G_cur on the window's sites: DL G0 DR - X Y^T restricted to (P_w, P_w)
#pragma unroll
    for (int it = 0; it < GIT; ++it) {
      const int e = tid + it * nthr;
      if (e < F * W * W) {
        const int a = e % W, b = (e / W) % W, f = e / (W * W);
        T v = zero_<T>();
        if (a < Wn && b < Wn) {
          const int pa = st_p[f * M + v0 + a], pb = st_p[f * M + v0 + b];
          const T* X = Xs + (long)f * KD * ldx; const T* Y = Ys + (long)f * KD * ldx;
          v = (dl[f * N + pa] * gwraw[it]) * dr[f * N + pb];
          for (int j = 0; j < nd; ++j) v = v - X[(long)j * ldx + pa] * Y[(long)j * ldx + pb];
        }
        Gw[(f * W + a) * LW + b] = v;
      }
    }
    __syncthreads();
    UPP_ACC(3)
    // ---- (b) one warp: the window's Metropolis decisions on the small block (Upgrade2 for rank-1 vertices).  This is the only
    // sequential part of the slice, so it is written for latency: per visit 1 shared load + a handful of dependent FP64 ops.
    if (warp == 0) {
      int nacc = 0, v = 0, pair_acc = 0;
      for (; v < Wn && nacc < cap; ++v) {
        const int s = v0 + v;
        T g[ALF_FMAX];
        int acc;
        const bool decide = !PAIR || !(v & 1);
        if (PAIR && decide && nacc + 2 > cap) break;                 // both factors of a pair go into the same batch
        for (int f = 0; f < F; ++f) g[f] = Gw[(f * W + v) * LW + v];
        if (decide) {
          T rfp = one_<T>();
          if (!PAIR) { for (int f = 0; f < F; ++f) rfp = rfp * (st_c1[f * M + s] - st_c2[f * M + s] * g[f]); }
          else {
            for (int f = 0; f < F; ++f) {     // Mat(n,m) = delta_nm (1 + d_m) - d_m G(P_n, P_m); 2 x 2 determinant as in upgrade_mod.F90:178-188
              const T d0 = st_d[f * M + s], d1 = st_d[f * M + s + 1];
              const T m00 = (d0 + one_<T>()) - d0 * g[f], m11 = (d1 + one_<T>()) - d1 * Gw[(f * W + v + 1) * LW + v + 1];
              const T m10 = -(d0 * Gw[(f * W + v + 1) * LW + v]), m01 = -(d1 * Gw[(f * W + v) * LW + v + 1]);
              const T s1 = m00 * m11, s2 = m10 * m01;
              const T D = (abs_(s1) > abs_(s2)) ? s1 * (one_<T>() - s2 / s1) : s2 * (s1 / s2 - one_<T>());
              rfp = rfp * (D * st_c1[f * M + s]);
            }
          }
          T rtT = rfp;
          for (int q = 1; q < n_sun; ++q) rtT = rtT * rfp;
          rtT = rtT * st_gr[s];
          const double weight = upd_weight(ph, rtT);
          acc = (weight > st_u[s]) ? 1 : 0;
          pair_acc = acc;
          if (lane == 0) {
            acc_flag[v] = (int8_t)acc; if (PAIR) acc_flag[v + 1] = (int8_t)acc;
            if (acc) fld[vertex_of(s)] = st_snew[s];
            if (acclog) acclog[(long)chain * Mtot + log_off + (PAIR ? (s >> 1) : s)] = (uint8_t)acc;
          }
          if (acc) { n_acc++; upd_phase(ph, rtT); }
        } else acc = pair_acc;
        if (lane == 0 && acc) acc_vis[nacc] = v;
        if (acc) {
          // the new factor on the window's sites (in the frame of this visit: UP applies Op_Wrapup N_type 1 to row / column v first),
          // then the rank-1 update of the entries that later visits of the window still read
          T yb[(ALF_FMAX * W + 31) / 32];
          int q = 0;
          for (int e = lane; e < F * W; e += 32, ++q) {
            const int f = e / W, b = e % W;
            const T d = st_d[f * M + s];
            const T xf = d / ((d + one_<T>()) - d * g[f]);
            T cv = Gw[(f * W + b) * LW + v], rv = Gw[(f * W + v) * LW + b];
            if (UP && b != v) { cv = cv * st_eoldi[f * M + s]; rv = rv * st_eold[f * M + s]; }
            yb[q] = ((b == v) ? one_<T>() : zero_<T>()) - rv;
            Xw[(nacc * F + f) * W + b] = xf * cv;
            Yw[(nacc * F + f) * W + b] = yb[q];
            if (b == 0) win_xf[nacc * F + f] = xf;
          }
          __syncwarp();
          q = 0;
          for (int e = lane; e < F * W; e += 32, ++q) {
            const int f = e / W, b = e % W;
            if (b > v) {
              const T* xw = Xw + (nacc * F + f) * W;
#pragma unroll
              for (int a = 1; a < W; ++a) if (a > v) Gw[(f * W + a) * LW + b] = Gw[(f * W + a) * LW + b] - xw[a] * yb[q];
            }
          }
          __syncwarp();
          nacc++;
        }
      }
      if (lane == 0) { win_nvis = v; win_nacc = nacc; }
    }
    __syncthreads();
    UPP_ACC(4)
    const int nvis = win_nvis, nacc = win_nacc, nd0 = nd;
    // ---- (c) all threads: full-length factors of the accepted flips; thread (f, i) owns element i of every factor of flavor f
    {
      T dli[IPT], dri[IPT], sxo[IPT], syo[IPT]; int ownv[IPT]; bool owndone[IPT];
#pragma unroll
      for (int it = 0; it < IPT; ++it) {
        const int item = tid + it * nthr; ownv[it] = -1; owndone[it] = true;
        sxo[it] = one_<T>(); syo[it] = one_<T>(); dli[it] = one_<T>(); dri[it] = one_<T>();
        if (item < items) {
          const int f = item / N, i = item - f * N;
          dli[it] = dl[item]; dri[it] = dr[item];
          for (int v = 0; v < nvis; ++v) if (st_p[f * M + v0 + v] == i) { ownv[it] = v; owndone[it] = false; }
        }
      }
      // own-site similarity: existing factors are rescaled lazily (sxo, syo), the ones created in this window immediately
      auto apply_own = [&](int it, int f, int i, int k_done) {
        const int so = v0 + ownv[it];
        T e, ei;
        if (UP) { e = st_eold[f * M + so]; ei = st_eoldi[f * M + so]; }
        else if (acc_flag[ownv[it]]) { ei = st_enew[f * M + so]; e = one_<T>() / ei; }            // Op_Wrapdo N_type 1 with the new field
        else { e = st_eoldi[f * M + so]; ei = st_eold[f * M + so]; }
        dli[it] = dli[it] * e; dri[it] = dri[it] * ei; sxo[it] = sxo[it] * e; syo[it] = syo[it] * ei;
        T* X = Xs + (long)f * KD * ldx; T* Y = Ys + (long)f * KD * ldx;
        for (int kk = 0; kk < k_done; ++kk) { X[(long)(nd0 + kk) * ldx + i] = X[(long)(nd0 + kk) * ldx + i] * e; Y[(long)(nd0 + kk) * ldx + i] = Y[(long)(nd0 + kk) * ldx + i] * ei; }
        owndone[it] = true;
      };
      for (int k0 = 0; k0 < nacc; k0 += ALF_CH) {
        T rc[IPT][ALF_CH], rr[IPT][ALF_CH];
#pragma unroll
        for (int it = 0; it < IPT; ++it) {
          const int item = tid + it * nthr;
#pragma unroll
          for (int c = 0; c < ALF_CH; ++c) {
            rc[it][c] = zero_<T>(); rr[it][c] = zero_<T>();
            if (item < items && k0 + c < nacc) {
              const int f = item / N, i = item - f * N; const int p = st_p[f * M + v0 + acc_vis[k0 + c]];
              const T* G0 = Gc + (long)f * N * N;
              rc[it][c] = G0[i + (long)p * N]; rr[it][c] = G0[p + (long)i * N];
            }
          }
        }
#pragma unroll
        for (int c = 0; c < ALF_CH; ++c) {
          const int k = k0 + c;
          if (k < nacc) {
            const int v = acc_vis[k], s = v0 + v;
#pragma unroll
            for (int it = 0; it < IPT; ++it) {
              const int item = tid + it * nthr;
              if (item < items) {
                const int f = item / N, i = item - f * N; const int p = st_p[f * M + s];
                if (!owndone[it] && (UP ? (ownv[it] <= v) : (ownv[it] < v))) apply_own(it, f, i, k);
                const T fx = UP ? st_eold[f * M + s] : one_<T>();       // frame of visit v at site p: X_j(p) * e, Y_j(p) / e, DL(p) = e, DR(p) = 1/e
                const T fy = UP ? st_eoldi[f * M + s] : one_<T>();
                const T* X = Xs + (long)f * KD * ldx; const T* Y = Ys + (long)f * KD * ldx;
                T colv = (dli[it] * rc[it][c]) * fy;
                T rowv = (fx * rr[it][c]) * dri[it];
                for (int j = 0; j < nd0; ++j) {
                  colv = colv - (X[(long)j * ldx + i] * sxo[it]) * (Y[(long)j * ldx + p] * fy);
                  rowv = rowv - (X[(long)j * ldx + p] * fx) * (Y[(long)j * ldx + i] * syo[it]);
                }
                for (int kk = 0; kk < k; ++kk) {
                  colv = colv - X[(long)(nd0 + kk) * ldx + i] * (Yw[(kk * F + f) * W + v] * fy);
                  rowv = rowv - (Xw[(kk * F + f) * W + v] * fx) * Y[(long)(nd0 + kk) * ldx + i];
                }
                const T xn = win_xf[k * F + f] * colv;
                const T yn = ((i == p) ? one_<T>() : zero_<T>()) - rowv;
                Xs[((long)f * KD + nd0 + k) * ldx + i] = xn; Ys[((long)f * KD + nd0 + k) * ldx + i] = yn;
              }
            }
          }
        }
      }
#pragma unroll
      for (int it = 0; it < IPT; ++it) {
        const int item = tid + it * nthr;
        if (item < items && !owndone[it]) { const int f = item / N, i = item - f * N; apply_own(it, f, i, nacc); }
      }
      __syncthreads();      // every read of an old factor at a foreign site is done: write the lazy rescaling back
#pragma unroll
      for (int it = 0; it < IPT; ++it) {
        const int item = tid + it * nthr;
        if (item < items) {
          const int f = item / N, i = item - f * N;
          if (ownv[it] >= 0) {
            T* X = Xs + (long)f * KD * ldx; T* Y = Ys + (long)f * KD * ldx;
            for (int j = 0; j < nd0; ++j) { X[(long)j * ldx + i] = X[(long)j * ldx + i] * sxo[it]; Y[(long)j * ldx + i] = Y[(long)j * ldx + i] * syo[it]; }
            dl[item] = dli[it]; dr[item] = dri[it];
          }
        }
      }
    }
    nd = nd0 + nacc; v0 += nvis;
    if (nd < KD && v0 < M) gw_prefetch(v0);       // next window's raw block (G0 is unchanged unless a flush comes first)
    __syncthreads();
    UPP_ACC(5)
  }
  do_flush();
  UPP_ACC(2)
  if (tid == 0) {
    phase[chain] = ph;
    counters[chain * 4 + 0] += (unsigned long long)cnt; // NC_up
    counters[chain * 4 + 1] += n_acc;                   // ACC_up
    counters[chain * 4 + 2] += (unsigned long long)cnt; // NC_eff_up
    counters[chain * 4 + 3] += n_acc;                   // ACC_eff_up
    counters[4 * (long)gridDim.x + chain] += n_flush;   // rank-KD rewrites of G0 (measurement support: algorithmic HBM bytes of this kernel)
  }
}
