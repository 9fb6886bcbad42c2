// Device-side time-displaced lattice observables: what ham%ObserT of the shipped Hamiltonians accumulates through
// Predefined_Obs_tau_Green/SpinMz/SpinSUN/Den_measure (Prog/Predefined_Obs_mod.F90:337-594; Hamiltonian_Hubbard_smod.F90:838-856)
// every time TAU_M / Tau_p hands it G(tau,0), G(0,tau), G(0,0), G(tau,tau) (tau_m_mod.F90:115-177, tau_p_mod.F90:173-250):
//     Obs(imj(I,J), nt, no_I, no_J) += f(I1, J1) * ZP * ZS          with I1, J1 over all sites, imj = lattice point r_I - r_J,
// channel 0 Green  f = sum_nf GT0(I1,J1,nf) / N_FL
// channel 1 SpinZ  Mz (N_FL = 2): (GTT(I1,I1,1) - GTT(I1,I1,2)) (G00(J1,J1,1) - G00(J1,J1,2)) - sum_nf G0T(J1,I1,nf) GT0(I1,J1,nf)
//                  SU(N) (N_FL = 1): - N_SUN G0T(J1,I1,1) GT0(I1,J1,1)
// channel 2 SpinXY Mz only: - G0T(J1,I1,1) GT0(I1,J1,2) - G0T(J1,I1,2) GT0(I1,J1,1)                    (SpinT = (2 XY + Z) / 3 is linear in these)
// channel 3 Den    ZI ZJ - N_SUN sum_nf G0T(J1,I1,nf) GT0(I1,J1,nf),  ZI = N_SUN sum_nf (1 - GTT(I1,I1,nf)), ZJ likewise with G00
// backgrounds Obs_Latt0(no_I): [0] SpinZ: GTT(I1,I1,2) - GTT(I1,I1,1) (Mz), [1] Den: ZI.
// One CTA per chain walks the N x N site pairs in 32 x 32 tiles (both G(tau,0) and the transposed G(0,tau) are read coalesced),
// bins into shared memory by imj, then adds its bins, times ZP ZS of the chain, to the global accumulators (sum over chains).
#pragma once
#include "alf_types.cuh"

#define OBST_NCH 4

struct LattDev { int n_unit = 0, norb = 1; const int* cell = nullptr; const int* orb = nullptr; const int* imj = nullptr; int shift32 = 0; };   // 0-based tables; shift32: see k_obs_tau_diag

// acc: [ch][nt][no_J][no_I][imj] complex; bg: [2][nt][norb] complex; cnt: [0] N (chain-measurements at nt = 0), [1] sum ZS
// EQ = 1: the equal-time variants Predefined_Obs_eq_Green / SpinMz / SpinSUN / Den_measure (Predefined_Obs_mod.F90:77-325) on the inputs
// GT0 = GTT = G00 = G, G0T = G - 1 (so that -G0T(J1,I1) = GRC(I1,J1)): SpinZ, SpinXY and Den coincide with the formulas above; Green is
// N_SUN sum_nf GRC(I1,J1,nf) and the SpinZ background GRC(I1,I1,2) - GRC(I1,I1,1) has the opposite sign.
template <typename T, int EQ>
__global__ void __launch_bounds__(256) k_obs_tau(const T* __restrict__ GT0, const T* __restrict__ G0T, const T* __restrict__ G00, const T* __restrict__ GTT,
                                                 long sM, int N, int F, int n_sun, const cplx* __restrict__ phase, LattDev lt, int nt, int ntau,
                                                 double* __restrict__ acc, double* __restrict__ bg, double* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_unit = lt.n_unit, norb = lt.norb, norb2 = norb * norb;
  double* bins = reinterpret_cast<double*>(smem_raw);                 // [ch][no][imj] (re, im)
  T* dTT = reinterpret_cast<T*>(bins + (size_t)2 * OBST_NCH * norb2 * n_unit);     // [f][N]
  T* d00 = dTT + (size_t)2 * N;
  typedef T Tile[32][33];
  Tile* tA = reinterpret_cast<Tile*>(d00 + (size_t)2 * N);      // [f]: GT0 tile per flavor, [j][i]
  Tile* tB = tA + 2;                                             // [f]: G0T tile per flavor as stored (rows J fastest), [i][j]
  const int c = blockIdx.x, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const long base = (long)c * F * sM;
  for (int e = tid; e < 2 * OBST_NCH * norb2 * n_unit; e += blockDim.x) bins[e] = 0.0;
  for (int e = tid; e < F * N; e += blockDim.x) { const int f = e / N, i = e % N; dTT[e] = GTT[base + f * sM + i + (long)i * N]; d00[e] = G00[base + f * sM + i + (long)i * N]; }
  __syncthreads();
  const int nt_ = (N + 31) / 32;
  for (int tile = 0; tile < nt_ * nt_; ++tile) {
    const int i0 = (tile % nt_) * 32, j0 = (tile / nt_) * 32;
    for (int f = 0; f < F && f < 2; ++f) {
      for (int r = ty; r < 32; r += 8) {
        const int i = i0 + tx, j = j0 + r;                 // GT0(i, j): i fastest
        tA[f][r][tx] = (i < N && j < N) ? GT0[base + f * sM + i + (long)j * N] : zero_<T>();
        const int jj = j0 + tx, ii = i0 + r;               // G0T(jj, ii): jj fastest
        tB[f][r][tx] = (ii < N && jj < N) ? G0T[base + f * sM + jj + (long)ii * N] : zero_<T>();
      }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + tx, j = j0 + r;
      if (i < N && j < N) {
        const int ci = lt.cell[i], cj = lt.cell[j];
        if (ci >= 0 && cj >= 0) {
          const int no = lt.orb[i] + norb * lt.orb[j];
          const int d = lt.imj[ci + (long)cj * n_unit];
          cplx gt0[2], g0t[2];
          for (int f = 0; f < F && f < 2; ++f) { const T a = tA[f][r][tx], b = tB[f][tx][r]; gt0[f] = cplx(real_(a), imag_(a)); g0t[f] = cplx(real_(b), imag_(b)); }
          cplx v[OBST_NCH];
          cplx zi = cplx(0.0, 0.0), zj = cplx(0.0, 0.0), zz = cplx(0.0, 0.0), gsum = cplx(0.0, 0.0);
          for (int f = 0; f < F && f < 2; ++f) {
            const T a = dTT[f * N + i], b = d00[f * N + j];
            zi = zi + (cplx(1.0, 0.0) - cplx(real_(a), imag_(a))); zj = zj + (cplx(1.0, 0.0) - cplx(real_(b), imag_(b)));
            zz = zz - g0t[f] * gt0[f]; gsum = gsum + gt0[f];
          }
          if (EQ) { cplx gc = cplx(0.0, 0.0); for (int f = 0; f < F && f < 2; ++f) gc = gc - g0t[f]; v[0] = gc * (double)n_sun; }
          else v[0] = gsum * (1.0 / (double)F);
          if (F >= 2) {
            const T a1 = dTT[i], a2 = dTT[N + i], b1 = d00[j], b2 = d00[N + j];
            const cplx da = cplx(real_(a1) - real_(a2), imag_(a1) - imag_(a2)), db = cplx(real_(b1) - real_(b2), imag_(b1) - imag_(b2));
            v[1] = da * db + zz;
            v[2] = cplx(0.0, 0.0) - g0t[0] * gt0[1] - g0t[1] * gt0[0];
          } else { v[1] = zz * (double)n_sun; v[2] = cplx(0.0, 0.0); }
          v[3] = (zi * (double)n_sun) * (zj * (double)n_sun) + zz * (double)n_sun;
#pragma unroll
          for (int ch = 0; ch < OBST_NCH; ++ch) {
            double* bp = bins + 2 * ((size_t)(ch * norb2 + no) * n_unit + d);
            atomicAdd(bp, v[ch].x); atomicAdd(bp + 1, v[ch].y);
          }
        }
      }
    }
    __syncthreads();
  }
  const cplx ph = phase[c]; const double zs = (ph.x >= 0.0) ? 1.0 : -1.0; const cplx zpzs = cplx(zs, zs * ph.y / ph.x);
  for (int e = tid; e < OBST_NCH * norb2 * n_unit; e += blockDim.x) {
    const int ch = e / (norb2 * n_unit), rest = e % (norb2 * n_unit);
    const cplx v = cplx(bins[2 * e], bins[2 * e + 1]) * zpzs;
    double* ap = acc + 2 * (((size_t)ch * ntau + nt) * norb2 * n_unit + rest);
    atomicAdd(ap, v.x); atomicAdd(ap + 1, v.y);
  }
  // backgrounds (Obs_Latt0) and counters
  if (tid < norb) {
    cplx bz = cplx(0.0, 0.0), bd = cplx(0.0, 0.0);
    for (int i = 0; i < N; ++i) if (lt.cell[i] >= 0 && lt.orb[i] == tid) {
      cplx zi = cplx(0.0, 0.0);
      for (int f = 0; f < F && f < 2; ++f) { const T a = dTT[f * N + i]; zi = zi + (cplx(1.0, 0.0) - cplx(real_(a), imag_(a))); }
      bd = bd + zi * (double)n_sun;
      if (F >= 2) { const T a1 = dTT[i], a2 = dTT[N + i]; const cplx dd = cplx(real_(a2) - real_(a1), imag_(a2) - imag_(a1)); bz = EQ ? bz - dd : bz + dd; }
    }
    bz = bz * zpzs; bd = bd * zpzs;
    double* b0 = bg + 2 * ((size_t)(0 * ntau + nt) * norb + tid); double* b1 = bg + 2 * ((size_t)(1 * ntau + nt) * norb + tid);
    atomicAdd(b0, bz.x); atomicAdd(b0 + 1, bz.y); atomicAdd(b1, bd.x); atomicAdd(b1 + 1, bd.y);
  }
  if (tid == 0 && nt == 0) { atomicAdd(cnt, 1.0); atomicAdd(cnt + 1, zs); }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_obs_tau_diag: the same accumulation for lattices whose site numbering is invariant under a shift by 32 sites (imj(i + 32, j + 32) = imj(i, j),
// same orbitals -- e.g. the 16 x 16 square lattice, where 32 sites are two lattice rows; checked on the host, LattDev::shift32).  k_obs_tau spends
// its time in eight shared-memory FP64 atomics per site pair (0.69 ms per time point against an HBM floor of 0.05 ms).  Here a CTA (chain, Delta)
// walks the N/32 tiles (I, J) = (J + Delta, J) of one block diagonal: a thread's site pair keeps its lattice displacement d from tile to tile, so the
// four channels accumulate in REGISTERS over the whole diagonal and go to the shared-memory bins once.
// ------------------------------------------------------------------------------------------------------------------------
template <typename T, int EQ>
__global__ void __launch_bounds__(256, 2) k_obs_tau_diag(const T* __restrict__ GT0, const T* __restrict__ G0T, const T* __restrict__ G00, const T* __restrict__ GTT,
                                                      long sM, int N, int F, int n_sun, const cplx* __restrict__ phase, LattDev lt, int nt, int ntau,
                                                      double* __restrict__ acc, double* __restrict__ bg, double* __restrict__ cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_unit = lt.n_unit, norb = lt.norb, norb2 = norb * norb;
  double* bins = reinterpret_cast<double*>(smem_raw);                 // [ch][no][imj] (re, im)
  T* dTT = reinterpret_cast<T*>(bins + (size_t)2 * OBST_NCH * norb2 * n_unit);     // [f][N]
  T* d00 = dTT + (size_t)2 * N;
  typedef T Tile[32][33];
  Tile* tB = reinterpret_cast<Tile*>(d00 + (size_t)2 * N);      // [f]: G0T tile as stored (rows J fastest), [i][j]
  const int c = blockIdx.y, delta = blockIdx.x, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5, ntl = N >> 5;
  const long base = (long)c * F * sM;
  for (int e = tid; e < 2 * OBST_NCH * norb2 * n_unit; e += blockDim.x) bins[e] = 0.0;
  for (int e = tid; e < F * N; e += blockDim.x) { const int f = e / N, i = e % N; dTT[e] = GTT[base + f * sM + i + (long)i * N]; d00[e] = G00[base + f * sM + i + (long)i * N]; }
  __syncthreads();
  cplx v[4][OBST_NCH];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int ch = 0; ch < OBST_NCH; ++ch) v[q][ch] = cplx(0.0, 0.0);
  // both operands of tile k + 1 are requested (into registers) before tile k is worked on, so a step costs one memory latency at most
  const int nfl = F < 2 ? F : 2;
  T pg0t[2][4], pgt0[2][4];           // [f][q]: G0T(j0 + tx, i0 + ty + 8 q) and GT0(i0 + tx, j0 + ty + 8 q) of the NEXT tile
  auto fetch = [&](int k) {
    const int J = k, I = (k + delta) % ntl, i0 = 32 * I, j0 = 32 * J;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pg0t[f][q] = zero_<T>(); pgt0[f][q] = zero_<T>();
        if (f < nfl) { pg0t[f][q] = G0T[base + f * sM + (j0 + tx) + (long)(i0 + ty + 8 * q) * N]; pgt0[f][q] = GT0[base + f * sM + (i0 + tx) + (long)(j0 + ty + 8 * q) * N]; }
      }
  };
  fetch(0);
  for (int k = 0; k < ntl; ++k) {
    const int J = k, I = (k + delta) % ntl, i0 = 32 * I, j0 = 32 * J;
    T cgt0[2][4];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int q = 0; q < 4; ++q) { if (f < nfl) tB[f][ty + 8 * q][tx] = pg0t[f][q]; cgt0[f][q] = pgt0[f][q]; }      // G0T(jj, ii): jj fastest
    __syncthreads();
    if (k + 1 < ntl) fetch(k + 1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty + 8 * q, i = i0 + tx, j = j0 + r;
      cplx gt0[2], g0t[2];
      for (int f = 0; f < F && f < 2; ++f) { const T a = cgt0[f][q], b2 = tB[f][tx][r]; gt0[f] = cplx(real_(a), imag_(a)); g0t[f] = cplx(real_(b2), imag_(b2)); }
      cplx zi = cplx(0.0, 0.0), zj = cplx(0.0, 0.0), zz = cplx(0.0, 0.0), gsum = cplx(0.0, 0.0);
      for (int f = 0; f < F && f < 2; ++f) {
        const T a = dTT[f * N + i], b2 = d00[f * N + j];
        zi = zi + (cplx(1.0, 0.0) - cplx(real_(a), imag_(a))); zj = zj + (cplx(1.0, 0.0) - cplx(real_(b2), imag_(b2)));
        zz = zz - g0t[f] * gt0[f]; gsum = gsum + gt0[f];
      }
      if (EQ) { cplx gc = cplx(0.0, 0.0); for (int f = 0; f < F && f < 2; ++f) gc = gc - g0t[f]; v[q][0] = v[q][0] + gc * (double)n_sun; }
      else v[q][0] = v[q][0] + gsum * (1.0 / (double)F);
      if (F >= 2) {
        const T a1 = dTT[i], a2 = dTT[N + i], b1 = d00[j], b2 = d00[N + j];
        const cplx da = cplx(real_(a1) - real_(a2), imag_(a1) - imag_(a2)), db = cplx(real_(b1) - real_(b2), imag_(b1) - imag_(b2));
        v[q][1] = v[q][1] + (da * db + zz);
        v[q][2] = v[q][2] - g0t[0] * gt0[1] - g0t[1] * gt0[0];
      } else v[q][1] = v[q][1] + zz * (double)n_sun;
      v[q][3] = v[q][3] + ((zi * (double)n_sun) * (zj * (double)n_sun) + zz * (double)n_sun);
    }
    __syncthreads();
  }
  // a thread's four site pairs (tile-independent displacement and orbital pair): registers -> bins
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = 32 * delta + tx, j = ty + 8 * q;          // representative pair of the diagonal: tile (Delta, 0)
    const int ci = lt.cell[i % N], cj = lt.cell[j];
    if (ci >= 0 && cj >= 0) {
      const int no = lt.orb[i % N] + norb * lt.orb[j]; const int d = lt.imj[ci + (long)cj * n_unit];
#pragma unroll
      for (int ch = 0; ch < OBST_NCH; ++ch) { double* bp = bins + 2 * ((size_t)(ch * norb2 + no) * n_unit + d); atomicAdd(bp, v[q][ch].x); atomicAdd(bp + 1, v[q][ch].y); }
    }
  }
  __syncthreads();
  const cplx ph = phase[c]; const double zs = (ph.x >= 0.0) ? 1.0 : -1.0; const cplx zpzs = cplx(zs, zs * ph.y / ph.x);
  for (int e = tid; e < OBST_NCH * norb2 * n_unit; e += blockDim.x) {
    if (bins[2 * e] == 0.0 && bins[2 * e + 1] == 0.0) continue;      // displacements this diagonal does not contain
    const int ch = e / (norb2 * n_unit), rest = e % (norb2 * n_unit);
    const cplx vv = cplx(bins[2 * e], bins[2 * e + 1]) * zpzs;
    double* ap = acc + 2 * (((size_t)ch * ntau + nt) * norb2 * n_unit + rest);
    atomicAdd(ap, vv.x); atomicAdd(ap + 1, vv.y);
  }
  if (delta == 0) {      // backgrounds (Obs_Latt0) and counters: once per chain
    if (tid < norb) {
      cplx bz = cplx(0.0, 0.0), bd = cplx(0.0, 0.0);
      for (int i = 0; i < N; ++i) if (lt.cell[i] >= 0 && lt.orb[i] == tid) {
        cplx zi = cplx(0.0, 0.0);
        for (int f = 0; f < F && f < 2; ++f) { const T a = dTT[f * N + i]; zi = zi + (cplx(1.0, 0.0) - cplx(real_(a), imag_(a))); }
        bd = bd + zi * (double)n_sun;
        if (F >= 2) { const T a1 = dTT[i], a2 = dTT[N + i]; const cplx dd = cplx(real_(a2) - real_(a1), imag_(a2) - imag_(a1)); bz = EQ ? bz - dd : bz + dd; }
      }
      bz = bz * zpzs; bd = bd * zpzs;
      double* b0 = bg + 2 * ((size_t)(0 * ntau + nt) * norb + tid); double* b1 = bg + 2 * ((size_t)(1 * ntau + nt) * norb + tid);
      atomicAdd(b0, bz.x); atomicAdd(b0 + 1, bz.y); atomicAdd(b1, bd.x); atomicAdd(b1 + 1, bd.y);
    }
    if (tid == 0 && nt == 0) { atomicAdd(cnt, 1.0); atomicAdd(cnt + 1, zs); }
  }
}

template <typename T>
static size_t obs_tau_smem(int N, int n_unit, int norb) {
  return sizeof(double) * 2 * OBST_NCH * (size_t)norb * norb * n_unit + sizeof(T) * ((size_t)4 * N + (size_t)4 * 32 * 33) + 64;
}
