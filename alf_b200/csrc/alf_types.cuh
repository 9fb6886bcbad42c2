// Scalar helpers shared by all kernels: the sweep is templated on T = double (real models,
// e.g. Mz-Hubbard) or T = cplx (SU(N) Hubbard, Kondo).  ALF itself always computes in
// complex(kind(0.d0)) (Prog/main.F90:152); the real instantiation is this build's choice when
// every Op_T / Op_V table is real (Prog/Operator_mod.F90:1103-1120 Op_is_real precedent).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#define ALF_HD __host__ __device__ __forceinline__
// cap for cudaFuncAttributeMaxDynamicSharedMemorySize (227 KB opt-in maximum of sm_100): the attribute is per-function state shared by
// all handles of the process, so it is always raised to the maximum instead of the size one particular launch needs
#define ALF_MAX_DYN_SMEM (227 * 1024)

struct __align__(16) cplx {
  double x, y;
  ALF_HD cplx() {}
  ALF_HD cplx(double a) : x(a), y(0.0) {}
  ALF_HD cplx(double a, double b) : x(a), y(b) {}
};
ALF_HD cplx operator+(cplx a, cplx b) { return cplx(a.x + b.x, a.y + b.y); }
ALF_HD cplx operator-(cplx a, cplx b) { return cplx(a.x - b.x, a.y - b.y); }
ALF_HD cplx operator-(cplx a) { return cplx(-a.x, -a.y); }
ALF_HD cplx operator*(cplx a, cplx b) { return cplx(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
ALF_HD cplx operator*(double a, cplx b) { return cplx(a * b.x, a * b.y); }
ALF_HD cplx operator*(cplx b, double a) { return cplx(a * b.x, a * b.y); }
ALF_HD cplx operator/(cplx a, cplx b) {
  double d = b.x * b.x + b.y * b.y;
  return cplx((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
ALF_HD cplx operator/(cplx a, double b) { return cplx(a.x / b, a.y / b); }
ALF_HD cplx& operator+=(cplx& a, cplx b) { a.x += b.x; a.y += b.y; return a; }
ALF_HD cplx& operator-=(cplx& a, cplx b) { a.x -= b.x; a.y -= b.y; return a; }
ALF_HD cplx& operator*=(cplx& a, cplx b) { a = a * b; return a; }
ALF_HD cplx& operator*=(cplx& a, double b) { a.x *= b; a.y *= b; return a; }

ALF_HD double conj_(double a) { return a; }
ALF_HD cplx conj_(cplx a) { return cplx(a.x, -a.y); }
ALF_HD double real_(double a) { return a; }
ALF_HD double real_(cplx a) { return a.x; }
ALF_HD double imag_(double) { return 0.0; }
ALF_HD double imag_(cplx a) { return a.y; }
ALF_HD double abs2_(double a) { return a * a; }
ALF_HD double abs2_(cplx a) { return a.x * a.x + a.y * a.y; }
ALF_HD double abs_(double a) { return fabs(a); }
ALF_HD double abs_(cplx a) { return sqrt(a.x * a.x + a.y * a.y); }
// fused multiply-add  c += a*b  and  c += conj(a)*b
ALF_HD void fma_(double& c, double a, double b) { c = fma(a, b, c); }
ALF_HD void fma_(cplx& c, cplx a, cplx b) {
  c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
ALF_HD void fmac_(double& c, double a, double b) { c = fma(a, b, c); }
ALF_HD void fmac_(cplx& c, cplx a, cplx b) {  // c += conj(a)*b
  c.x = fma(a.x, b.x, c.x); c.x = fma(a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y); c.y = fma(-a.y, b.x, c.y);
}
template <typename T> ALF_HD T make_(double re, double im);
template <> ALF_HD double make_<double>(double re, double) { return re; }
template <> ALF_HD cplx make_<cplx>(double re, double im) { return cplx(re, im); }
ALF_HD double exp_(double a) { return exp(a); }
ALF_HD cplx exp_(cplx a) { const double e = exp(a.x); double sn, cs; sincos(a.y, &sn, &cs); return cplx(e * cs, e * sn); }
template <typename T> ALF_HD T zero_() { return make_<T>(0.0, 0.0); }
template <typename T> ALF_HD T one_() { return make_<T>(1.0, 0.0); }
ALF_HD bool isnan_(double a) { return a != a; }
ALF_HD bool isnan_(cplx a) { return a.x != a.x || a.y != a.y; }

#ifdef __CUDACC__
__device__ __forceinline__ double shfl_xor_(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_xor_(cplx v, int m) {
  return cplx(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ double shfl_(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ cplx shfl_(cplx v, int l) {
  return cplx(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l));
}
// FP64 tensor-core tile product D(8x8) += A(8x4, row) B(4x8, col).  Lane l (g = l >> 2, q = l & 3) holds A[g][q], B[q][g] and
// C[g][2q], C[g][2q+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__host__ __device__ inline int ld_pad(int rows) { int l = rows; while (l % 16 != 4) ++l; return l; }   // conflict-free DMMA fragment loads
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = v + shfl_xor_(v, m);
  return v;
}
#endif

// complex scalar carried per chain (Phase) regardless of T
typedef cplx phase_t;
