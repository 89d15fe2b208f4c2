// tfg_math.cuh -- lean double-precision elementary functions for the FAST arithmetic mode.
//
// Why not libdevice: in the melt kernel ~70 % of all issued instructions were not FP64 math but the
// 2 x UMOV immediates libdevice spends per polynomial coefficient, plus branches around special cases
// (ncu, profiles/).  These versions keep coefficients in __constant__ memory (one LDCU.128 feeds two
// DFMAs), are branch-free on their fast path, use a MUFU seed + two Newton steps for reciprocals, and are
// accurate to <= ~2 ulp on the argument ranges the physics produces.  Anything outside the fast path
// (non-finite, huge, non-positive log argument ...) falls back to libdevice, so semantics are preserved.
//
// The strict mode never uses this file: it stays on libdevice + single-rounding IEEE operations.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TFG_HD __host__ __device__ __forceinline__
#else
#define TFG_HD inline
#define __constant__ static const
#endif

namespace tfg {
namespace fm {

#include "tfg_math_coeffs.inc"

#if defined(__CUDA_ARCH__)
#define TFG_COEF(tab, i) tab[i]
__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }
__device__ __forceinline__ int lo32(double x) { return __double2loint(x); }
__device__ __forceinline__ double mk64(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
#define TFG_COEF(tab, i) tab[i]
inline int hi32(double x) { long long b; __builtin_memcpy(&b, &x, 8); return (int)(b >> 32); }
inline int lo32(double x) { long long b; __builtin_memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
inline double mk64(int hi, int lo) {
  long long b = ((long long)hi << 32) | (unsigned int)lo; double x; __builtin_memcpy(&x, &b, 8); return x;
}
#endif

template <int N>
TFG_HD double horner(const double (&c)[N], double x) {
  double p = c[0];
#pragma unroll
  for (int i = 1; i < N; ++i) p = fma(p, x, c[i]);
  return p;
}

// The same polynomial evaluated as K interleaved Horner chains in x^K (K = 2 or 4): K independent dependency
// chains of length ~N/K instead of one of length N.  The melt kernel runs ~3 warps per scheduler and is bound by
// DFMA latency, not throughput (ncu: "wait" is the dominant stall), so instruction-level parallelism pays.
template <int K, int N>
TFG_HD double horner_k(const double (&c)[N], double x) {
  static_assert(K == 2 || K == 4, "K");
  const double x2 = x * x;
  const double y = (K == 2) ? x2 : x2 * x2;
  double S[K];
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const int jtop = ((N - 1 - r) / K) * K + r;  // highest power congruent to r (mod K); coefficient of x^j is c[N-1-j]
    double acc = c[N - 1 - jtop];
#pragma unroll
    for (int j = jtop - K; j >= 0; j -= K) acc = fma(acc, y, c[N - 1 - j]);
    S[r] = acc;
  }
  if (K == 2) return fma(x, S[1], S[0]);
  return fma(x2, fma(x, S[3], S[2]), fma(x, S[1], S[0]));
}

// 1/b for finite, normal, non-zero b: MUFU.RCP64H seed (~2^-23) + two Newton steps.
TFG_HD double rcp(double b) {
  double r;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
  r = (double)(1.0f / (float)b);
#endif
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// a/b with one residual correction (<= ~0.6 ulp); b finite, normal, non-zero
TFG_HD double div(double a, double b) {
  const double r = rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// sqrt(w) for w >= 0 (w below 1e-290 is treated as 1e-290): MUFU.RSQ64H seed + Newton
TFG_HD double sqrt_pos(double w) {
  w = (w > 1e-290) ? w : 1e-290;
  double y;
#if defined(__CUDA_ARCH__)
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
#else
  y = (double)(1.0f / sqrtf((float)w));
#endif
  double h = 0.5 * w;
  y = y * fma(-h * y, y, 1.5);   // two Newton steps on 1/sqrt
  y = y * fma(-h * y, y, 1.5);
  double s = w * y;
  return fma(fma(-s, s, w), 0.5 * y, s);  // Heron correction
}

// exp(x), |x| < 700
TFG_HD double exp_core(double x) {
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);  // low word = rint(x*log2(e))
  const int n = lo32(t);
  const double fn = t - 6755399441055744.0;
  double r = fma(fn, -6.93147180369123816490e-01, x);
  r = fma(fn, -1.90821492927058770002e-10, r);
  const double q = horner_k<2>(kExpQ, r);
  const double p = fma(q, r * r, r) + 1.0;  // in [0.70, 1.42]
  return mk64(hi32(p) + (n << 20), lo32(p));
}
TFG_HD double exp_f(double x) { return (fabs(x) < 700.0) ? exp_core(x) : exp(x); }

// log(x) for positive, normal, finite x
TFG_HD double log_core(double x) {
  int hi = hi32(x);
  int e = (hi >> 20) - 1023;
  double m = mk64((hi & 0x000fffff) | 0x3ff00000, lo32(x));  // [1, 2)
  const bool up = m > 1.4142135623730951;
  m = up ? 0.5 * m : m;                                        // [0.707, 1.414]
  e = up ? e + 1 : e;
  const double f = m - 1.0;                                    // exact
  const double s = div(f, 2.0 + f);
  const double z = s * s;
  const double lm = fma(s * z, horner_k<2>(kLogP, z), 2.0 * s);     // log(m) = 2 atanh(s)
  const double ed = (double)e;
  return fma(ed, 6.93147180369123816490e-01, fma(ed, 1.90821492927058770002e-10, lm));
}
TFG_HD double log_f(double x) {
  const unsigned hi = (unsigned)hi32(x);
  return (hi - 0x00100000u < 0x7fe00000u) ? log_core(x) : log(x);
}

// x**y for x > 0 (normal, finite): exp(y*log(x)); relative error ~ (|y*log x| + 2) ulp
TFG_HD double pow_f(double x, double y) { return exp_f(y * log_f(x)); }

// asin(x) for 0 <= x <= 1 (values slightly above 1 are clamped)
TFG_HD double asin01(double x) {
  const bool big = x > 0.5;
  const double w = big ? 0.5 * (1.0 - fmin(x, 1.0)) : x * x;
  const double s = big ? sqrt_pos(w) : x;
  const double a = fma(s * w, horner_k<2>(kAsinP, w), s);
  // pi/2 = 1.5707963267948966 + 6.123233995736766e-17
  return big ? (fma(-2.0, a, 1.5707963267948966) + 6.123233995736766e-17) : a;
}

// atan(x), |x| < 1e150
TFG_HD double atan_core(double x) {
  const double ax = fabs(x);
  const bool inv = ax > 1.0;
  const double t = inv ? rcp(ax) : ax;
  const double a0 = t * horner_k<4>(kAtanP, t * t);
  const double a = inv ? ((1.5707963267948966 - a0) + 6.123233995736766e-17) : a0;
  return copysign(a, x);
}
TFG_HD double atan_f(double x) { return (fabs(x) < 1e150) ? atan_core(x) : atan(x); }

// Stull (2011) wet-bulb temperature with RH as the reference feeds it (a fraction, reference
// bmi_topoflow_glacier.py:1514-1520), for 0 <= RH <= 2:
//   T*atan(0.151977*sqrt(RH+8.313659)) + atan(T+RH) - atan(RH-1.676331) + 0.00391838*RH^1.5*atan(0.023101*RH) - 4.86035
// The first arctangent is a smooth function of RH on [0, 2] (direct polynomial, no sqrt), the last has a tiny
// argument (4-term series); only the two middle ones need the general routine.
TFG_HD double stull_wet_bulb(double T, double RH) {
  const double a1 = horner_k<2>(kStull1, RH);
  const double u = 0.023101 * RH, u2 = u * u;
  const double a4 = u * fma(u2, fma(u2, fma(u2, fma(u2, 1.0 / 9.0, -1.0 / 7.0), 0.2), -1.0 / 3.0), 1.0);
  const double t4 = (0.00391838 * (RH * sqrt_pos(RH))) * a4;
  return ((((T * a1) + atan_core(T + RH)) - atan_core(RH - 1.676331)) + t4) - 4.86035;
}

// ---- float32 counterparts (coefficients are FFMA immediates; MUFU reciprocal / rsqrt) ------------------------
template <class T>
TFG_HD float horner32(float x) {
  float p = T::c(0);
#pragma unroll
  for (int i = 1; i < T::N; ++i) p = fmaf(p, x, T::c(i));
  return p;
}
TFG_HD float rcp32(float b) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(b);
#else
  return 1.0f / b;
#endif
}
TFG_HD float atan32(float x) {  // |x| < 1e30
  const float ax = fabsf(x);
  const bool inv = ax > 1.0f;
  const float t = inv ? rcp32(ax) : ax;
  const float a0 = t * horner32<kAtanP32>(t * t);
  return copysignf(inv ? 1.5707963267948966f - a0 : a0, x);
}
TFG_HD float asin01_32(float x) {  // 0 <= x <= 1 (clamped)
  const bool big = x > 0.5f;
  const float w = big ? 0.5f * (1.0f - fminf(x, 1.0f)) : x * x;
  const float s = big ? sqrtf(w) : x;
  const float a = fmaf(s * w, horner32<kAsinP32>(w), s);
  return big ? fmaf(-2.0f, a, 1.5707963267948966f) : a;
}
TFG_HD float stull_wet_bulb32(float T, float RH) {  // 0 <= RH <= 2
  const float a1 = horner32<kStull132>(RH);
  const float u = 0.023101f * RH;
  const float a4 = u * fmaf(u * u, -1.0f / 3.0f, 1.0f);
  const float t4 = (0.00391838f * (RH * sqrtf(RH))) * a4;
  return ((((T * a1) + atan32(T + RH)) - atan32(RH - 1.676331f)) + t4) - 4.86035f;
}

}  // namespace fm
}  // namespace tfg
