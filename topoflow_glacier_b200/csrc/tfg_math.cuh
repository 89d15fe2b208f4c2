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
#define TFG_HD __device__ __forceinline__
#define TFG_DEVTAB static __device__ const   // piecewise-polynomial tables in global memory, read through L1 (__ldg)
#else
#define TFG_HD inline
#define __constant__ static const
#define TFG_DEVTAB static const
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

// Literals of the routines below whose low word is non-zero (such a float64 immediate costs two UMOV / IMAD.MOV
// per use; from the constant bank it is half an LDCU).  Literals like 1.0, 0.5 or 2^52*1.5 encode in the
// instruction and stay inline.
struct MathLit {
  double exp_scale, exp_nl2h, exp_nl2l, ln2h, ln2l, pio2h, pio2l, pih, pil, tiny, st_u, st_c9, st_c7, st_c5, st_c3,
      st_d, st_c, st_f, r7_c1, r7_c2;
};
__constant__ MathLit kML = {92.332482616893656877, -1.08304246932675596327e-02, -2.98158582698529328128e-12,
                            6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.5707963267948966,
                            6.123233995736766e-17, 3.141592653589793, 1.2246467991473532e-16, 1e-290, 0.023101,
                            1.0 / 9.0, -1.0 / 7.0, 0.2, -1.0 / 3.0, 0.00391838, 1.676331, 4.86035, 1.0 / 7.0,
                            4.0 / 49.0};

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

#if !defined(__CUDA_ARCH__)
// host stand-in for MUFU.RCP64H: reciprocal of the HIGH WORD of b (relative error up to 2^-20), low word cleared
inline double rcp_seed_host(double b) { const double r = 1.0 / mk64(hi32(b), 0); return mk64(hi32(r), 0); }
#endif
// 1/b for finite, normal, non-zero b: MUFU.RCP64H seed (~2^-20) + two Newton steps.
TFG_HD double rcp(double b) {
  double r;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
  r = rcp_seed_host(b);
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
// sqrt(w) for w >= 0 (w below 1e-290 is treated as 1e-290): MUFU.RSQ64H seed (~2^-20), one Newton step on
// 1/sqrt (-> 2^-39), Heron correction (-> 2^-77)
TFG_HD double sqrt_pos(double w) {
  w = (w > kML.tiny) ? w : kML.tiny;
  double y;
#if defined(__CUDA_ARCH__)
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
#else
  { const double q = 1.0 / sqrt(mk64(hi32(w), 0)); y = mk64(hi32(q), 0); }  // stand-in for MUFU.RSQ64H
#endif
  const double h = 0.5 * w;
  y = y * fma(-h * y, y, 1.5);
  const double s = w * y;
  return fma(fma(-s, s, w), 0.5 * y, s);
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

// ---- table-driven exp / log --------------------------------------------------------------------------------
// One lookup (64 x 8 B resp. 128 x 16 B, kept in shared memory by the melt kernel) shrinks the reduced argument
// to |r| <= ln2/128 resp. 1/128, so a degree-5/7 polynomial replaces the degree-11/15 ones above and the log
// needs no division: 10 resp. 13 FP64 instructions instead of 16 resp. 25, <= ~1 ulp.
#if defined(__CUDACC__)
extern __shared__ double tfg_tabs[];  // dynamic shared memory of the melt kernel: [0,64) 2^(j/64), [64,320) {1/c_i, log c_i}
#endif
#if defined(__CUDA_ARCH__)
#define TFG_EXPTAB(j) tfg_tabs[j]
#define TFG_LOGTAB(i, c) tfg_tabs[64 + 2 * (i) + (c)]
#else
#define TFG_EXPTAB(j) kExpTab[j]
#define TFG_LOGTAB(i, c) kLogTab[i][c]
#endif
constexpr int kTabDoubles = 64 + 2 * 128;

// 1/b, one cubic step from the MUFU.RCP64H seed (relative error e0 <= 2^-20  ->  e0^3)
TFG_HD double rcp3(double b) {
  double r;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
  r = rcp_seed_host(b);
#endif
  const double e = fma(-b, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// a/b without the final residual correction (<= ~1.5 ulp)
TFG_HD double div_fast(double a, double b) { return a * rcp3(b); }

// exp(x), |x| < 700
TFG_HD double exp_tab(double x) {
  const double t = fma(x, kML.exp_scale, 6755399441055744.0);  // low word = rint(x * 64/ln2)
  const int k = lo32(t);
  const double fn = t - 6755399441055744.0;
  double r = fma(fn, kML.exp_nl2h, x);   // -ln2hi/64 (32 significant bits)
  r = fma(fn, kML.exp_nl2l, r);          // -ln2lo/64
  const double T = TFG_EXPTAB(k & 63);
  const double r2 = r * r;
  const double a = fma(r, TFG_COEF(kExpTQ, 2), TFG_COEF(kExpTQ, 3));
  const double b = fma(r, TFG_COEF(kExpTQ, 0), TFG_COEF(kExpTQ, 1));
  const double p = fma(r2, fma(r2, b, a), r);           // exp(r) - 1
  const double y = fma(T, p, T);                        // in [1, 2.01)
  return mk64(hi32(y) + ((k >> 6) << 20), lo32(y));
}

// log(x) for positive, normal, finite x
TFG_HD double log_tab(double x) {
  const int hi = hi32(x);
  const int tmp = hi - 0x3fe60000;                      // bits(x) - bits(0.6875): z = x / 2^e in [0.6875, 1.375)
  const int i = (tmp >> 13) & 127;
  const int e = tmp >> 20;
  const double z = mk64(hi - (tmp & (int)0xfff00000), lo32(x));
  const double invc = TFG_LOGTAB(i, 0), logc = TFG_LOGTAB(i, 1);
  const double r = fma(z, invc, -1.0);
  const double ed = (double)e;
  const double w = fma(ed, kML.ln2h, logc);
  const double r2 = r * r;
  const double a0 = fma(r, TFG_COEF(kLogTA, 4), TFG_COEF(kLogTA, 5));
  const double a1 = fma(r, TFG_COEF(kLogTA, 2), TFG_COEF(kLogTA, 3));
  const double a2 = fma(r, TFG_COEF(kLogTA, 0), TFG_COEF(kLogTA, 1));
  const double pa = fma(r2, fma(r2, a2, a1), a0);
  const double lo = fma(r2, pa, fma(ed, kML.ln2l, r));
  return w + lo;
}

// ---- N independent arguments at once -----------------------------------------------------------------------
// ptxas schedules the melt kernel for register pressure (96 registers), which serialises the elementary functions:
// a warp then issues one dependent FP64 chain at a time and waits out the pipe latency between instructions
// (ncu: `wait` is the top stall).  These variants evaluate N independent arguments stage by stage, so the N chains
// are adjacent in program order and overlap in the FP64 pipe.
template <int N>
TFG_HD void rcp3_n(const double (&b)[N], double (&r)[N]) {
  double e[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(b[i]));
#else
    r[i] = rcp_seed_host(b[i]);
#endif
  }
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(-b[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < N; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(r[i], e[i], r[i]);
}

template <int N>
TFG_HD void exp_tab_n(const double (&x)[N], double (&y)[N]) {
  double t[N], fn[N], r[N], T[N], r2[N], a[N], b[N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = fma(x[i], kML.exp_scale, 6755399441055744.0);
#pragma unroll
  for (int i = 0; i < N; ++i) { k[i] = lo32(t[i]); fn[i] = t[i] - 6755399441055744.0; }
#pragma unroll
  for (int i = 0; i < N; ++i) { r[i] = fma(fn[i], kML.exp_nl2h, x[i]); T[i] = TFG_EXPTAB(k[i] & 63); }
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = fma(fn[i], kML.exp_nl2l, r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    r2[i] = r[i] * r[i];
    a[i] = fma(r[i], TFG_COEF(kExpTQ, 2), TFG_COEF(kExpTQ, 3));
    b[i] = fma(r[i], TFG_COEF(kExpTQ, 0), TFG_COEF(kExpTQ, 1));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = fma(r2[i], b[i], a[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) a[i] = fma(r2[i], a[i], r[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double v = fma(T[i], a[i], T[i]);
    y[i] = mk64(hi32(v) + ((k[i] >> 6) << 20), lo32(v));
  }
}

template <int N>
TFG_HD void log_tab_n(const double (&x)[N], double (&y)[N]) {
  double z[N], invc[N], logc[N], r[N], ed[N], w[N], r2[N], a0[N], a1[N], a2[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int hi = hi32(x[i]);
    const int tmp = hi - 0x3fe60000;
    const int j = (tmp >> 13) & 127;
    z[i] = mk64(hi - (tmp & (int)0xfff00000), lo32(x[i]));
    invc[i] = TFG_LOGTAB(j, 0);
    logc[i] = TFG_LOGTAB(j, 1);
    ed[i] = (double)(tmp >> 20);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) { r[i] = fma(z[i], invc[i], -1.0); w[i] = fma(ed[i], kML.ln2h, logc[i]); }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    r2[i] = r[i] * r[i];
    a0[i] = fma(r[i], TFG_COEF(kLogTA, 4), TFG_COEF(kLogTA, 5));
    a1[i] = fma(r[i], TFG_COEF(kLogTA, 2), TFG_COEF(kLogTA, 3));
    a2[i] = fma(r[i], TFG_COEF(kLogTA, 0), TFG_COEF(kLogTA, 1));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) { a1[i] = fma(r2[i], a2[i], a1[i]); ed[i] = fma(ed[i], kML.ln2l, r[i]); }
#pragma unroll
  for (int i = 0; i < N; ++i) a0[i] = fma(r2[i], a1[i], a0[i]);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = w[i] + fma(r2[i], a0[i], ed[i]);
}

// x**(1/7) for 1e-30 < x < 1e30 (the Brutsaert emissivity, reference bmi_topoflow_glacier.py:1179): a float32
// MUFU seed of w = x**(-1/7) (relative error d <= ~1e-6), ONE third-order Householder step on f(w) = w**-7 - x
// (no division: e = 1 - x w**7, w <- w (1 + e/7 + 4 e*e/49), error 20 d**3), then x**(1/7) = x w**6.
// 12 FP64 instructions instead of the 24 of exp(log(x)/7).  `seed_scale` perturbs the seed (host tests only).
TFG_HD double root7(double x, float seed_scale = 1.0f) {
  float w0f;
#if defined(__CUDA_ARCH__)
  asm("{ .reg .f32 l; lg2.approx.ftz.f32 l, %1; mul.f32 l, l, 0fBE124925; ex2.approx.ftz.f32 %0, l; }"
      : "=f"(w0f) : "f"((float)x));   // 0fBE124925 = -1/7
  w0f *= seed_scale;
#else
  w0f = exp2f(log2f((float)x) * (-1.0f / 7.0f)) * seed_scale;
#endif
  const double w0 = (double)w0f;
  const double w2 = w0 * w0, w4 = w2 * w2;
  const double w7 = (w4 * w2) * w0;
  const double e = fma(-x, w7, 1.0);
  const double w1 = fma(w0, e * fma(e, kML.r7_c2, kML.r7_c1), w0);
  const double v2 = w1 * w1, v4 = v2 * v2;
  return x * (v4 * v2);
}

// exp_tab_n<2> and root7 in one pass, stage by stage: the 7th root is a serial chain of twelve multiplications behind two
// MUFU round trips; evaluated beside the two exponentials its latency is filled with their work.  Same operations per
// value as the separate routines (bit-identical results).
TFG_HD void exp_tab2_root7(const double (&x)[2], double (&y)[2], double xr, double& r7) {
  float w0f;
#if defined(__CUDA_ARCH__)
  asm("{ .reg .f32 l; lg2.approx.ftz.f32 l, %1; mul.f32 l, l, 0fBE124925; ex2.approx.ftz.f32 %0, l; }" : "=f"(w0f) : "f"((float)xr));
#else
  w0f = exp2f(log2f((float)xr) * (-1.0f / 7.0f));
#endif
  double t[2], fn[2], r[2], T[2], r2[2], a[2], b[2];
  int k[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) t[i] = fma(x[i], kML.exp_scale, 6755399441055744.0);
  const double w0 = (double)w0f;
#pragma unroll
  for (int i = 0; i < 2; ++i) { k[i] = lo32(t[i]); fn[i] = t[i] - 6755399441055744.0; }
  const double w2 = w0 * w0;
#pragma unroll
  for (int i = 0; i < 2; ++i) { r[i] = fma(fn[i], kML.exp_nl2h, x[i]); T[i] = TFG_EXPTAB(k[i] & 63); }
  const double w4 = w2 * w2;
#pragma unroll
  for (int i = 0; i < 2; ++i) r[i] = fma(fn[i], kML.exp_nl2l, r[i]);
  const double w6 = w4 * w2;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    r2[i] = r[i] * r[i];
    a[i] = fma(r[i], TFG_COEF(kExpTQ, 2), TFG_COEF(kExpTQ, 3));
    b[i] = fma(r[i], TFG_COEF(kExpTQ, 0), TFG_COEF(kExpTQ, 1));
  }
  const double w7 = w6 * w0;
  const double e = fma(-xr, w7, 1.0);
#pragma unroll
  for (int i = 0; i < 2; ++i) a[i] = fma(r2[i], b[i], a[i]);
  const double u = e * fma(e, kML.r7_c2, kML.r7_c1);
#pragma unroll
  for (int i = 0; i < 2; ++i) a[i] = fma(r2[i], a[i], r[i]);
  const double w1 = fma(w0, u, w0);
  const double v2 = w1 * w1;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double v = fma(T[i], a[i], T[i]);
    y[i] = mk64(hi32(v) + ((k[i] >> 6) << 20), lo32(v));
  }
  const double v4 = v2 * v2;
  r7 = xr * (v4 * v2);
}

// asin(x) for 0 <= x <= 1 (values slightly above 1 are clamped)
TFG_HD double asin01(double x) {
  const bool big = x > 0.5;
  const double w = big ? 0.5 * (1.0 - fmin(x, 1.0)) : x * x;
  const double s = big ? sqrt_pos(w) : x;
  const double a = fma(s * w, horner_k<2>(kAsinP, w), s);
  // pi/2 = 1.5707963267948966 + 6.123233995736766e-17
  return big ? (fma(-2.0, a, kML.pio2h) + kML.pio2l) : a;
}

// atan(x), |x| < 1e150
TFG_HD double atan_core(double x) {
  const double ax = fabs(x);
  const bool inv = ax > 1.0;
  const double t = inv ? rcp(ax) : ax;
  const double a0 = t * horner_k<4>(kAtanP, t * t);
  const double a = inv ? ((kML.pio2h - a0) + kML.pio2l) : a0;
  return copysign(a, x);
}
TFG_HD double atan_f(double x) { return (fabs(x) < 1e150) ? atan_core(x) : atan(x); }

// atan(a) - atan(b) = atan2(a - b, 1 + a*b): one polynomial and one reciprocal instead of two of each.
// |a|, |b| < 1e150; absolute error <= ~2e-16 (+ 1 ulp of the quotient where a*b ~ -1, i.e. the result ~ +-pi/2).
TFG_HD double atan_diff(double a, double b) {
  const double y = a - b, x = fma(a, b, 1.0);
  const double ay = fabs(y), ax = fabs(x);
  const bool swap = ay > ax;
  const double hi = swap ? ay : ax, lo = swap ? ax : ay;
  const double t = (hi > 0.0) ? lo * rcp3(hi) : 0.0;     // in [0, 1]
  double r = t * horner_k<4>(kAtanP, t * t);
  r = swap ? ((kML.pio2h - r) + kML.pio2l) : r;
  r = (x < 0.0) ? ((kML.pih - r) + kML.pil) : r;
  return copysign(r, y);
}

// Stull (2011) wet-bulb temperature with RH as the reference feeds it (a fraction, reference
// bmi_topoflow_glacier.py:1514-1520), for 0 <= RH <= 2:
//   T*atan(0.151977*sqrt(RH+8.313659)) + atan(T+RH) - atan(RH-1.676331) + 0.00391838*RH^1.5*atan(0.023101*RH) - 4.86035
// The first arctangent is a smooth function of RH on [0, 2] (direct polynomial, no sqrt), the last has a tiny
// argument (4-term series), and the two middle ones are one atan2.
TFG_HD double stull_wet_bulb(double T, double RH) {
  const double a1 = horner_k<2>(kStull1, RH);
  const double u = kML.st_u * RH, u2 = u * u;
  const double a4 = u * fma(u2, fma(u2, fma(u2, fma(u2, kML.st_c9, kML.st_c7), kML.st_c5), kML.st_c3), 1.0);
  const double t4 = (kML.st_d * (RH * sqrt_pos(RH))) * a4;
  return (((T * a1) + atan_diff(T + RH, RH - kML.st_c)) + t4) - kML.st_f;
}

// ---- piecewise polynomials on uniform bins (tables kWetBulbTab / kAtanTab / kAirMassC / kAirMassS) --------------
// Bin k = rint(S x) is found with the 2^52 * 1.5 rounding trick (as in exp_tab); the local variable u = S x - k lies
// in [-1/2, 1/2].  A row holds the coefficients highest degree first; even and odd powers are two Horner chains.
// The tables sit in global memory and are read with 128-bit read-only loads: they are ~14 KB, indexed almost
// uniformly across a warp (cos Z) or by the few snowing lanes (wet bulb), and stay L1-resident -- the forcing stream
// uses evict-first loads.
TFG_HD void ld_pair(const double* p, double& a, double& b) {
#if defined(__CUDA_ARCH__)
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  a = t.x; b = t.y;
#else
  a = p[0]; b = p[1];
#endif
}
TFG_HD int bin_of(double x, double S, double& u) {
  const double t = fma(x, S, 6755399441055744.0);
  u = fma(x, S, -(t - 6755399441055744.0));
  return lo32(t);
}
TFG_HD double poly7_row(const double* row, double u) {
  double c0, c1, c2, c3, c4, c5, c6, c7;
  ld_pair(row, c0, c1); ld_pair(row + 2, c2, c3); ld_pair(row + 4, c4, c5); ld_pair(row + 6, c6, c7);
  const double u2 = u * u;
  const double ev = fma(fma(fma(c1, u2, c3), u2, c5), u2, c7);
  const double od = fma(fma(fma(c0, u2, c2), u2, c4), u2, c6);
  return fma(od, u, ev);
}
TFG_HD double poly5_row(const double* row, double u) {
  double c0, c1, c2, c3, c4, c5;
  ld_pair(row, c0, c1); ld_pair(row + 2, c2, c3); ld_pair(row + 4, c4, c5);
  const double u2 = u * u;
  return fma(fma(fma(c0, u2, c2), u2, c4), u, fma(fma(c1, u2, c3), u2, c5));
}

// Stull (2011) wet bulb, table-driven, for 3/64 <= RH <= 2 and |T| <= 100 (see stull_wet_bulb for the formula):
//   T_wb = T a1(RH) + atan(T + RH) + g(RH),  g = -atan(RH - 1.676331) + 0.00391838 RH^1.5 atan(0.023101 RH) - 4.86035
// a1 and g from one table row, atan from the atan table after an argument inversion: ~40 FP64 instructions instead
// of ~80 (no square root, no degree-20 arctangent).  Absolute error <= 3e-14 (tests/test_host_math.py).
TFG_HD double stull_wet_bulb_tab(double T, double RH) {
  double u;
  const int k = bin_of(RH, 16.0, u);
  const double* row = kWetBulbTab[k];
  const double a1 = poly5_row(row, u);
  const double g = poly7_row(row + 6, u);
  const double x = T + RH, ax = fabs(x);
  const bool inv = ax > 1.0;
  const double t = inv ? rcp3(ax) : ax;
  double v;
  const int k2 = bin_of(t, 16.0, v);
  double a = poly7_row(kAtanTab[k2], v);
  a = inv ? ((kML.pio2h - a) + kML.pio2l) : a;
  return fma(T, a1, g) + copysign(a, x);
}

// 1 / M for the Kasten & Young (1989) optical air mass M = 1/(sin g + 0.50572 (g + 6.07995)^-1.6364), g = asin(c) in
// degrees, c = cos Z in [0, 1] (reference solar_funcs.py:540-568): replaces asin + log + exp of the closed form by one
// table row (plus a square root above 70 deg elevation).  Relative error <= 6e-14 (5e-14 in the first bin above the
// horizon, <= 3e-15 above 3 deg elevation; tests/test_host_math.py).
TFG_HD double inv_air_mass(double c) {
  double u;
  if (c <= 0.9375) {
    const int k = bin_of(c, 128.0, u);
    return poly7_row(kAirMassC[k], u);
  }
  const double s = sqrt_pos(0.5 * (1.0 - fmin(c, 1.0)));
  const int k = bin_of(s, 64.0, u);
  return poly7_row(kAirMassS[k], u);
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
  return __fdividef(1.0f, b);  // MUFU.RCP, ~1 ulp (the IEEE __frcp_rn costs a Newton step and a slow path)
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
TFG_HD float atan_diff32(float a, float b) {  // atan(a) - atan(b) = atan2(a - b, 1 + a*b), as atan_diff above
  const float y = a - b, x = fmaf(a, b, 1.0f);
  const float ay = fabsf(y), ax = fabsf(x);
  const bool swap = ay > ax;
  const float hi = swap ? ay : ax, lo = swap ? ax : ay;
  const float t = (hi > 0.0f) ? lo * rcp32(hi) : 0.0f;
  float r = t * horner32<kAtanP32>(t * t);
  r = swap ? 1.5707963267948966f - r : r;
  r = (x < 0.0f) ? 3.141592653589793f - r : r;
  return copysignf(r, y);
}
TFG_HD float stull_wet_bulb32(float T, float RH) {  // 0 <= RH <= 2
  const float a1 = horner32<kStull132>(RH);
  const float u = 0.023101f * RH;
  const float a4 = u * fmaf(u * u, -1.0f / 3.0f, 1.0f);
  const float t4 = (0.00391838f * (RH * sqrtf(RH))) * a4;
  return (((T * a1) + atan_diff32(T + RH, RH - 1.676331f)) + t4) - 4.86035f;
}

}  // namespace fm
}  // namespace tfg
