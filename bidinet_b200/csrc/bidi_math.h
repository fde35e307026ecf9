// bidi_math.h -- scalar helpers that must round exactly like the reference's x86-64
// build (g++ -O2, no FMA): compiled for the device with -fmad=false so a*b+c stays
// a rounded multiply followed by a rounded add, as in sdr/RSDR.cpp:124-131.
#ifndef BIDI_MATH_H
#define BIDI_MATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define BIDI_HD __host__ __device__ __forceinline__
#else
#define BIDI_HD static inline
#endif

// std::max / std::min as the reference uses them (first argument wins ties).
BIDI_HD float bidi_max(float a, float b) { return (a < b) ? b : a; }
BIDI_HD float bidi_min(float a, float b) { return (b < a) ? b : a; }
BIDI_HD float bidi_clamp01(float x) { return bidi_min(1.0f, bidi_max(0.0f, x)); }

// expf with the algorithm of the C library the reference links (glibc >= 2.27,
// sysdeps/ieee754/flt-32/e_expf.c, the FMA ifunc variant x86-64 selects): exp(x) =
// 2^(k/32) * 2^(r/32) with a 32-entry table and a cubic in double, rounded once to
// float.  The reference's sigmoids (sdr/RSDR.h:55-57, neo/Column.h:71-73) gate
// learning rates and exploration, so last-ulp agreement here is what lets whole
// trajectories match the oracle bit for bit.  Table: 2^(i/32) correctly rounded
// to double, minus i<<47 (checked against the running libm in tests/test_math.py).
#define BIDI_EXP2F_TABLE                                                                         \
  {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
   0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
   0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
   0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
   0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
   0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
   0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
   0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}
static const uint64_t bidi_exp2f_tab_host[32] = BIDI_EXP2F_TABLE;
#if defined(__CUDACC__)
__constant__ uint64_t bidi_exp2f_tab_dev[32] = BIDI_EXP2F_TABLE;
#endif

BIDI_HD float bidi_expf(float x) {
#if defined(__CUDA_ARCH__)
  const uint64_t* T = bidi_exp2f_tab_dev;
#else
  const uint64_t* T = bidi_exp2f_tab_host;
#endif
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
  const double Shift = 0x1.8p52;
  const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
  const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
  const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  if (!(x == x)) return x + x;
  if (x > 0x1.62e42ep6f) return INFINITY;    // overflow
  if (x < -0x1.9fe368p6f) return 0.0f;       // underflow (incl. -inf)
  double xd = (double)x;
  double z = InvLn2N * xd;
  double kd = z + Shift;
  uint64_t ki;
#if defined(__CUDA_ARCH__)
  ki = (uint64_t)__double_as_longlong(kd);
#else
  memcpy(&ki, &kd, sizeof(ki));
#endif
  kd -= Shift;
  double r = z - kd;
  uint64_t t = T[ki % 32];
  t += ki << (52 - 5);
  double s;
#if defined(__CUDA_ARCH__)
  s = __longlong_as_double((long long)t);
#else
  memcpy(&s, &t, sizeof(s));
#endif
  z = fma(C0, r, C1);
  double r2 = r * r;
  double y = fma(C2, r, 1.0);
  y = fma(z, r2, y);
  y = y * s;
  return (float)y;
}

// 1/(1+exp(-x)) exactly as sdr/RSDR.h:55-57.
BIDI_HD float bidi_sigmoid(float x) { return 1.0f / (1.0f + bidi_expf(-x)); }

#endif
