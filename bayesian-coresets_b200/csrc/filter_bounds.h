// Rigorous bounds for the half-precision pre-filter of the selection scan (filter16 in loop_kernel.cuh).
//
// The persistent greedy kernel may stream a float16 COPY of the unit rows (2 N S bytes per iteration instead of 4 N S)
// and then re-scan, with the unchanged float32 arithmetic of scan_core.cuh, only those row groups that can still hold
// the float32 arg-max or a row inside its near-tie window.  The selection is therefore bit-identical to the plain
// float32 scan provided the float32 score of EVERY row lies inside the interval [lb, ub] computed here from its
// float16 inner products.  This header is that proof obligation, host- and device-compilable so that
// tests/hostcheck can check it against brute force.
//
// Notation: a = float32 unit row (|a|_2 <= 1 + 1e-4; zero rows allowed), h = fl16(a) (round to nearest),
// d = float32 direction, s = fl32 sum a_i d_i (any summation order, fused multiply-adds), t = fl32 sum h_i d_i.
//   quantisation   |h_i - a_i| <= 2^-11 |a_i| + 2^-25        (normal range / subnormal spacing 2^-24, |a_i| <= 1.0001)
//   =>             |sum h_i d_i - sum a_i d_i| <= 2^-11 |a|_2 |d|_2 + 2^-25 sqrt(S) |d|_2          (Cauchy-Schwarz)
//   accumulation   each float32 sum of n = S fused terms errs by <= gamma_S |a|_2 |d|_2,  gamma_S = S u / (1 - S u), u = 2^-24
//   =>             |t - s| <= E = (2^-11 * 1.0001 + 2^-25 sqrt(S) + 2.2 gamma_S) * |d|_2 * 1.001    (filter_eps below)
// Scores (scan_core.cuh):  Frank-Wolfe / OrthoPursuit  score = s0;   GIGA  den = 1 - s1^2,
//   score = (s1 > -1 && den > 0) ? s0 * rsqrtf(den) : 0.   For GIGA the float32 evaluation of den is only trusted when the
//   interval of den stays above kDenFloor (relative error of den <= 2.4e-7 / kDenFloor); below it the row is declared
//   unbounded (ub = +inf when its numerator can be positive) and is simply re-scanned.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BCG_HD __host__ __device__ __forceinline__
#else
#define BCG_HD inline
#endif

namespace bcg {

// Budget of kScoreRel (all relative to the score, den >= kDenFloor = 1e-3):
//   the scan's own evaluation      den_f32 = fl(1 - s1^2) is off by <= 1.2e-7 absolute = 1.2e-4 relative  -> 0.60e-4 on rsqrt
//                                  rsqrtf: 2 ulp = 2.4e-7;  the product s0 * rsqrtf(den): 0.6e-7           -> 0.62e-4 in total
//   this header's evaluation       dmin / dmax carry the same 1.2e-7 absolute error                         -> 0.60e-4 on rmin / rmax
//                                  (rsqrtf / the product again 3e-7; the rounding of |t1| +- e1 and of t0 +- e0 is covered by
//                                   the 0.1 % by which filter_eps_unit inflates E: 5e-7 >> 6e-8)
//   sum 1.25e-4 < kScoreRel = 2e-4.  kScoreAbs covers scores near zero.
constexpr float kDenFloor = 1e-3f;
constexpr float kScoreRel = 2e-4f;     // relative slack of the float32 score evaluation when den >= kDenFloor
constexpr float kScoreAbs = 1e-7f;

// E / |d|_2 for rows of S elements
BCG_HD float filter_eps_unit(int S) {
  const double u = 5.9604644775390625e-8;                 // 2^-24
  const double gamma = (double)S * u / (1. - (double)S * u);
  const double e = 4.8828125e-4 * 1.0001 + 2.98023223876953125e-8 * sqrt((double)S) + 2.2 * gamma;
  return (float)(e * 1.001);
}

// near-tie window of the float32 arg-max (loop_kernel.cuh / scan_kernel.cuh: 2e-5 + 1e-5 |top|), with slack; a row group
// is re-scanned when its upper bound reaches filter_threshold(L), L = a lower bound of the float32 maximum
BCG_HD float filter_threshold(float L) {
  if (!(L < 3.0e38f)) return 3.0e38f;                     // +inf (a row with an infinite score): only unbounded rows pass
  if (!(L > -3.0e38f)) return -INFINITY;                  // nothing bounded from below yet: everything passes
  return L - (2.5e-5f + 1.2e-5f * fabsf(L));
}

// Frank-Wolfe / OrthoPursuit: score = s0
BCG_HD void filter_bounds_lin(float t0, float e0, float* lb, float* ub) {
  *lb = t0 - e0;
  *ub = t0 + e0;
}

// GIGA: bounds of (s1 > -1 && 1 - s1^2 > 0) ? s0 * rsqrtf(1 - s1^2) : 0 over s0 in t0 +- e0, s1 in t1 +- e1
BCG_HD void filter_bounds_giga(float t0, float t1, float e0, float e1, float* lb, float* ub) {
  const float a1 = fabsf(t1);
  const float a1hi = fminf(a1 + e1, 1.f);
  const float a1lo = fmaxf(a1 - e1, 0.f);
  const float dmin = 1.f - a1hi * a1hi;                   // smallest possible den
  const float dmax = 1.f - a1lo * a1lo;                   // largest possible den
  const float u0 = t0 + e0, l0 = t0 - e0;
  if (!(dmin >= kDenFloor)) {                             // (also taken when t1 is NaN)
    *ub = (u0 > 0.f || !(u0 == u0)) ? INFINITY : 0.f;     // masked rows score exactly 0
    *lb = -INFINITY;
    return;
  }
#if defined(__CUDA_ARCH__)
  const float rmin = rsqrtf(dmax), rmax = rsqrtf(dmin);                // 2 ulp; inside kScoreRel together with the float32
#else                                                                  // evaluation of dmin / dmax (<= 1.3e-4 relative)
  const float rmin = 1.f / sqrtf(dmax), rmax = 1.f / sqrtf(dmin);      // rmin <= 1/sqrt(den) <= rmax
#endif
  const float u = (u0 > 0.f) ? u0 * rmax : u0 * rmin;
  const float l = (l0 > 0.f) ? l0 * rmin : l0 * rmax;
  *ub = u + (kScoreRel * fabsf(u) + kScoreAbs);
  *lb = l - (kScoreRel * fabsf(l) + kScoreAbs);
}

}  // namespace bcg
