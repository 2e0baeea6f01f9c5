// Fast float64 link functions for the projection kernels (K3 / K3b).
//
// The logistic-regression and Poisson log-likelihoods of the reference
//   examples/common/model_lr.py:25-32      -log1p(exp(m)), m = -z.theta  (linear branch for m >= 100)
//   examples/common/model_poiss.py:25-38   s = log softplus(x.theta),  y s - exp(s)
// are built on softplus.  Written with libdevice's exp + log1p the link costs ~110 instructions per matrix
// element and the d = 10 projection kernel spent most of its issue slots there (profiles/r01_final_project_lr_*:
// 208 instructions per element, 5 % of the HBM write rate).  Here g(t) = log1p(exp(t)), t <= 0, is a Taylor
// expansion about the nearest node t_i = -i/32 of a table that holds only (g(t_i), sigma(t_i)): every derivative of
// softplus is a polynomial in sigma -- with q = sigma (1 - sigma):
//   g1 = sigma, g2 = q, g3 = q (1 - 2 sigma), g4 = q (1 - 6 q), g5 = q (1 - 2 sigma)(1 - 12 q),
//   g6 = q (1 - 30 q + 120 q^2)
// so ONE 16-byte load feeds an order-6 expansion (|delta| <= 1/64: remainder < 1e-16 g; all derivatives scale like
// exp(t) in the tail, so tiny values keep their RELATIVE accuracy; measured max relative error 2.5e-16 over [-37, 0]).
// The first version used 64-byte records of degree-7 Chebyshev coefficients: four gathered 16-byte loads per element
// saturated the L1 data pipe of the projection kernels (ncu: l1tex data-pipe wavefronts 97 % of peak, 13.6 wavefronts per
// load request), which is why the table shrank to one load per element at the price of ~15 more float64 operations.
// Below t = -37, log1p(exp(t)) == exp(t) in float64 (the next term is exp(t) * 2^-54): rare out-of-line tail.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SP_HD __host__ __device__ __forceinline__
#else
#define SP_HD inline
#endif

namespace bcg {

constexpr int kSpPerUnit = 32;                           // nodes per unit of t
constexpr int kSpRange = 37;                             // the table covers [-37, 0]
constexpr int kSpNodes = kSpRange * kSpPerUnit + 1;      // 1185 nodes t_i = -i / 32
constexpr size_t kSpTableDoubles = (size_t)kSpNodes * 2; // (g, sigma) per node: 16-byte records, 19 KB

// host: tab[2 i] = log1p(exp(t_i)), tab[2 i + 1] = 1 / (1 + exp(-t_i)), evaluated in long double
inline void softplus_table_build(double* tab) {
  for (int i = 0; i < kSpNodes; ++i) {
    const long double t = -(long double)i / kSpPerUnit;
    tab[2 * (size_t)i] = (double)log1pl(expl(t));
    tab[2 * (size_t)i + 1] = (double)(1.L / (1.L + expl(-t)));
  }
}

// rare tails, out of line on the device (one copy of libdevice's exp instead of one per inlined link)
#if defined(__CUDACC__)
__host__ __device__ __noinline__ inline double exp_tail(double t) { return exp(t); }
#else
inline double exp_tail(double t) { return exp(t); }
#endif

// g(t) = log1p(exp(t)) for t <= 0 (NaN propagates)
SP_HD double softplus_neg(const double* tab, double t) {
  if (t < -(double)kSpRange) return exp_tail(t);
#ifdef __CUDA_ARCH__
  int i = __double2int_rn(t * -(double)kSpPerUnit);      // nearest node; NaN -> 0
#else
  int i = (t == t) ? (int)nearbyint(t * -(double)kSpPerUnit) : 0;
#endif
  i = i < 0 ? 0 : (i > kSpNodes - 1 ? kSpNodes - 1 : i);
  const double dl = fma((double)i, 1. / kSpPerUnit, t);  // t - t_i, |dl| <= 1/64
#ifdef __CUDA_ARCH__
  const double2 n = __ldg(reinterpret_cast<const double2*>(tab) + i);
  const double g = n.x, s = n.y;
#else
  const double g = tab[2 * (size_t)i], s = tab[2 * (size_t)i + 1];
#endif
  const double q = fma(-s, s, s);                        // sigma (1 - sigma)
  const double r = fma(-2., s, 1.);
  const double g3 = q * r;
  const double g4 = q * fma(-6., q, 1.);
  const double g5 = g3 * fma(-12., q, 1.);
  const double g6 = q * fma(fma(120., q, -30.), q, 1.);
  double p = fma(dl * (1. / 6.), g6, g5);
  p = fma(dl * 0.2, p, g4);
  p = fma(dl * 0.25, p, g3);
  p = fma(dl * (1. / 3.), p, q);
  p = fma(dl * 0.5, p, s);
  return fma(dl, p, g);
}

// ---- branch-free forms ---------------------------------------------------------------------------------------------
// A projection kernel evaluates 16-64 links per lane back to back.  With the tail test of softplus_neg inside, every
// evaluation is its own branch region (SASS: BSSY / @P BRA / CALL / BSYNC around ~55 instructions) and the compiler
// cannot interleave the independent Horner chains: the float64 pipe sees a dependent chain at a time and the warps sit
// in fixed-latency "wait" stalls (ncu, profiles/r02_k3_mma_vs_fast_lr_N2e6_S512.csv: 2.1-2.5 wait-stall cycles per
// issue, float64 pipe 37-53 % busy).  The *_nb forms have no branch: they are exact for |lin| <= 37 and return
// garbage beyond; the caller ORs link_needs_tail() over its elements and, when any lane of the warp saw one (rare:
// |z.theta| > 37), re-evaluates that batch with the branching forms below.
SP_HD bool link_needs_tail(double lin) { return fabs(lin) > (double)kSpRange; }

SP_HD double softplus_neg_nb(const double* tab, double t) {
#ifdef __CUDA_ARCH__
  int i = __double2int_rn(t * -(double)kSpPerUnit);
#else
  int i = (t == t) ? (int)nearbyint(fmax(fmin(t * -(double)kSpPerUnit, 1e9), -1e9)) : 0;
#endif
  i = i < 0 ? 0 : (i > kSpNodes - 1 ? kSpNodes - 1 : i);
  const double dl = fma((double)i, 1. / kSpPerUnit, t);
#ifdef __CUDA_ARCH__
  const double2 n = __ldg(reinterpret_cast<const double2*>(tab) + i);
  const double g = n.x, s = n.y;
#else
  const double g = tab[2 * (size_t)i], s = tab[2 * (size_t)i + 1];
#endif
  const double q = fma(-s, s, s);
  const double r = fma(-2., s, 1.);
  const double g3 = q * r;
  const double g4 = q * fma(-6., q, 1.);
  // (Evaluating the two highest orders in float32 -- they enter as dl^5/120 (g5 + dl/6 g6) <= 7.8e-12 of g, so float32 would
  // leave ~1e-18 -- moves 7 of the 23 float64 operations to the FP32 pipe but adds 4 conversions on the quarter-rate XU pipe:
  // measured 22.5 vs 21.9 ms for the N = 1e7 LR projection, so it is NOT done.)
  const double g5 = g3 * fma(-12., q, 1.);
  const double g6 = q * fma(fma(120., q, -30.), q, 1.);
  double p = fma(dl * (1. / 6.), g6, g5);
  p = fma(dl * 0.2, p, g4);
  p = fma(dl * 0.25, p, g3);
  p = fma(dl * (1. / 3.), p, q);
  p = fma(dl * 0.5, p, s);
  return fma(dl, p, g);
}

SP_HD double lr_link_nb(const double* tab, double lin) {
  const double m = -lin;
  return -(fmax(m, 0.) + softplus_neg_nb(tab, -fabs(m)));
}

SP_HD double poisson_link_nb(const double* tab, double lin, double y) {
  const double sp = fmax(lin, 0.) + softplus_neg_nb(tab, -fabs(lin));
  return y * log(sp) - sp;
}

// model_lr.py:28-31 as -(max(m, 0) + log1p(exp(-|m|))), m = -lin: the same function in float64 (for m >= 100 the
// log1p term is < 4e-44 and vanishes against m, which is the reference's linear branch)
SP_HD double lr_link_fast(const double* tab, double lin) {
  const double m = -lin;
  return -(fmax(m, 0.) + softplus_neg(tab, -fabs(m)));
}

// model_poiss.py:26-38 without the row-constant gammaln(y+1): s = log(softplus(lin)) for lin > -100, else lin;
// y s - exp(s).  exp(log(softplus)) is softplus itself (the reference's round trip differs from it by
// <= |s| 2^-53 relative, below the comparison tolerance of 1e-9), so one log replaces exp + log1p + log + exp.
SP_HD double poisson_link_fast(const double* tab, double lin, double y) {
  if (lin > -100.) {
    const double sp = fmax(lin, 0.) + softplus_neg(tab, -fabs(lin));
    return y * log(sp) - sp;
  }
  return y * lin - exp_tail(lin);                        // (also the NaN path)
}

}  // namespace bcg
