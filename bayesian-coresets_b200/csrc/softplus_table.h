// Fast float64 link functions for the projection kernels (K3 / K3b).
//
// The logistic-regression and Poisson log-likelihoods of the reference
//   examples/common/model_lr.py:25-32      -log1p(exp(m)), m = -z.theta  (linear branch for m >= 100)
//   examples/common/model_poiss.py:25-38   s = log softplus(x.theta),  y s - exp(s)
// are built on softplus.  Written with libdevice's exp + log1p the link costs ~110 instructions per matrix
// element and the d = 10 projection kernel spent most of its issue slots there (profiles/r01_final_project_lr_*:
// 208 instructions per element, 5 % of the HBM write rate).  Here g(t) = log1p(exp(t)), t <= 0, comes from a table:
// [-37, 0] is cut into 296 intervals of width 1/8, each with the degree-7 polynomial that interpolates g at the
// interval's 8 Chebyshev nodes (built on the host in long double).  The interpolation error is below
// |g^(8)| (1/16)^8 / (8! 2^7) ~ 1e-17 g, i.e. below double rounding, RELATIVE to g on every interval (all derivatives
// of g scale like exp(t) in the tail), so tiny values keep their relative accuracy like the libdevice path.
// Below t = -37, log1p(exp(t)) == exp(t) in float64 (the next term is exp(t) * 2^-54): rare out-of-line tail.
// One evaluation: 4 x 16-byte table loads (one 64-byte record) + ~14 float64 instructions.
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

constexpr int kSpPerUnit = 8;                            // intervals per unit of t
constexpr int kSpRange = 37;                             // the table covers [-37, 0]
constexpr int kSpIntervals = kSpRange * kSpPerUnit;      // 296
constexpr int kSpStride = 8;                             // coefficients per interval (degree 7): one 64-byte record
constexpr size_t kSpTableDoubles = (size_t)kSpIntervals * kSpStride;

// host: coefficients c_0..c_7 of p(x) = sum c_k x^k, x = 2 (8 t - floor(8 t)) - 1 in [-1, 1), per interval
inline void softplus_table_build(double* tab) {
  const int n = kSpStride;
  const long double pi = 3.14159265358979323846264338327950288L;
  // monomial coefficients of the Chebyshev polynomials T_0..T_7
  long double Tm[8][8];
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < n; ++j) Tm[k][j] = 0.L;
  Tm[0][0] = 1.L;
  Tm[1][1] = 1.L;
  for (int k = 2; k < n; ++k) {
    for (int j = 0; j < n; ++j) Tm[k][j] = -Tm[k - 2][j];
    for (int j = 1; j < n; ++j) Tm[k][j] += 2.L * Tm[k - 1][j - 1];
  }
  for (int i = 0; i < kSpIntervals; ++i) {
    const long double lo = (long double)(i - kSpIntervals) / kSpPerUnit;       // interval [lo, lo + 1/8]
    const long double half = 0.5L / kSpPerUnit;
    long double f[8], cheb[8];
    for (int j = 0; j < n; ++j) {
      const long double xj = cosl(pi * (j + 0.5L) / n);
      f[j] = log1pl(expl(lo + half * (xj + 1.L)));
    }
    for (int k = 0; k < n; ++k) {
      long double s = 0.L;
      for (int j = 0; j < n; ++j) s += f[j] * cosl(pi * k * (j + 0.5L) / n);
      cheb[k] = s * 2.L / n;
    }
    cheb[0] *= 0.5L;
    for (int j = 0; j < n; ++j) {
      long double c = 0.L;
      for (int k = 0; k < n; ++k) c += cheb[k] * Tm[k][j];
      tab[(size_t)i * kSpStride + j] = (double)c;
    }
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
  const double u = t * (double)kSpPerUnit;               // exact
#ifdef __CUDA_ARCH__
  int i = __double2int_rd(u);                            // floor; NaN -> 0
#else
  int i = (u == u) ? (int)floor(u) : 0;
#endif
  i += kSpIntervals;
  i = i < 0 ? 0 : (i > kSpIntervals - 1 ? kSpIntervals - 1 : i);   // t = 0 evaluates the last interval at x = 1
  const double x = fma(2., u - (double)(i - kSpIntervals), -1.);
  const double* c = tab + (size_t)i * kSpStride;
#ifdef __CUDA_ARCH__
  const double2 c01 = __ldg(reinterpret_cast<const double2*>(c));
  const double2 c23 = __ldg(reinterpret_cast<const double2*>(c) + 1);
  const double2 c45 = __ldg(reinterpret_cast<const double2*>(c) + 2);
  const double2 c67 = __ldg(reinterpret_cast<const double2*>(c) + 3);
  double p = c67.y;
  p = fma(p, x, c67.x);
  p = fma(p, x, c45.y);
  p = fma(p, x, c45.x);
  p = fma(p, x, c23.y);
  p = fma(p, x, c23.x);
  p = fma(p, x, c01.y);
  p = fma(p, x, c01.x);
  return p;
#else
  double p = c[7];
  for (int k = 6; k >= 0; --k) p = fma(p, x, c[k]);
  return p;
#endif
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
