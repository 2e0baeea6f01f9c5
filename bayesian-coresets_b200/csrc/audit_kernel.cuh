// Independent float64 audit scorer: the selection pass of the reference recomputed from the RAW data.
//
// For every data row n the kernel re-evaluates, in float64 and with libdevice's exp / log1p / log / lgamma,
//   projector.py:19-21            lls = loglikelihood(pts, samples); lls -= lls.mean(axis=1)
//   model_lr.py:25-32             m = -z.theta ; m < 100 ? -log1p(exp(m)) : -m
//   model_poiss.py:25-38          s = log softplus(x.theta) (s > -100) ; y s - gammaln(y + 1) - exp(s)
//   model_gaussian.py:4-10        -1/2 (x Si x + theta Si theta - 2 x Si theta)   (+ constants that cancel)
//   giga.py:10-13                 row norm, unit row
// and scores the unit row against the direction(s) of ONE greedy iteration
//   giga.py:31-38                 s = An^T [cdir | xw] ; masked s0 / sqrt(1 - s1^2)          (kind = BCG_ALG_GIGA)
//   frankwolfe.py:17, orthopursuit.py:19, sparsevi.py:51   An^T residual                      (kind = FW / OMP)
// It writes one float64 score (and optionally the norm) per row and accumulates the column sums b.
//
// Purpose: parity evidence at the benchmarked sizes (N = 1e7, S = 512), where a float64 host oracle of the
// N x S matrix does not fit the test budget.  Deliberately shares NOTHING with the production path: no float32
// storage, no TMA, no scan_core.cuh, no softplus table, no step_logic.h -- one warp per row, plain loads.  It is
// also the first half of the never-materialising select of SURVEY.md section 8(f) rank 2 (bytes 8 N d_in instead
// of 4 N S; float64-pipe bound: ~5 ps per matrix element).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/bcg.h"

namespace bcg {

struct AuditArgs {
  const double* Z;        // n x zld
  int64_t n;
  int32_t zld, d, S;
  int32_t model;          // BCG_MODEL_*
  int32_t kind;           // BCG_ALG_GIGA: dirs = [cdir | xw]; otherwise dirs = [residual]
  const double* thetaT;   // d x S (LR / Poisson: samples^T; Gaussian: (theta Siginv)^T)
  const double* tt;       // S: theta_s Siginv theta_s (Gaussian) or null
  const double* Siginv;   // d x d (Gaussian) or null
  const double* dirs;     // ndir x S or null (no scores requested)
  double* scores;         // n or null
  double* norms;          // n or null
  double* colsum;         // S, zero-initialised, or null
};

__device__ __forceinline__ double audit_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// centred log-likelihood vector of data row `row` (lane owns columns lane + 32 j) and its norm; shared by the audit
// scorer and the never-materialising select
template <int J>
__device__ __forceinline__ double audit_row(const AuditArgs& a, int64_t row, int lane, double (&v)[J]) {
  const int S = a.S, d = a.d;
  const double* z = a.Z + row * a.zld;
#pragma unroll
  for (int j = 0; j < J; ++j) v[j] = 0.;
  for (int k = 0; k < d; ++k) {
    const double zk = z[k];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < S) v[j] = fma(zk, __ldg(a.thetaT + (size_t)k * S + s), v[j]);
    }
  }
  double xsx = 0., y = 0., lg = 0.;
  if (a.model == BCG_MODEL_GAUSSIAN) {
    // x Siginv x (row constant; kept because the reference evaluates it before centring)
    double part = 0.;
    for (int i = lane; i < d; i += 32) {
      double t = 0.;
      for (int k = 0; k < d; ++k) t = fma(__ldg(a.Siginv + (size_t)i * d + k), z[k], t);
      part = fma(z[i], t, part);
    }
    xsx = audit_warp_sum(part);
  } else if (a.model == BCG_MODEL_POISSON) {
    y = z[d];
    lg = lgamma(y + 1.);
  }
  double sum = 0.;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) {
      double ll;
      if (a.model == BCG_MODEL_LR) {
        const double m = -v[j];
        ll = (m < 100.) ? -log1p(exp(m)) : -m;
      } else if (a.model == BCG_MODEL_POISSON) {
        double t = v[j];
        if (t > -100.) t = log(fmax(t, 0.) + log1p(exp(-fabs(t))));
        ll = y * t - lg - exp(t);
      } else {
        ll = -0.5 * (xsx + __ldg(a.tt + s) - 2. * v[j]);
      }
      v[j] = ll;
      sum += ll;
    }
  }
  const double mean = audit_warp_sum(sum) / (double)S;
  double ss = 0.;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) { v[j] -= mean; ss = fma(v[j], v[j], ss); } else v[j] = 0.;
  }
  return sqrt(audit_warp_sum(ss));
}

// score of a centred row against the direction(s): giga.py:33-38 / frankwolfe.py:17
template <int J>
__device__ __forceinline__ double audit_score(const double (&v)[J], double norm, const double* dirs, int S, bool giga, int lane) {
  double p0 = 0., p1 = 0.;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) {
      p0 = fma(v[j], dirs[s], p0);
      if (giga) p1 = fma(v[j], dirs[S + s], p1);
    }
  }
  p0 = audit_warp_sum(p0);
  double score;
  if (giga) {
    p1 = audit_warp_sum(p1);
    const double s0 = p0 / norm, s1 = p1 / norm;
    const bool ok = (s1 > -1. + 1e-14) && (1. - s1 * s1 > 0.);            // giga.py:33
    score = ok ? s0 / sqrt(1. - s1 * s1) : s0 / INFINITY;                 // giga.py:34-38
  } else {
    score = p0 / norm;
  }
  if (!(norm > 0.)) score = -INFINITY;      // the reference rejects zero rows at construction (giga.py:11-12)
  return score;
}

template <int J>
__global__ void __launch_bounds__(256) audit_score_kernel(const AuditArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int S = a.S;
  double cs[J];
#pragma unroll
  for (int j = 0; j < J; ++j) cs[j] = 0.;
  for (int64_t row = warp; row < a.n; row += nwarps) {
    double v[J];
    const double norm = audit_row<J>(a, row, lane, v);
#pragma unroll
    for (int j = 0; j < J; ++j) cs[j] += v[j];
    if (a.norms && lane == 0) a.norms[row] = norm;
    if (a.scores) {
      const double score = audit_score<J>(v, norm, a.dirs, S, a.kind == BCG_ALG_GIGA, lane);
      if (lane == 0) a.scores[row] = score;
    }
  }
  if (a.colsum) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < S) atomicAdd(a.colsum + s, cs[j]);
    }
  }
}

}  // namespace bcg
