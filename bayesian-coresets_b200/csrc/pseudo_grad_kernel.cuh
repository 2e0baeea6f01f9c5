// (n, S, d) gradients of the log-likelihood with respect to the DATAPOINTS, for the K pseudo-points of BatchPSVI
// (bpsvi.py:37,53 -> projector.py:23-28 -> grad_loglikelihood):
//   model_lr.py:50-57         glls[k,s,:] = sigma(m_ks) theta_s ,  m = -z.theta  (sigma = e^m / (1 + e^m), 1 for m >= 100)
//   model_poiss.py:58-67      glls[k,s,:] = g_ks [theta_s, 0]  with the reference's broadcast defect repaired:
//                             g = y - e^s, and where e^s > 1e-15: (y e^-s - 1)(1 - exp(-e^s)),  s = log softplus(x.theta)
//   model_gaussian.py:12-15   glls[k,s,:] = theta_s Siginv - x_k Siginv
// centred over the LAST axis (projector.py:26: `glls -= glls.mean(axis=2)`, replicated as is), and their contraction
//   bpsvi.py:53               ugrad[k,:] = -(1/S) sum_s w_k resid_s glls[k,s,:]
// Every model has the form glls[k,s,:] = g_ks U_s + V_k; the host passes the row-centred U (S x dz) and V (K x dz), the
// kernel evaluates g (the transcendental part, float64 libdevice) and either writes the (K, S, dz) array or contracts it
// on the fly -- the (K, S, dz) array is never formed for the gradient step.  One CTA per pseudo-point; K-sized work.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/bcg.h"

namespace bcg {

struct PseudoGradArgs {
  const double* pts;      // K x zld
  const double* theta;    // S x d (row major, NOT transposed)
  const double* Uc;       // S x dz, rows centred over dz
  const double* Vc;       // K x dz, rows centred over dz, or null
  const double* w;        // K      (contraction) or null
  const double* resid;    // S      (contraction) or null
  double* glls;           // K x S x dz or null
  double* ugrad;          // K x dz or null
  int32_t K, zld, d, dz, S, model;
};

__global__ void __launch_bounds__(256) pseudo_grad_kernel(const PseudoGradArgs a) {
  extern __shared__ double pg_smem[];           // g[S], then z[d + 1]
  double* g = pg_smem;
  double* z = pg_smem + a.S;
  const int k = blockIdx.x, t = threadIdx.x;
  const int S = a.S, d = a.d, dz = a.dz;
  for (int i = t; i <= d && i < a.zld; i += blockDim.x) z[i] = a.pts[(size_t)k * a.zld + i];
  __syncthreads();
  const double y = a.model == BCG_MODEL_POISSON ? z[d] : 0.;
  for (int s = t; s < S; s += blockDim.x) {
    double gs = 1.;
    if (a.model != BCG_MODEL_GAUSSIAN) {
      const double* th = a.theta + (size_t)s * d;
      double lin = 0.;
      for (int i = 0; i < d; ++i) lin = fma(z[i], th[i], lin);
      if (a.model == BCG_MODEL_LR) {
        const double m = -lin;                                          // model_lr.py:53-56
        gs = 1.;
        if (m < 100.) { const double e = exp(m); gs = e / (1. + e); }
      } else {
        double sv = lin;                                                // model_poiss.py:25-30 (compute_s)
        if (sv > -100.) sv = log(fmax(sv, 0.) + log1p(exp(-fabs(sv))));
        const double es = exp(sv);
        gs = y - es;                                                    // model_poiss.py:64-66
        if (es > 1e-15) gs = (y * exp(-sv) - 1.) * (1. - exp(-es));
      }
    }
    g[s] = gs;
  }
  __syncthreads();
  if (a.glls) {
    double* out = a.glls + (size_t)k * S * dz;
    for (size_t i = t; i < (size_t)S * dz; i += blockDim.x) {
      const int s = (int)(i / dz), j = (int)(i - (size_t)s * dz);
      out[i] = g[s] * a.Uc[(size_t)s * dz + j] + (a.Vc ? a.Vc[(size_t)k * dz + j] : 0.);
    }
  }
  if (a.ugrad) {
    double rsum = 0.;
    if (a.Vc)
      for (int s = 0; s < S; ++s) rsum += a.resid[s];
    const double scale = -a.w[k] / (double)S;
    for (int j = t; j < dz; j += blockDim.x) {
      double acc = 0.;
      for (int s = 0; s < S; ++s) acc = fma(a.resid[s] * g[s], a.Uc[(size_t)s * dz + j], acc);    // coalesced over j
      if (a.Vc) acc = fma(rsum, a.Vc[(size_t)k * dz + j], acc);
      a.ugrad[(size_t)k * dz + j] = scale * acc;
    }
  }
}

}  // namespace bcg
