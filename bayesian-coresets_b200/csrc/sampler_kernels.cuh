// Device side of the host-side samplers that feed Projector.update().
//
// Weighted Gaussian posterior (examples/common/model_gaussian.py:23-30 `weighted_post`, used by the SparseVI / BatchPSVI
// sampler of examples/gaussian/main.py:107-113):
//   L = chol(Sig0inv + sum(w) Siginv)                      (lower)
//   U = (L^-1)^T                                           Sigp = U U^T
//   mup = U U^T (Sig0inv th0 + Siginv sum_k w_k x_k)       (th0 when there are no points)
//   theta = mup + E U^T = mup + E L^-1                     E = the caller's standard normals (drawn on the HOST, in the
//                                                          reference's order, so seeded runs consume the global RNG alike)
// d is the parameter dimension (200 in the Gaussian example): the d x d factorisation is latency-bound single-CTA work, the
// point of doing it here is that at d = 200 NumPy spends ~3 ms per call in inv / cholesky / the S x d x d product, 101
// times per SparseVI build iteration.  All float64.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace bcg {

struct GaussPostArgs {
  const double* th0;      // d
  const double* Sig0inv;  // d x d
  const double* Siginv;   // d x d
  const double* pts;      // K x d (K may be 0)
  const double* w;        // K
  const double* E;        // S x d standard normals
  double* L;              // d x d work: precision -> its lower Cholesky factor
  double* Linv;           // d x d work: L^-1 (lower)
  double* rhs;            // 2 d work
  double* mup;            // d   out
  double* theta;          // S x d out
  int32_t d, K, S;
  int32_t* status;        // 0 ok, 1 precision not positive definite
};

// one CTA: precision, right-hand side, Cholesky, triangular inverse, posterior mean
__global__ void __launch_bounds__(1024) gauss_post_factor_kernel(const GaussPostArgs a) {
  __shared__ double s_red[32];
  __shared__ double s_piv;
  const int d = a.d, t = threadIdx.x, nt = blockDim.x;
  // wsum (fixed order) and xw = sum_k w_k x_k
  double wsum = 0.;
  for (int k = 0; k < a.K; ++k) wsum += a.w[k];
  for (int i = t; i < d * d; i += nt) a.L[i] = a.Sig0inv[i] + wsum * a.Siginv[i];
  double* xw = a.rhs + d;
  for (int j = t; j < d; j += nt) {
    double acc = 0.;
    for (int k = 0; k < a.K; ++k) acc += a.w[k] * a.pts[(size_t)k * d + j];
    xw[j] = acc;
  }
  __syncthreads();
  for (int i = t; i < d; i += nt) {
    double acc = 0.;
    for (int j = 0; j < d; ++j) acc += a.Sig0inv[(size_t)i * d + j] * a.th0[j] + a.Siginv[(size_t)i * d + j] * xw[j];
    a.rhs[i] = acc;
  }
  // left-looking Cholesky (Cholesky-Crout), lower, in place: column j needs the dot products of row i (i >= j) with row j
  // over the finished columns k < j -- both rows are contiguous, so a WARP per row reads them coalesced and reduces by
  // shuffles; two block barriers per column (the first version -- a right-looking rank-1 update of the trailing block by
  // all threads with a div/mod per element -- took ~4 ms at d = 200, slower than NumPy)
  const int lane = t & 31, warp = t >> 5, nw = nt >> 5;
  for (int j = 0; j < d; ++j) {
    __syncthreads();
    const double* rj = a.L + (size_t)j * d;
    for (int i = j + warp; i < d; i += nw) {
      double* ri = a.L + (size_t)i * d;
      double acc = 0.;
      for (int k = lane; k < j; k += 32) acc = fma(ri[k], rj[k], acc);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) {
        const double v = ri[j] - acc;
        if (i == j) {
          s_piv = v > 0. ? sqrt(v) : 0.;
          if (!(v > 0.)) *a.status = 1;
        } else {
          ri[j] = v;                                          // scaled by the pivot after the barrier
        }
      }
    }
    __syncthreads();
    const double piv = s_piv;
    if (piv == 0.) return;
    for (int i = j + t; i < d; i += nt) a.L[(size_t)i * d + j] = (i == j) ? piv : a.L[(size_t)i * d + j] / piv;
  }
  __syncthreads();
  for (int q = t; q < d * d; q += nt) { const int r = q / d, c = q % d; if (c > r) a.L[q] = 0.; }
  __syncthreads();
  // Linv: column c of L^-1 by forward substitution, one thread per column
  for (int c = t; c < d; c += nt) {
    for (int r = 0; r < d; ++r) {
      double acc = (r == c) ? 1. : 0.;
      for (int k = c; k < r; ++k) acc -= a.L[(size_t)r * d + k] * a.Linv[(size_t)k * d + c];
      a.Linv[(size_t)r * d + c] = (r < c) ? 0. : acc / a.L[(size_t)r * d + r];
    }
  }
  __syncthreads();
  // mup = Linv^T (Linv rhs)
  double* y = a.rhs + d;
  for (int i = t; i < d; i += nt) {
    double acc = 0.;
    for (int j = 0; j <= i; ++j) acc += a.Linv[(size_t)i * d + j] * a.rhs[j];
    y[i] = acc;
  }
  __syncthreads();
  for (int i = t; i < d; i += nt) {
    double acc = 0.;
    for (int j = i; j < d; ++j) acc += a.Linv[(size_t)j * d + i] * y[j];
    a.mup[i] = a.K > 0 ? acc : a.th0[i];                   // model_gaussian.py:26-29
  }
  (void)s_red;
}

// theta = mup + E Linv     (E U^T with U^T = L^-1); one thread per output element, coalesced over the columns
__global__ void gauss_post_sample_kernel(const GaussPostArgs a) {
  const int d = a.d;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.S * d) return;
  const int s = (int)(i / d), c = (int)(i - (int64_t)s * d);
  const double* e = a.E + (size_t)s * d;
  double acc = 0.;
  for (int k = c; k < d; ++k) acc = fma(e[k], a.Linv[(size_t)k * d + c], acc);
  a.theta[i] = a.mup[c] + acc;
}

}  // namespace bcg

namespace bcg {

// Weighted log-joint of the GLM example models over the K coreset points, with its gradient and Hessian in theta -- the
// reductions of the Laplace sampler (examples/logistic_poisson_regression/main.py:16-41 `get_laplace`, which minimises
// -log_joint with its gradient and factorises -hess at the optimum):
//   model_lr.py:25-32, 34-39, 41-48, 59-80       log_joint / grad_th_log_joint / hess_th_log_joint (logistic regression)
//   model_poiss.py:32-38, 40-45, 47-56, 69-93    the same for Poisson regression with the softplus rate
// value = sum_k w_k ll_k - d/2 log(2 pi) - |theta|^2 / 2 ; grad = -theta + sum_k w_k g_k x_k ; hess = -I + sum_k w_k h_k x_k x_k^T
// One CTA: per-point scalars (ll, g, h) into shared memory, then one thread per gradient / Hessian entry.
struct GlmJointArgs {
  const double* Z;        // K x zld  (LR: z = y x; Poisson: [x, y])
  const double* w;        // K
  const double* theta;    // d
  double* value;          // 1
  double* grad;           // d
  double* hess;           // d x d
  int32_t K, zld, d, model;
};

__global__ void __launch_bounds__(256) glm_joint_kernel(const GlmJointArgs a) {
  extern __shared__ double gj_smem[];           // theta[d], then (w ll, w g, w h)[K]
  double* th = gj_smem;
  double* sl = gj_smem + a.d;
  double* sg = sl + a.K;
  double* sh = sg + a.K;
  __shared__ double s_red[8];
  const int t = threadIdx.x, nt = blockDim.x, d = a.d, K = a.K;
  for (int i = t; i < d; i += nt) th[i] = a.theta[i];
  __syncthreads();
  double part = 0.;
  for (int k = t; k < K; k += nt) {
    const double* x = a.Z + (size_t)k * a.zld;
    double lin = 0.;
    for (int i = 0; i < d; ++i) lin = fma(x[i], th[i], lin);
    double ll, g, h;
    if (a.model == BCG_MODEL_LR) {
      const double m = -lin;                                             // model_lr.py:28-31, 44-47, 63-66
      if (m < 100.) {
        const double e = exp(m);
        ll = -log1p(e); g = e / (1. + e); h = -(e / ((1. + e) * (1. + e)));
      } else {
        ll = -m; g = 1.; h = -0.;
      }
    } else {
      const double y = x[d];
      double s = lin;                                                    // model_poiss.py:25-30
      if (s > -100.) s = log(fmax(s, 0.) + log1p(exp(-fabs(s))));
      const double es = exp(s);
      ll = y * s - lgamma(y + 1.) - es;                                  // model_poiss.py:38
      g = y - es;                                                        // model_poiss.py:53-55
      h = -(1. + y) * es;                                                // model_poiss.py:82-84
      if (es > 1e-15) {
        g = (y * exp(-s) - 1.) * (1. - exp(-es));
        h = (y * exp(-s) * (1. - exp(-s + es) + exp(-s)) - 1.) * (exp(-es) - exp(-2. * es));
      }
    }
    const double wk = a.w[k];
    sl[k] = wk * ll; sg[k] = wk * g; sh[k] = wk * h;
    part += wk * ll;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if ((t & 31) == 0) s_red[t >> 5] = part;
  __syncthreads();
  if (t == 0) {
    double v = 0., n2 = 0.;
    for (int wv = 0; wv < (nt >> 5); ++wv) v += s_red[wv];
    for (int i = 0; i < d; ++i) n2 += th[i] * th[i];
    *a.value = v - 0.5 * d * log(2. * 3.14159265358979323846) - 0.5 * n2;   // + log_prior (model_lr.py:34-36)
  }
  for (int j = t; j < d; j += nt) {
    double acc = 0.;
    for (int k = 0; k < K; ++k) acc = fma(sg[k], a.Z[(size_t)k * a.zld + j], acc);
    a.grad[j] = acc - th[j];
  }
  for (int q = t; q < d * d; q += nt) {
    const int i = q / d, j = q - i * d;
    double acc = 0.;
    for (int k = 0; k < K; ++k) acc = fma(sh[k] * a.Z[(size_t)k * a.zld + i], a.Z[(size_t)k * a.zld + j], acc);
    a.hess[q] = acc - (i == j ? 1. : 0.);
  }
}

}  // namespace bcg
