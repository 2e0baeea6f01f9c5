// K3: build the N x S matrix on the device.
//
// Replaces (reference, float64 NumPy):
//   projector.py:20-21           lls = loglikelihood(pts, samples); lls -= lls.mean(axis=1)
//   model_lr.py:25-32            LR log-likelihood          (MODEL_LR)
//   model_gaussian.py:4-10       Gaussian log-likelihood    (MODEL_LINEAR: after row-centring only
//                                x.Siginv.theta_s - 0.5 theta_s.Siginv.theta_s survives)
//   model_poiss.py:25-38         Poisson log-likelihood     (MODEL_POISSON: gammaln(y+1) cancels)
//   giga.py:10-13                Anorms = sqrt((A**2).sum(0)); An = A / Anorms
//   hilbert.py:24                b = vecs.sum(axis=0)
//
// One warp owns one row at a time; lane l owns columns l, l+32, ... (J per lane, in registers),
// so the row mean / norm are two warp-shuffle reductions and the column sums accumulate in
// registers with no atomics.  All arithmetic is float64 (the reference's precision); the only
// rounding to float32 is the final store of the unit-norm row.  Column sums are reduced
// warp -> block -> grid in a fixed order, so b is bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "softplus_table.h"

namespace bcg {

enum { MODEL_LR = 0, MODEL_LINEAR = 1, MODEL_POISSON = 2 };
constexpr int kProjWarps = 16;
constexpr int kProjKTile = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// v[j] holds the centred, unnormalised value of column lane + 32 j of one row.  Outputs are optional:
// out_row (unit float32 row) + norm_out for the resident matrix, out64 (float64 row, unnormalised) for
// small host read-backs, neither for the column-sum-only pass (K3b).
template <int J>
__device__ __forceinline__ void finalize_row(const double (&v)[J], int S, int ld, int lane, float* out_row,
                                             double* norm_out, double* out64, double (&colsum)[J], double& normsum,
                                             unsigned long long* zero_rows) {
  double ss = 0.;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) { ss += v[j] * v[j]; colsum[j] += v[j]; }
  }
  if (out64) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < S) out64[s] = v[j];
    }
  }
  if (!out_row && !norm_out) return;
  ss = warp_sum(ss);
  const double norm = sqrt(ss);
  const double inv = norm > 0. ? 1. / norm : 0.;
  if (out_row) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < ld) out_row[s] = (s < S) ? (float)(v[j] * inv) : 0.f;
    }
  }
  if (lane == 0) {
    if (norm_out) *norm_out = norm;
    normsum += norm;
    if (norm == 0.) atomicAdd(zero_rows, 1ull);
  }
}

// block-level reduction of the per-warp column sums into partial[blockIdx][S+1]
template <int J>
__device__ __forceinline__ void flush_colsum(const double (&colsum)[J], double normsum, int S, double* smem_cs,
                                             double* partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) smem_cs[(size_t)warp * (S + 1) + s] = colsum[j];
  }
  if (lane == 0) smem_cs[(size_t)warp * (S + 1) + S] = normsum;
  __syncthreads();
  for (int s = threadIdx.x; s < S + 1; s += blockDim.x) {
    double t = 0.;
    for (int w = 0; w < nw; ++w) t += smem_cs[(size_t)w * (S + 1) + s];
    partial[(size_t)blockIdx.x * (S + 1) + s] = t;
  }
}

// out[s] = sum_b partial[b][s]: 32 columns x 32 row slices per block, each thread sums its slice in order and the
// slices are added in order, so the result is bit-reproducible (one thread per column with a serial loop over the
// ~7000 partial rows of a pipelined N = 1e7 projection took 0.7 ms)
constexpr int kCsrSlices = 32;
__global__ void __launch_bounds__(32 * kCsrSlices) colsum_reduce_kernel(const double* partial, int nblocks, int S1, double* out) {
  __shared__ double part[kCsrSlices][33];
  const int s = blockIdx.x * 32 + threadIdx.x;
  double t = 0.;
  if (s < S1)
    for (int b = threadIdx.y; b < nblocks; b += kCsrSlices) t += partial[(size_t)b * S1 + s];
  part[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && s < S1) {
    double tot = 0.;
    for (int y = 0; y < kCsrSlices; ++y) tot += part[y][threadIdx.x];
    out[s] = tot;
  }
}

// ---- ingest: float64 rows already on the device -> unit float32 rows + norms + column sums ----
template <int J>
__global__ void __launch_bounds__(kProjWarps * 32) ingest_kernel(const double* src, int64_t src_ld, int64_t n,
                                                               int S, int ld, float* An, double* norms,
                                                               double* partial, unsigned long long* zero_rows) {
  extern __shared__ double smem_cs[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * kProjWarps + warp, GW = (int64_t)gridDim.x * kProjWarps;
  double colsum[J];
#pragma unroll
  for (int j = 0; j < J; ++j) colsum[j] = 0.;
  double normsum = 0.;
  for (int64_t row = gw; row < n; row += GW) {
    double v[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      v[j] = (s < S) ? src[row * src_ld + s] : 0.;
    }
    finalize_row<J>(v, S, ld, lane, An + (size_t)row * ld, norms + row, nullptr, colsum, normsum, zero_rows);
  }
  flush_colsum<J>(colsum, normsum, S, smem_cs, partial);
}

// ---- model projection ---------------------------------------------------------------------
struct ProjectArgs {
  const double* Z;       // rows x zld  (LR: z = y x; LINEAR: x; POISSON: [x, y])
  const int64_t* rowidx; // optional gather: output row r reads data row rowidx[r] (subsampled tangent spaces)
  const double* theta;   // d x S, TRANSPOSED samples (LINEAR: Siginv theta^T)
  const double* coff;    // S        per-column offset (LINEAR: -0.5 theta Siginv theta) or null
  float* An;             // unit float32 rows (null: not materialised)
  uint16_t* An16 = nullptr;  // optional float16 copy of the unit rows (fl16 of the float32 value; project_fast_kernel only)
  int32_t ld16 = 0;          // halves per row of An16
  double* norms;         // row norms (null with An)
  double* out64;         // n x S float64 centred rows (null unless requested)
  double* partial;
  unsigned long long* zero_rows;
  int64_t n;
  int32_t zld, d, S, ld, model, ktile;
  const double* sp_tab;  // softplus table (softplus_table.h) or null: libdevice exp / log1p links
};

__device__ __forceinline__ double link_value(int model, double lin, double y) {
  if (model == MODEL_LR) {
    // model_lr.py:28-31: -log1p(exp(m)) for m < 100, else -m, with m = -z.theta.  Evaluated as
    // -(max(m,0) + log1p(exp(-|m|))): the same function (for m >= 100 the log1p term is < 4e-44 and
    // vanishes in float64, which is the reference's linear branch), but exp/log1p always take the
    // same code path for every lane -- the direct form diverges inside log1p and ran 9x slower.
    const double m = -lin;
    return -(fmax(m, 0.) + log1p(exp(-fabs(m))));
  }
  if (model == MODEL_POISSON) {
    double s = lin;                                      // model_poiss.py:26-29
    if (s > -100.) s = log(fmax(s, 0.) + log1p(exp(-fabs(s))));
    return y * s - exp(s);                               // - gammaln(y+1): constant per row
  }
  return lin;
}

// table-driven links (softplus_table.h) when the context carries the table
__device__ __forceinline__ double link_value_tab(const double* tab, int model, double lin, double y) {
  if (model == MODEL_LR) return lr_link_fast(tab, lin);
  if (model == MODEL_POISSON) return poisson_link_fast(tab, lin, y);
  return lin;
}

template <int J>
__global__ void __launch_bounds__(kProjWarps * 32) project_kernel(const ProjectArgs a) {
  extern __shared__ double smem[];
  // smem: theta tile [ktile][S] , then column-sum scratch [warps][S+1]
  double* th = smem;
  double* smem_cs = smem + (size_t)a.ktile * a.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = a.S, d = a.d;
  const int KT = a.ktile;
  const int ntile = (d + KT - 1) / KT;
  const int64_t rows_per_batch = kProjWarps;
  const int64_t nbatch = (a.n + rows_per_batch - 1) / rows_per_batch;

  double colsum[J];
#pragma unroll
  for (int j = 0; j < J; ++j) colsum[j] = 0.;
  double normsum = 0.;

  bool tile_loaded = false;
  for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int64_t row = batch * rows_per_batch + warp;
    const bool live = row < a.n;
    double acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.;
    for (int t = 0; t < ntile; ++t) {
      const int k0 = t * KT;
      const int kn = (d - k0 < KT) ? d - k0 : KT;
      if (ntile > 1 || !tile_loaded) {
        __syncthreads();
        for (int i = threadIdx.x; i < kn * S; i += blockDim.x) th[i] = a.theta[(size_t)k0 * S + i];
        __syncthreads();
        tile_loaded = true;
      }
      if (live) {
        const int64_t zr = a.rowidx ? a.rowidx[row] : row;
        const double zreg = (lane < kn) ? a.Z[zr * a.zld + k0 + lane] : 0.;
        for (int k = 0; k < kn; ++k) {
          const double zk = __shfl_sync(0xffffffffu, zreg, k);
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int s = lane + 32 * j;
            if (s < S) acc[j] = fma(zk, th[(size_t)k * S + s], acc[j]);
          }
        }
      }
    }
    if (live) {
      const double y = (a.model == MODEL_POISSON) ? a.Z[(a.rowidx ? a.rowidx[row] : row) * a.zld + d] : 0.;
      double sum = 0.;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int s = lane + 32 * j;
        if (s < S) {
          double lin = acc[j];
          if (a.coff) lin += a.coff[s];
          acc[j] = a.sp_tab ? link_value_tab(a.sp_tab, a.model, lin, y) : link_value(a.model, lin, y);
          sum += acc[j];
        }
      }
      const double mean = warp_sum(sum) / (double)S;     // projector.py:21
#pragma unroll
      for (int j = 0; j < J; ++j) acc[j] -= mean;
      finalize_row<J>(acc, S, a.ld, lane, a.An ? a.An + (size_t)row * a.ld : nullptr, a.An ? a.norms + row : nullptr,
                      a.out64 ? a.out64 + (size_t)row * S : nullptr, colsum, normsum, a.zero_rows);
    }
  }
  flush_colsum<J>(colsum, normsum, S, smem_cs, a.partial);
}

// unnormalised float64 rows (norm * unit row) for host read-back
__global__ void expand_rows_kernel(const float* An, const double* norms, int64_t row0, int64_t nrows, int S, int ld,
                                   double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * S) return;
  const int64_t r = i / S;
  const int s = (int)(i - r * S);
  out[i] = norms[row0 + r] * (double)An[(size_t)(row0 + r) * ld + s];
}

__global__ void expand_active_kernel(const float* rows, const double* norms, int64_t first, int64_t count, int S,
                                     int ld, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count * S) return;
  const int64_t r = i / S;
  const int s = (int)(i - r * S);
  out[i] = norms[first + r] * (double)rows[(size_t)(first + r) * ld + s];
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace bcg
