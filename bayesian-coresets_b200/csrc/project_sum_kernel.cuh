// K3b for large feature dimension: column sums of the row-centred log-likelihood matrix WITHOUT
// materialising it -- the project(data).sum(axis=0) inside every SparseVI / BatchPSVI optimisation step
// (reference: coreset/sparsevi.py:71-72, coreset/bpsvi.py:49-51 with projector.py:19-21).
//
//   colsum_s = sum_n (ll_ns - mean_t ll_nt) = rawsum_s - (1/S) sum_t rawsum_t ,  rawsum_s = sum_n ll_ns
// so no per-row centring pass is needed: the kernel is a float64 GEMM (rows x d) . (d x S) with the
// model's link applied to the accumulators in registers and a column reduction as its epilogue.
//
// Register-tiled float64 GEMM on the CUDA cores (the float32 unit rows of the resident matrix allow
// ~1e-7, but these sums feed gradients that are differences of O(N) sums, so they stay float64):
// CTA tile 128 rows x 128 columns, 8 warps as 4 x 2, lane grid 4 x 8, 8 x 8 accumulators per thread;
// per k: 8 LDS.128 (broadcast-friendly layouts) for 64 DFMA.  The CTA walks all column tiles of its
// row block (Z tile re-read from L2), row blocks are grid-strided; column sums are reduced
// lane -> warp -> CTA in registers / shared memory and written as one partial row per CTA.
// Bound: float64 FMA pipe (2 N d S flops) plus N S link evaluations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "project_kernels.cuh"

namespace bcg {

constexpr int kPsBM = 128, kPsBN = 128, kPsKT = 8, kPsThreads = 256;
constexpr int kPsZs = kPsBM + 2;     // padded row length of the transposed z tile (conflict-free stores)

struct ProjectSumArgs {
  const double* Z;       // rows x zld
  const int64_t* rowidx; // optional gather (n entries)
  const double* thetaT;  // d x S
  const double* coff;    // S or null
  double* partial;       // gridDim.x x S raw column sums
  int64_t n;
  int32_t zld, d, S, model;
  const double* sp_tab;  // softplus table or null
};

// out-of-line link: the epilogue applies it to 64 accumulators per thread; inlining 64 copies of the
// float64 exp/log1p bodies made the kernel ~600 KB of SASS and instruction-fetch bound
template <int MODEL>
__device__ __noinline__ double link_call(double lin, double y) { return link_value(MODEL, lin, y); }
template <>
__device__ __forceinline__ double link_call<MODEL_LINEAR>(double lin, double) { return lin; }
// table-driven link inline (~20 instructions), libdevice link out of line
template <int MODEL>
__device__ __forceinline__ double link_apply(const double* tab, double lin, double y) {
  if (MODEL == MODEL_LINEAR) return lin;
  if (tab) return MODEL == MODEL_LR ? lr_link_fast(tab, lin) : poisson_link_fast(tab, lin, y);
  return link_call<MODEL>(lin, y);
}

template <int MODEL>
__global__ void __launch_bounds__(kPsThreads, 1) project_sum_kernel(const ProjectSumArgs a) {
  __shared__ __align__(16) double zs[kPsKT][kPsZs];
  __shared__ __align__(16) double ts[kPsKT][kPsBN];
  __shared__ double ys[kPsBM];
  __shared__ double colacc[4][kPsBN];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wr = warp >> 1, wc = warp & 1;
  const int lr = lane >> 3, lc = lane & 7;
  const int S = a.S, d = a.d;
  const int ncoltiles = (S + kPsBN - 1) / kPsBN;
  const int64_t nrowblocks = (a.n + kPsBM - 1) / kPsBM;
  // loader roles
  const int zrow = t >> 1, zhalf = t & 1;            // z tile: row, which 4 of the 8 k's
  const int tk = t >> 5, tcol = (t & 31) * 4;        // theta tile: k, 4 consecutive columns

  // per-thread running column sum: thread t < 128 owns column t of each column tile
  double mysum[4] = {0., 0., 0., 0.};                // up to 4 column tiles (S <= 512); more handled below

  for (int64_t rb = blockIdx.x; rb < nrowblocks; rb += gridDim.x) {
    const int64_t row0 = rb * kPsBM;
    if (MODEL == MODEL_POISSON) {
      __syncthreads();
      if (t < kPsBM) ys[t] = (row0 + t < a.n) ? a.Z[(a.rowidx ? a.rowidx[row0 + t] : row0 + t) * a.zld + d] : 0.;
    }
    for (int ct = 0; ct < ncoltiles; ++ct) {
      const int col0 = ct * kPsBN;
      double acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.;

      // register prefetch of the first k tile
      double zreg[4], treg[4];
      auto gload = [&](int k0) {
        const int64_t r = row0 + zrow;
        const int64_t zr = (r < a.n && a.rowidx) ? a.rowidx[r] : r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = k0 + zhalf * 4 + q;
          zreg[q] = (r < a.n && k < d) ? a.Z[zr * a.zld + k] : 0.;
        }
        const int k = k0 + tk;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = col0 + tcol + q;
          treg[q] = (k < d && c < S) ? a.thetaT[(size_t)k * S + c] : 0.;
        }
      };
      gload(0);
      for (int k0 = 0; k0 < d; k0 += kPsKT) {
        __syncthreads();                                 // previous tile fully consumed
#pragma unroll
        for (int q = 0; q < 4; ++q) zs[zhalf * 4 + q][zrow] = zreg[q];
        *reinterpret_cast<double2*>(&ts[tk][tcol]) = make_double2(treg[0], treg[1]);
        *reinterpret_cast<double2*>(&ts[tk][tcol + 2]) = make_double2(treg[2], treg[3]);
        __syncthreads();
        if (k0 + kPsKT < d) gload(k0 + kPsKT);           // overlap the next tile's global loads with the math
#pragma unroll
        for (int k = 0; k < kPsKT; ++k) {
          double zv[8], tv[8];
          const double2* zp = reinterpret_cast<const double2*>(&zs[k][wr * 32 + lr * 8]);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const double2 v = zp[i]; zv[2 * i] = v.x; zv[2 * i + 1] = v.y; }
#pragma unroll
          for (int i = 0; i < 4; ++i) {                  // columns lc*2 + 16 i, +1 within the warp's 64
            const double2 v = *reinterpret_cast<const double2*>(&ts[k][wc * 64 + 16 * i + lc * 2]);
            tv[2 * i] = v.x; tv[2 * i + 1] = v.y;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(zv[i], tv[j], acc[i][j]);
        }
      }

      // ---- epilogue: link, mask rows beyond n, column sums --------------------------------------
      double cs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[j] = 0.;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = wr * 32 + lr * 8 + i;
        const bool live = row0 + rl < a.n;
        const double y = (MODEL == MODEL_POISSON) ? ys[rl] : 0.;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = col0 + wc * 64 + 16 * (j >> 1) + lc * 2 + (j & 1);
          double lin = acc[i][j];
          if (a.coff && c < S) lin += a.coff[c];
          const double v = link_apply<MODEL>(a.sp_tab, lin, y);
          cs[j] += (live && c < S) ? v : 0.;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {                      // over the 4 row groups of the warp
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
        cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
      }
      __syncthreads();
      if (lr == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) colacc[wr][wc * 64 + 16 * (j >> 1) + lc * 2 + (j & 1)] = cs[j];
      }
      __syncthreads();
      if (t < kPsBN) {
        const double v = colacc[0][t] + colacc[1][t] + colacc[2][t] + colacc[3][t];
        if (ct < 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (q == ct) mysum[q] += v;
        } else if (col0 + t < S) {
          a.partial[(size_t)blockIdx.x * S + col0 + t] += v;   // S > 512: accumulate in place (zero-initialised)
        }
      }
    }
  }
  if (t < kPsBN) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = q * kPsBN + t;
      if (q < ncoltiles && c < S) a.partial[(size_t)blockIdx.x * S + c] = mysum[q];
    }
  }
}

// rawsum -> centred column sums: out_s = sum_b partial[b][s] - (1/S) sum_t sum_b partial[b][t]
__global__ void project_sum_finish_kernel(const double* partial, int nblocks, int S, double* out) {
  __shared__ double tot[1024];
  __shared__ double red[32];
  const int t = threadIdx.x;
  double mine = 0.;
  for (int s = t; s < S; s += blockDim.x) {
    double v = 0.;
    for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * S + s];
    tot[s] = v;
    mine += v;
  }
  mine = warp_sum(mine);
  if ((t & 31) == 0) red[t >> 5] = mine;
  __syncthreads();
  double all = 0.;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) all += red[w];
  const double mean = all / (double)S;
  for (int s = t; s < S; s += blockDim.x) out[s] = tot[s] - mean;
}

}  // namespace bcg
