// K3b for large feature dimension on the float64 tensor cores: column sums of the row-centred log-likelihood matrix
// WITHOUT materialising it -- the project(data).sum(axis=0) inside every SparseVI / BatchPSVI optimisation step
// (reference: coreset/sparsevi.py:71-72, coreset/bpsvi.py:49-51 with projector.py:19-21).
//
//   colsum_s = sum_n (ll_ns - mean_t ll_nt) = rawsum_s - (1/S) sum_t rawsum_t ,  rawsum_s = sum_n ll_ns
// so no per-row centring pass is needed: the kernel is a float64 GEMM (rows x d) . (d x S) with the model's link
// applied to the accumulators in registers and a column reduction as its epilogue.  The contraction is issued as DMMA
// (mma.sync.m8n8k4.f64: 256 FMA per warp instruction instead of 32), the only tensor-core path that keeps float64
// accumulation -- tcgen05 has no f64 kind, and these sums feed gradients that are differences of O(N) sums.
//
// CTA tile 128 rows x 128 columns, 8 warps as 4 x 2, warp tile 32 x 64 = 4 x 8 DMMA tiles, 2 accumulators per lane and
// tile (64 doubles per lane).  History (profiles/r01_analysis.md, r01b_projsum_*): a register-tiled kernel on the FMA
// pipe reached 50 % of the float64 pipe; the first DMMA version 22.9 TFLOP/s with single-buffered operand tiles (DMMA
// pipe 64 % active, 0.9 barrier-stall cycles per issue, 208 M shared-memory bank conflicts); this version:
//   * operand tiles double-buffered in dynamic shared memory: the stores of k tile kt+1 go to the other buffer while
//     tile kt is being multiplied, ONE block barrier per k tile;
//   * conflict-free tile stores: the z loader is warp-uniform in k (threads 0-127 carry k 0-7 of rows 0-127, threads
//     128-255 carry k 8-15), so a warp writes 32 consecutive doubles of one k row; the theta loader takes columns
//     (t & 15) + 16 q of k row t >> 4, so a half-warp writes 16 consecutive doubles;
//   * fragment loads from rows of 132 doubles: lane = 4 g + t reads [k0 + t][base + g], conflict-free per half-warp.
// Measured back to back on one box (profiles/r01b_projsum_mma2.txt): 9.1 / 11.1 / 15.8 ms per pass (Gaussian / LR /
// Poisson, N = 1e6, d = 200, S = 512) against 9.9 / 12.0 / 16.7 ms for the single-buffered version, which it replaced.
// Bound: float64 tensor pipe (2 N d S flops) plus N S link evaluations (LR / Poisson).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "project_kernels.cuh"

namespace bcg {

constexpr int kPsBM = 128, kPsBN = 128, kPsThreads = 256;
constexpr int kPmKT = 16;            // k per shared-memory tile (4 MMA k-steps)
constexpr int kPmLd = 132;           // padded row length of both operand tiles (doubles)

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

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

constexpr size_t kPm2TileDoubles = (size_t)kPmKT * kPmLd;                       // one operand tile
constexpr size_t kPm2SmemBytes = (4 * kPm2TileDoubles + kPsBM + 4 * kPsBN) * sizeof(double);   // 2 x (z, theta) + ys + colacc

template <int MODEL>
__global__ void __launch_bounds__(kPsThreads, 1) project_sum_mma_kernel(const ProjectSumArgs a) {
  extern __shared__ __align__(16) double pm2_smem[];
  double* zs0 = pm2_smem;                               // [2][kPmKT][kPmLd]
  double* ts0 = pm2_smem + 2 * kPm2TileDoubles;         // [2][kPmKT][kPmLd]
  double* ys = pm2_smem + 4 * kPm2TileDoubles;          // [kPsBM]
  double* colacc = ys + kPsBM;                          // [4][kPsBN]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wr = warp >> 1, wc = warp & 1;
  const int g = lane >> 2, tq = lane & 3;               // MMA fragment coordinates
  const int S = a.S, d = a.d;
  const int ncoltiles = (S + kPsBN - 1) / kPsBN;
  const int64_t nrowblocks = (a.n + kPsBM - 1) / kPsBM;
  const int nk = (d + kPmKT - 1) / kPmKT;
  // loader roles: z tile 128 rows x 16 k (8 consecutive k of one row per thread, warp-uniform k half),
  // theta tile 16 k x 128 columns (k row t >> 4, columns (t & 15) + 16 q)
  const int zrow = t & 127, zhalf = t >> 7;
  const int tk = t >> 4, tc0 = t & 15;

  double mysum[4] = {0., 0., 0., 0.};

  for (int64_t rb = blockIdx.x; rb < nrowblocks; rb += gridDim.x) {
    const int64_t row0 = rb * kPsBM;
    if (MODEL == MODEL_POISSON) {
      __syncthreads();
      if (t < kPsBM) ys[t] = (row0 + t < a.n) ? a.Z[(a.rowidx ? a.rowidx[row0 + t] : row0 + t) * a.zld + d] : 0.;
    }
    const int64_t zr_live = row0 + zrow;
    const int64_t zr = (zr_live < a.n) ? (a.rowidx ? a.rowidx[zr_live] : zr_live) : 0;
    for (int ct = 0; ct < ncoltiles; ++ct) {
      const int col0 = ct * kPsBN;
      double acc[4][8][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.; acc[mi][ni][1] = 0.; }

      double zreg[8], treg[8];
      auto gload = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = k0 + zhalf * 8 + q;
          zreg[q] = (zr_live < a.n && k < d) ? a.Z[zr * a.zld + k] : 0.;
        }
        const int k = k0 + tk;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = col0 + tc0 + 16 * q;
          treg[q] = (k < d && c < S) ? a.thetaT[(size_t)k * S + c] : 0.;
        }
      };
      auto sstore = [&](int buf) {
        double* zs = zs0 + (size_t)buf * kPm2TileDoubles;
        double* ts = ts0 + (size_t)buf * kPm2TileDoubles;
#pragma unroll
        for (int q = 0; q < 8; ++q) zs[(size_t)(zhalf * 8 + q) * kPmLd + zrow] = zreg[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) ts[(size_t)tk * kPmLd + tc0 + 16 * q] = treg[q];
      };
      gload(0);
      sstore(0);
      __syncthreads();
      for (int kt = 0; kt < nk; ++kt) {
        const bool more = kt + 1 < nk;
        if (more) gload((kt + 1) * kPmKT);                // global loads in flight during this tile's DMMAs
        const double* zs = zs0 + (size_t)(kt & 1) * kPm2TileDoubles;
        const double* ts = ts0 + (size_t)(kt & 1) * kPm2TileDoubles;
#pragma unroll
        for (int kk = 0; kk < kPmKT; kk += 4) {
          double af[4], bf[8];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) af[mi] = zs[(size_t)(kk + tq) * kPmLd + wr * 32 + mi * 8 + g];
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) bf[ni] = ts[(size_t)(kk + tq) * kPmLd + wc * 64 + ni * 8 + g];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        if (more) sstore((kt + 1) & 1);                   // the other buffer: last read before the previous barrier
        __syncthreads();
      }

      // ---- epilogue: accumulator (mi, ni, e) is row wr*32 + mi*8 + g, column wc*64 + ni*8 + tq*2 + e.  First with the
      // branch-free links (64 independent evaluations the compiler can interleave); when any lane of the warp met an
      // argument outside their range (|lin| > 37: rare) the tile's sums are redone with the branching links.
      double cs[8][2];
      auto epilogue = [&](auto link) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) { cs[ni][0] = 0.; cs[ni][1] = 0.; }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const int rl = wr * 32 + mi * 8 + g;
          const bool live = row0 + rl < a.n;
          const double y = (MODEL == MODEL_POISSON) ? ys[rl] : 0.;
#pragma unroll
          for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = col0 + wc * 64 + ni * 8 + tq * 2 + e;
              double lin = acc[mi][ni][e];
              if (a.coff && c < S) lin += a.coff[c];
              const double v = link(lin, y);
              cs[ni][e] += (live && c < S) ? v : 0.;
            }
        }
      };
      if (MODEL == MODEL_LINEAR || !a.sp_tab) {
        epilogue([&](double lin, double y) { return link_apply<MODEL>(a.sp_tab, lin, y); });
      } else {
        bool tail = false;
        epilogue([&](double lin, double y) {
          tail |= link_needs_tail(lin);
          return MODEL == MODEL_LR ? lr_link_nb(a.sp_tab, lin) : poisson_link_nb(a.sp_tab, lin, y);
        });
        if (__any_sync(0xffffffffu, tail)) epilogue([&](double lin, double y) { return link_apply<MODEL>(a.sp_tab, lin, y); });
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e) {                      // over the 8 row groups g (lane bits 2..4)
          double v = cs[ni][e];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          cs[ni][e] = v;
        }
      __syncthreads();
      if (g == 0) {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          colacc[wr * kPsBN + wc * 64 + ni * 8 + tq * 2] = cs[ni][0];
          colacc[wr * kPsBN + wc * 64 + ni * 8 + tq * 2 + 1] = cs[ni][1];
        }
      }
      __syncthreads();
      if (t < kPsBN) {
        const double v = colacc[t] + colacc[kPsBN + t] + colacc[2 * kPsBN + t] + colacc[3 * kPsBN + t];
        if (ct < 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (q == ct) mysum[q] += v;
        } else if (col0 + t < S) {
          a.partial[(size_t)blockIdx.x * S + col0 + t] += v;
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
