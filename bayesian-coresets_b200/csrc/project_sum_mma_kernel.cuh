// K3b on the float64 tensor cores: the same computation as project_sum_kernel.cuh (column sums of the
// row-centred log-likelihood matrix without materialising it; sparsevi.py:71-72, bpsvi.py:49-51 with
// projector.py:19-21), but the (rows x d) . (d x S) contraction is issued as DMMA
// (mma.sync.m8n8k4.f64: 256 FMA per warp instruction instead of 32), which is the only tensor-core path
// that keeps float64 accumulation -- tcgen05 has no f64 kind, and these sums feed gradients that are
// differences of O(N) sums, so float32 accumulators (TF32 / BF16 splits) are not an option.
//
// CTA tile 128 rows x 128 columns, 8 warps as 4 x 2, warp tile 32 x 64 = 4 x 8 MMA tiles, 2 accumulators per
// lane and tile (64 doubles per lane).  Operand tiles live in shared memory as zs[k][row], ts[k][col] with a
// row length of 132 doubles, which makes the fragment loads (lane = 4 g + t reads [k0 + t][base + g])
// bank-conflict free.  Link, row masking and the column reduction are the epilogue, as in the CUDA-core kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "project_sum_kernel.cuh"

namespace bcg {

constexpr int kPmKT = 16;            // k per shared-memory tile (4 MMA k-steps)
constexpr int kPmLd = 132;           // padded row length of both operand tiles (doubles)

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int MODEL>
__global__ void __launch_bounds__(kPsThreads, 1) project_sum_mma_kernel(const ProjectSumArgs a) {
  __shared__ __align__(16) double zs[kPmKT][kPmLd];
  __shared__ __align__(16) double ts[kPmKT][kPmLd];
  __shared__ double ys[kPsBM];
  __shared__ double colacc[4][kPsBN];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wr = warp >> 1, wc = warp & 1;
  const int g = lane >> 2, tq = lane & 3;             // MMA fragment coordinates
  const int S = a.S, d = a.d;
  const int ncoltiles = (S + kPsBN - 1) / kPsBN;
  const int64_t nrowblocks = (a.n + kPsBM - 1) / kPsBM;
  // loader roles: z tile 128 rows x 16 k (8 doubles per thread), theta tile 16 k x 128 cols (8 per thread)
  const int zrow = t >> 1, zhalf = t & 1;
  const int tk = t >> 4, tcol = (t & 15) * 8;

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
          const int c = col0 + tcol + q;
          treg[q] = (k < d && c < S) ? a.thetaT[(size_t)k * S + c] : 0.;
        }
      };
      gload(0);
      for (int k0 = 0; k0 < d; k0 += kPmKT) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) zs[zhalf * 8 + q][zrow] = zreg[q];
#pragma unroll
        for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2*>(&ts[tk][tcol + q]) = make_double2(treg[q], treg[q + 1]);
        __syncthreads();
        if (k0 + kPmKT < d) gload(k0 + kPmKT);
#pragma unroll
        for (int kk = 0; kk < kPmKT; kk += 4) {
          double af[4], bf[8];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) af[mi] = zs[kk + tq][wr * 32 + mi * 8 + g];
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) bf[ni] = ts[kk + tq][wc * 64 + ni * 8 + g];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
      }

      // ---- epilogue: accumulator (mi, ni, e) is row wr*32 + mi*8 + g, column wc*64 + ni*8 + tq*2 + e
      double cs[8][2];
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
            const double v = link_apply<MODEL>(a.sp_tab, lin, y);
            cs[ni][e] += (live && c < S) ? v : 0.;
          }
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
          colacc[wr][wc * 64 + ni * 8 + tq * 2] = cs[ni][0];
          colacc[wr][wc * 64 + ni * 8 + tq * 2 + 1] = cs[ni][1];
        }
      }
      __syncthreads();
      if (t < kPsBN) {
        const double v = colacc[0][t] + colacc[1][t] + colacc[2][t] + colacc[3][t];
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

}  // namespace bcg
