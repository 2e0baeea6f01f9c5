// K3, specialised: the common shapes of the materialising projection -- S = 64 * J2 in {64, 128, 256, 512}, the whole
// d x S sample tile resident in shared memory (d <= 32), table-driven links -- without the generality tax of
// project_kernel<J>.  Same reference lines (projector.py:19-21 around model_lr.py:25-32 / model_gaussian.py:4-10 /
// model_poiss.py:25-38, giga.py:10-13, hilbert.py:24), same outputs and the same float64 arithmetic.
//
// What the ncu source view of project_kernel<16> charged per matrix element at d = 10 (profiles/r01b_project_lr_*:
// 124 warp instructions per element, issue slots 37 % busy) and what this kernel does instead:
//   * inner product: one LDS.64 + predicate per FMA (32 instr)  -> lane owns column PAIRS (2 lane + 64 j, +1): one
//     conflict-free LDS.128 per two FMAs, no bounds checks (S is a multiple of 64 by construction)
//   * model / table dispatch per element                         -> compile-time MODEL, table pointer in a register
//   * float32 store per element with a bounds check               -> one 8-byte store per column pair (a warp writes
//     256 contiguous bytes per instruction)
//   * the next row's z is fetched before the current row's link is evaluated (the loads have a row's worth of math
//     to hide behind)
// Bytes per row: 8 d_in read + 4 S + 8 written; bound: float64 pipe + issue (about 50 instructions per element), not HBM.
#pragma once
#include <cuda_fp16.h>
#include "project_kernels.cuh"

namespace bcg {

template <int MODEL>
__device__ __forceinline__ double fast_link(const double* tab, double lin, double y) {
  if (MODEL == MODEL_LR) return lr_link_fast(tab, lin);
  if (MODEL == MODEL_POISSON) return poisson_link_fast(tab, lin, y);
  return lin;
}

template <int J2, int MODEL>
__global__ void __launch_bounds__(kProjWarps * 32, 1) project_fast_kernel(const ProjectArgs a) {
  extern __shared__ __align__(16) double smem[];
  constexpr int S = 64 * J2;
  double* th = smem;                                     // [d][S]
  double* smem_cs = smem + (size_t)a.d * S;              // [warps][S + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = a.d;
  const double* __restrict__ tab = a.sp_tab;
  for (int i = threadIdx.x; i < d * S; i += blockDim.x) th[i] = a.theta[i];
  // per-column offsets of the linear (Gaussian) model: kept in the column-sum scratch, which is idle until the flush
  double* coff_s = smem_cs;
  if (MODEL == MODEL_LINEAR)
    for (int i = threadIdx.x; i < S; i += blockDim.x) coff_s[i] = a.coff ? a.coff[i] : 0.;
  __syncthreads();

  double colsum[J2][2];
#pragma unroll
  for (int j = 0; j < J2; ++j) colsum[j][0] = colsum[j][1] = 0.;
  double normsum = 0.;

  const int64_t gw = (int64_t)blockIdx.x * kProjWarps + warp, GW = (int64_t)gridDim.x * kProjWarps;
  auto fetch = [&](int64_t row, double& z, double& y) {
    z = 0.; y = 0.;
    if (row < a.n) {
      const int64_t zr = a.rowidx ? a.rowidx[row] : row;
      if (lane < d) z = a.Z[zr * a.zld + lane];
      if (MODEL == MODEL_POISSON) y = a.Z[zr * a.zld + d];
    }
  };
  double zreg, y;
  fetch(gw, zreg, y);
  for (int64_t row = gw; row < a.n; row += GW) {
    double znext, ynext;
    fetch(row + GW, znext, ynext);                       // in flight during this row's math
    double acc[J2][2];
    auto contract = [&]() {
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        if (MODEL == MODEL_LINEAR) {
          const double2 c = reinterpret_cast<const double2*>(coff_s)[32 * j + lane];
          acc[j][0] = c.x; acc[j][1] = c.y;
        } else {
          acc[j][0] = acc[j][1] = 0.;
        }
      }
      for (int k = 0; k < d; ++k) {
        const double zk = __shfl_sync(0xffffffffu, zreg, k);
        const double2* tk = reinterpret_cast<const double2*>(th + (size_t)k * S) + lane;
#pragma unroll
        for (int j = 0; j < J2; ++j) {
          const double2 t = tk[32 * j];
          acc[j][0] = fma(zk, t.x, acc[j][0]);
          acc[j][1] = fma(zk, t.y, acc[j][1]);
        }
      }
    };
    contract();
    double sum = 0.;
    if (MODEL != MODEL_LINEAR) {
      // branch-free links first (independent chains the compiler can interleave, softplus_table.h); a row with an
      // argument beyond their range (|lin| > 37: rare) is contracted again and evaluated with the branching links
      bool tail = false;
#pragma unroll
      for (int j = 0; j < J2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double lin = acc[j][e];
          tail |= link_needs_tail(lin);
          acc[j][e] = MODEL == MODEL_LR ? lr_link_nb(tab, lin) : poisson_link_nb(tab, lin, y);
        }
      if (__any_sync(0xffffffffu, tail)) {
        contract();
#pragma unroll
        for (int j = 0; j < J2; ++j) {
          acc[j][0] = fast_link<MODEL>(tab, acc[j][0], y);
          acc[j][1] = fast_link<MODEL>(tab, acc[j][1], y);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < J2; ++j) sum += acc[j][0] + acc[j][1];
    const double mean = warp_sum(sum) * (1. / (double)S);   // projector.py:21 (S is a power of two: exact reciprocal)
    double ss = 0.;
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      acc[j][0] -= mean; acc[j][1] -= mean;
      ss = fma(acc[j][0], acc[j][0], ss);
      ss = fma(acc[j][1], acc[j][1], ss);
      colsum[j][0] += acc[j][0]; colsum[j][1] += acc[j][1];
    }
    ss = warp_sum(ss);
    const double norm = sqrt(ss);
    const double inv = norm > 0. ? 1. / norm : 0.;
    if (a.An) {
      float2* out = reinterpret_cast<float2*>(a.An + (size_t)row * S) + lane;
#pragma unroll
      for (int j = 0; j < J2; ++j) out[32 * j] = make_float2((float)(acc[j][0] * inv), (float)(acc[j][1] * inv));
      if (a.An16) {                              // the pre-filter's copy: fl16 of the float32 value just stored
        __half2* o16 = reinterpret_cast<__half2*>(a.An16 + (size_t)row * a.ld16) + lane;
#pragma unroll
        for (int j = 0; j < J2; ++j) o16[32 * j] = __floats2half2_rn((float)(acc[j][0] * inv), (float)(acc[j][1] * inv));
      }
      if (lane == 0) a.norms[row] = norm;
    }
    if (lane == 0) {
      normsum += norm;
      if (norm == 0.) atomicAdd(a.zero_rows, 1ull);
    }
    zreg = znext; y = ynext;
  }
  // block-level reduction of the per-warp column sums into partial[blockIdx][S + 1] (fixed order)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J2; ++j) {
    smem_cs[(size_t)warp * (S + 1) + 64 * j + 2 * lane] = colsum[j][0];
    smem_cs[(size_t)warp * (S + 1) + 64 * j + 2 * lane + 1] = colsum[j][1];
  }
  if (lane == 0) smem_cs[(size_t)warp * (S + 1) + S] = normsum;
  __syncthreads();
  for (int s = threadIdx.x; s < S + 1; s += blockDim.x) {
    double t = 0.;
    for (int w = 0; w < kProjWarps; ++w) t += smem_cs[(size_t)w * (S + 1) + s];
    a.partial[(size_t)blockIdx.x * (S + 1) + s] = t;
  }
}

// (A two-warps-per-row variant -- half the per-lane state, 32 warps per SM, row reductions completed across the pair
// through shared memory and an mbarrier -- was measured at 0.546 ms per 209715 x 512 chunk against 0.506 ms for this kernel and
// removed: the bound was the L1 data pipe, not latency.  The DMMA kernel of project_mma_kernel.cuh is what relieved it.)

}  // namespace bcg
