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
    double sum = 0.;
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      acc[j][0] = fast_link<MODEL>(tab, acc[j][0], y);
      acc[j][1] = fast_link<MODEL>(tab, acc[j][1], y);
      sum += acc[j][0] + acc[j][1];
    }
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

// ---------------------------------------------------------------------------------------------------------------
// Two warps per row -- an experiment that did NOT pay off, kept selectable (BCG_PROJ_FAST=2).  With the first link
// table project_fast_kernel ran with issue slots 24 % busy at 16 warps per SM (the 64 registers of per-lane state --
// 16 accumulators + 16 column sums in float64 -- pin it at 128 registers per thread), which looked latency-bound.  It
// was not: doubling the warps doubled the stall cycles per issue and ncu showed the L1 data pipe at 97 % of its
// wavefront rate (gathered table loads + sample-tile LDS); the fix was the one-load link table (softplus_table.h), after
// which one warp per row takes 0.506 ms per 209715 x 512 chunk and this kernel 0.546 ms.  Here a row is shared by a PAIR of warps, each owning half
// of the columns (8 + 8 float64 of state at S = 512), so 32 warps fit on an SM at 64 registers; the two row-wide
// reductions (mean over the S samples, row norm) are completed across the pair through shared memory and an
// mbarrier per exchange (lane 0 of each warp arrives, all lanes wait; the two exchanges of a row alternate, so a slot
// is never overwritten before the partner has read it).  Both warps add the two halves in the same order: they hold
// bit-identical means and norms.  S = 128 * J2 in {128, 256, 512}.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double pair_total(double part, double* slot, uint64_t* bar, int half, int lane, uint32_t parity) {
  const uint32_t addr = pk_smem_u32(bar);
  if (lane == 0) {
    slot[half] = part;
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
  const volatile double* vs = slot;
  return vs[0] + vs[1];
}

constexpr int kPairThreads = 1024;                       // 32 warps = kProjWarps row pairs

template <int J2, int MODEL>
__global__ void __launch_bounds__(kPairThreads, 1) project_pair_kernel(const ProjectArgs a) {
  extern __shared__ __align__(16) double smem[];
  constexpr int S = 128 * J2, H = S / 2;
  __shared__ double xval[kProjWarps][2][2];              // [pair][exchange][half]
  __shared__ __align__(8) uint64_t xbar[kProjWarps][2];
  double* th = smem;                                     // [d][S]
  double* smem_cs = smem + (size_t)a.d * S;              // [pairs][S + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp >> 1, half = warp & 1;
  const int d = a.d;
  const double* __restrict__ tab = a.sp_tab;
  for (int i = threadIdx.x; i < d * S; i += blockDim.x) th[i] = a.theta[i];
  double* coff_s = smem_cs;                              // idle until the flush (see project_fast_kernel)
  if (MODEL == MODEL_LINEAR)
    for (int i = threadIdx.x; i < S; i += blockDim.x) coff_s[i] = a.coff ? a.coff[i] : 0.;
  if (threadIdx.x < 2 * kProjWarps) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pk_smem_u32(&xbar[threadIdx.x >> 1][threadIdx.x & 1])), "r"(2) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double colsum[J2][2];
#pragma unroll
  for (int j = 0; j < J2; ++j) colsum[j][0] = colsum[j][1] = 0.;
  double normsum = 0.;

  const int64_t gp = (int64_t)blockIdx.x * kProjWarps + pair, GP = (int64_t)gridDim.x * kProjWarps;
  auto fetch = [&](int64_t row, double& z, double& y) {
    z = 0.; y = 0.;
    if (row < a.n) {
      const int64_t zr = a.rowidx ? a.rowidx[row] : row;
      if (lane < d) z = a.Z[zr * a.zld + lane];
      if (MODEL == MODEL_POISSON) y = a.Z[zr * a.zld + d];
    }
  };
  double zreg, y;
  fetch(gp, zreg, y);
  uint32_t parity = 0;
  for (int64_t row = gp; row < a.n; row += GP) {
    double znext, ynext;
    fetch(row + GP, znext, ynext);
    double acc[J2][2];
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      if (MODEL == MODEL_LINEAR) {
        const double2 c = reinterpret_cast<const double2*>(coff_s + half * H)[32 * j + lane];
        acc[j][0] = c.x; acc[j][1] = c.y;
      } else {
        acc[j][0] = acc[j][1] = 0.;
      }
    }
    for (int k = 0; k < d; ++k) {
      const double zk = __shfl_sync(0xffffffffu, zreg, k);
      const double2* tk = reinterpret_cast<const double2*>(th + (size_t)k * S + half * H) + lane;
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        const double2 t = tk[32 * j];
        acc[j][0] = fma(zk, t.x, acc[j][0]);
        acc[j][1] = fma(zk, t.y, acc[j][1]);
      }
    }
    double sum = 0.;
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      acc[j][0] = fast_link<MODEL>(tab, acc[j][0], y);
      acc[j][1] = fast_link<MODEL>(tab, acc[j][1], y);
      sum += acc[j][0] + acc[j][1];
    }
    const double mean = pair_total(warp_sum(sum), xval[pair][0], &xbar[pair][0], half, lane, parity) * (1. / (double)S);
    double ss = 0.;
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      acc[j][0] -= mean; acc[j][1] -= mean;
      ss = fma(acc[j][0], acc[j][0], ss);
      ss = fma(acc[j][1], acc[j][1], ss);
      colsum[j][0] += acc[j][0]; colsum[j][1] += acc[j][1];
    }
    ss = pair_total(warp_sum(ss), xval[pair][1], &xbar[pair][1], half, lane, parity);
    parity ^= 1u;
    const double norm = sqrt(ss);
    const double inv = norm > 0. ? 1. / norm : 0.;
    if (a.An) {
      float2* out = reinterpret_cast<float2*>(a.An + (size_t)row * S + half * H) + lane;
#pragma unroll
      for (int j = 0; j < J2; ++j) out[32 * j] = make_float2((float)(acc[j][0] * inv), (float)(acc[j][1] * inv));
      if (lane == 0 && half == 0) a.norms[row] = norm;
    }
    if (lane == 0 && half == 0) {
      normsum += norm;
      if (norm == 0.) atomicAdd(a.zero_rows, 1ull);
    }
    zreg = znext; y = ynext;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < J2; ++j) {
    smem_cs[(size_t)pair * (S + 1) + half * H + 64 * j + 2 * lane] = colsum[j][0];
    smem_cs[(size_t)pair * (S + 1) + half * H + 64 * j + 2 * lane + 1] = colsum[j][1];
  }
  if (lane == 0 && half == 0) smem_cs[(size_t)pair * (S + 1) + S] = normsum;
  __syncthreads();
  for (int s = threadIdx.x; s < S + 1; s += blockDim.x) {
    double t = 0.;
    for (int w = 0; w < kProjWarps; ++w) t += smem_cs[(size_t)w * (S + 1) + s];
    a.partial[(size_t)blockIdx.x * (S + 1) + s] = t;
  }
}

}  // namespace bcg
