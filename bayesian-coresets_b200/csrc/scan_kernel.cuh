// K1: the fused residual-row inner product + score + argmax scan over the N x S matrix.
//
// Replaces (reference, float64 NumPy/BLAS):
//   snnls/giga.py:31-38          An.T.dot([cdir|xw]) + mask + sqrt + divide + argmax   (NDIR = 2)
//   snnls/frankwolfe.py:17       An.T.dot(residual).argmax()                            (NDIR = 1)
//   snnls/orthopursuit.py:19-26  same positive-direction scan                           (NDIR = 1)
//
// HBM-bound streaming kernel (algorithmic traffic 4*N*S bytes per launch, 0.5-1 flop/byte):
//  * every warp is an independent streaming engine with a private ring of `stages` shared-memory
//    buffers; lane 0 issues one TMA bulk copy (cp.async.bulk -> UBLKCP) per stage for a
//    contiguous group of rows and all lanes wait on the stage's mbarrier -- no CTA-wide barriers
//    in the steady state, (warps * stages * stage_bytes) bytes in flight per SM;
//  * the 1-2 direction vectors live in registers (lane l owns float4 chunks l, l+lpr, ...), rows
//    are read from shared memory as float4, reduced with warp shuffles over `lpr` lanes;
//  * each warp keeps its best (score, row) -- first maximum wins, as ndarray.argmax -- and writes
//    ONE candidate; the step kernel (K2) resolves the global winner and re-scores near ties in
//    float64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bcg_state.h"

namespace bcg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// arm the stage barrier with the byte count, then start the bulk copy global -> shared
__device__ __forceinline__ void tma_load_rows(uint64_t* bar, float* dst, const float* src, uint32_t bytes,
                                              uint64_t policy) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(b), "l"(policy)
      : "memory");
}

struct ScanArgs {
  const float* An;        // n_rows x ld, unit rows
  const float* dir;       // NDIR x ld
  ScanCand* cands;        // gridDim.x * warps_per_block
  const int32_t* skip0;   // scan is skipped when *skip0 or *skip1 is non-zero (halted / select failed)
  const int32_t* skip1;
  int64_t n_rows;
  int32_t ld;             // floats per row (multiple of 4)
  int32_t lpr;            // lanes per row (power of two, <= 32)
  int32_t rps;            // rows per stage
  int32_t stages;
  int32_t evict_first;    // L2 policy of the streaming loads
};

template <int CH, int NDIR>
__global__ void __launch_bounds__(512, 1) scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
  const int64_t GW = (int64_t)gridDim.x * wpb;

  if (*a.skip0 != 0 || *a.skip1 != 0) return;

  const uint32_t stage_floats = (uint32_t)a.rps * (uint32_t)a.ld;
  float* wbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * a.stages * stage_floats;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem_raw + (size_t)wpb * a.stages * stage_floats * sizeof(float)) + warp * a.stages;

  if (lane == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  uint64_t policy;
  if (a.evict_first)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));

  const int lpr = a.lpr;
  const int g = lane & (lpr - 1);      // lane within the row group
  const int grp = lane / lpr;          // which row of the batch this lane works on
  const int ngrp = 32 / lpr;
  const int nchunk = a.ld >> 2;

  // direction vectors -> registers
  float4 d0[CH];
  float4 d1[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int c = g + j * lpr;
    d0[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    d1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nchunk) {
      d0[j] = reinterpret_cast<const float4*>(a.dir)[c];
      if (NDIR == 2) d1[j] = reinterpret_cast<const float4*>(a.dir + a.ld)[c];
    }
  }

  const int64_t n_chunks = (a.n_rows + a.rps - 1) / a.rps;
  const int64_t n_my = (gw < n_chunks) ? (n_chunks - gw + GW - 1) / GW : 0;

  auto issue = [&](int64_t k) {
    const int st = (int)(k % a.stages);
    const int64_t row0 = (gw + k * GW) * a.rps;
    const int64_t left = a.n_rows - row0;
    const uint32_t nr = (uint32_t)(left < a.rps ? left : a.rps);
    tma_load_rows(&bars[st], wbuf + (size_t)st * stage_floats, a.An + (size_t)row0 * a.ld,
                  nr * (uint32_t)a.ld * 4u, policy);
  };

  if (lane == 0) {
    const int64_t pre = n_my < a.stages ? n_my : a.stages;
    for (int64_t k = 0; k < pre; ++k) issue(k);
  }

  float best = -INFINITY;
  uint32_t brow = 0xffffffffu;

  for (int64_t k = 0; k < n_my; ++k) {
    const int st = (int)(k % a.stages);
    const uint32_t parity = (uint32_t)((k / a.stages) & 1);
    mbar_wait(&bars[st], parity);

    const int64_t row0 = (gw + k * GW) * a.rps;
    const int64_t left = a.n_rows - row0;
    const int nr = (int)(left < a.rps ? left : a.rps);
    const float* tile = wbuf + (size_t)st * stage_floats;

#pragma unroll 2
    for (int rb = 0; rb < nr; rb += ngrp) {
      const int r = rb + grp;
      const float4* rowp = reinterpret_cast<const float4*>(tile + (size_t)(r < a.rps ? r : 0) * a.ld);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int c = g + j * lpr;
        if (c < nchunk) {
          const float4 x = rowp[c];
          a0 = fmaf(x.x, d0[j].x, a0); a0 = fmaf(x.y, d0[j].y, a0);
          a0 = fmaf(x.z, d0[j].z, a0); a0 = fmaf(x.w, d0[j].w, a0);
          if (NDIR == 2) {
            a1 = fmaf(x.x, d1[j].x, a1); a1 = fmaf(x.y, d1[j].y, a1);
            a1 = fmaf(x.z, d1[j].z, a1); a1 = fmaf(x.w, d1[j].w, a1);
          }
        }
      }
      for (int off = lpr >> 1; off > 0; off >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, off);
        if (NDIR == 2) a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      }
      float score;
      if (NDIR == 2) {
        // giga.py:33-38 in float32 (candidate generation; K2 re-scores near ties in float64)
        const float den = 1.f - a1 * a1;
        score = (a1 > -1.f && den > 0.f) ? a0 * rsqrtf(den) : 0.f;
      } else {
        score = a0;
      }
      if (r < nr && score > best) { best = score; brow = (uint32_t)(row0 + r); }
    }
    __syncwarp();   // every lane's shared-memory reads of this stage are complete
    if (lane == 0 && k + a.stages < n_my) issue(k + a.stages);
  }

  // merge the row groups of the warp (ties -> lowest row)
  for (int off = lpr; off < 32; off <<= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, best, off);
    const uint32_t r2 = __shfl_xor_sync(0xffffffffu, brow, off);
    if (s2 > best || (s2 == best && r2 < brow)) { best = s2; brow = r2; }
  }
  if (lane == 0) {
    ScanCand c;
    c.score = best;
    c.row = brow;
    a.cands[gw] = c;
  }
}

}  // namespace bcg
