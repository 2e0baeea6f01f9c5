// Shared device core of the N x S scan (used by the stand-alone scan kernel and by the persistent
// greedy-loop kernel): TMA bulk-copy / mbarrier helpers and the register-tiled row-batch
// inner product + score + running arg-max.
//
// Replaces (reference, float64 NumPy/BLAS):
//   snnls/giga.py:31-38          An.T.dot([cdir|xw]) + mask + sqrt + divide + argmax   (NDIR = 2)
//   snnls/frankwolfe.py:17       An.T.dot(residual).argmax()                            (NDIR = 1)
//   snnls/orthopursuit.py:19-26  same positive-direction scan                           (NDIR = 1)
//
// Work decomposition.  A warp owns a private ring of shared-memory stages filled by TMA bulk
// copies (contiguous groups of rows).  Inside a stage it works on batches of RB = R * (32/LPR)
// rows: LPR lanes cooperate on one row (lane g owns float4 chunks g, g+LPR, ...; the direction
// vectors for exactly those chunks live in registers), 32/LPR row groups run side by side, and
// every lane accumulates R rows at once (R independent FMA chains per direction).  The R x LPR
// partial sums are reduced with a recursive-halving shuffle network (R-1 + log2(LPR/R) shuffles
// per direction for R rows instead of R*log2(LPR)), after which each lane holds the complete
// inner products of one row, evaluates the score and keeps a running best -- first maximum wins,
// as ndarray.argmax.  Measured motivation: the first version (one row at a time, runtime LPR)
// executed 150 warp-instructions per 1 KB row and was issue-bound at 55 % of HBM bandwidth
// (profiles/r01_v1_scan_N1e6_S256_full.csv).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "bcg_state.h"
#include "filter_bounds.h"
#include "kernel_args.h"

namespace bcg {

constexpr uint32_t kNoRowU = 0xffffffffu;

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

// Arm the stage barrier with the byte count, then start the bulk copy global -> shared (UBLKCP).
// Called by ALL lanes with identical arguments; only the lane with `leader` set issues, through
// PTX predicates rather than a branch.  (A lane-0 `if` around this, with loop-carried lane-0
// state, left lane 0 permanently split from lanes 1-31 in the persistent kernel: every
// instruction of the scan executed twice and the shuffles took the divergent slow path --
// profiles/r01_analysis.md.)
__device__ __forceinline__ void tma_load_rows(uint64_t* bar, float* dst, const float* src, uint32_t bytes,
                                              uint64_t policy, bool leader) {
  const uint32_t b = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t"
      "}"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(b), "l"(policy), "r"((uint32_t)leader)
      : "memory");
}

__device__ __forceinline__ uint64_t l2_policy(int evict_first) {
  uint64_t policy;
  if (evict_first)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}

// (score descending, row ascending); "no row" never beats a real row
__device__ __forceinline__ bool cand_better(float s2, uint32_t r2, float s1, uint32_t r1) {
  if (r2 == kNoRowU) return false;
  if (r1 == kNoRowU) return true;
  return s2 > s1 || (s2 == s1 && r2 < r1);
}

// recursive-halving reduction of V per-lane values over the lanes {g ^ OFF, g ^ OFF/2, ...}
template <int V, int OFF>
struct HalvingReduce {
  static __device__ __forceinline__ void run(float* a, int g, int& rid) {
    if constexpr (OFF >= 1) {
      if constexpr (V > 1) {
        const bool hi = (g & OFF) != 0;
#pragma unroll
        for (int t = 0; t < V / 2; ++t) {
          const float keep = hi ? a[t + V / 2] : a[t];
          const float send = hi ? a[t] : a[t + V / 2];
          a[t] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if (hi) rid += V / 2;
        HalvingReduce<V / 2, OFF / 2>::run(a, g, rid);
      } else {
        a[0] += __shfl_xor_sync(0xffffffffu, a[0], OFF);
        HalvingReduce<1, OFF / 2>::run(a, g, rid);
      }
    }
  }
};

constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x / 2); }
constexpr int cmin(int a, int b) { return a < b ? a : b; }

template <int CH, int NDIR, int LPR, int R>
struct ScanCore {
  static constexpr int NGRP = 32 / LPR;
  static constexpr int RB = R * NGRP;                          // rows per batch
  static constexpr int HALVINGS = cmin(ilog2c(R), ilog2c(LPR));
  static constexpr int VREM = R >> HALVINGS;                   // complete rows per lane after the reduce

  static __device__ __forceinline__ void load_dirs(const float* dir, int ld, int nchunk, int g, float4 (&d0)[CH],
                                                   float4 (&d1)[CH]) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int c = g + j * LPR;
      d0[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      d1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nchunk) {
        d0[j] = __ldcg(reinterpret_cast<const float4*>(dir) + c);
        if (NDIR == 2) d1[j] = __ldcg(reinterpret_cast<const float4*>(dir + ld) + c);
      }
    }
  }

  // one batch: rows [0, nvalid) of `tile` (RB rows capacity, stride ld floats); global row index of
  // tile row 0 is row_base
  static __device__ __forceinline__ void batch(const float* tile, int ld, int nchunk, int g, int grp, int nvalid,
                                               uint32_t row_base, const float4 (&d0)[CH], const float4 (&d1)[CH],
                                               float& best, uint32_t& brow, float& lost) {
    float a0[R];
    float a1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
    const float4* base = reinterpret_cast<const float4*>(tile + (size_t)grp * ld);
    const int rstride = (NGRP * ld) >> 2;                      // float4 stride between this group's rows
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int c = g + j * LPR;
      if (c < nchunk) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 x = base[(size_t)r * rstride + c];
          a0[r] = fmaf(x.x, d0[j].x, a0[r]); a0[r] = fmaf(x.y, d0[j].y, a0[r]);
          a0[r] = fmaf(x.z, d0[j].z, a0[r]); a0[r] = fmaf(x.w, d0[j].w, a0[r]);
          if (NDIR == 2) {
            a1[r] = fmaf(x.x, d1[j].x, a1[r]); a1[r] = fmaf(x.y, d1[j].y, a1[r]);
            a1[r] = fmaf(x.z, d1[j].z, a1[r]); a1[r] = fmaf(x.w, d1[j].w, a1[r]);
          }
        }
      }
    }
    int rid = 0;
    HalvingReduce<R, LPR / 2>::run(a0, g, rid);
    if (NDIR == 2) { int rid2 = 0; HalvingReduce<R, LPR / 2>::run(a1, g, rid2); }
#pragma unroll
    for (int t = 0; t < VREM; ++t) {
      const int rin = (rid + t) * NGRP + grp;                  // row within the batch
      float score;
      if (NDIR == 2) {
        // giga.py:33-38 in float32 (candidate generation; near ties are re-scored in float64)
        const float den = 1.f - a1[t] * a1[t];
        score = (a1[t] > -1.f && den > 0.f) ? a0[t] * rsqrtf(den) : 0.f;
      } else {
        score = a0[t];
      }
      // running best per lane; `lost` = the best score this lane saw and does NOT carry (exactness check of the
      // candidate set: see SolverState::cand_lost)
      if (rin < nvalid) {
        if (score > best) { lost = fmaxf(lost, best); best = score; brow = row_base + (uint32_t)rin; }
        else lost = fmaxf(lost, score);
      }
    }
  }

  // merge the per-lane running bests of a warp and fold what the merge drops into `lost` (lanes that scored the
  // same row -- LPR > R leaves several lanes with the complete sums of one row -- are not "dropped")
  static __device__ __forceinline__ void warp_merge_lost(float& best, uint32_t& brow, float& lost) {
    const float mine = best;
    const uint32_t myrow = brow;
    warp_merge(best, brow);
    float l = lost;
    if (myrow != brow && myrow != kNoRowU) l = fmaxf(l, mine);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, off));
    lost = l;
  }

  // merge the per-lane running bests of a warp
  static __device__ __forceinline__ void warp_merge(float& best, uint32_t& brow) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, best, off);
      const uint32_t r2 = __shfl_xor_sync(0xffffffffu, brow, off);
      if (cand_better(s2, r2, best, brow)) { best = s2; brow = r2; }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// float16 pre-filter pass (filter_bounds.h): the same register-tiled row batch over a float16 copy of the rows.
// All 32 lanes cooperate on one row (lane g owns the 8-element groups g, g + 32, ...; one 16-byte shared-memory load
// each); R rows run side by side.  Instead of a running arg-max the pass keeps, per row group (= ring stage), the largest
// upper bound of the float32 score, and per warp the largest lower bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_sortable(float x) {          // monotone float -> int (no NaN)
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float sortable_float(int i) {
  return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max_f(float x) {
  return sortable_float(__reduce_max_sync(0xffffffffu, float_sortable(x)));
}

template <int CH16, int NDIR, int R>
struct Filter16Core {
  static constexpr int RB = R;                                   // rows per batch (LPR = 32)
  static constexpr int HALVINGS = cmin(ilog2c(R), 5);
  static constexpr int VREM = R >> HALVINGS;

  // direction registers for this lane's element groups + |d|_2 of each direction (warp-uniform)
  static __device__ __forceinline__ void load_dirs(const float* dir, int ld, int S, int g, float (&d0)[CH16][8],
                                                   float (&d1)[CH16][8], float* n0, float* n1) {
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < CH16; ++j) {
      const int e0 = 8 * (g + 32 * j);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e = e0 + k;
        d0[j][k] = (e < S) ? __ldcg(dir + e) : 0.f;
        d1[j][k] = (NDIR == 2 && e < S) ? __ldcg(dir + ld + e) : 0.f;
        q0 = fmaf(d0[j][k], d0[j][k], q0);
        q1 = fmaf(d1[j][k], d1[j][k], q1);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      q0 += __shfl_xor_sync(0xffffffffu, q0, off);
      q1 += __shfl_xor_sync(0xffffffffu, q1, off);
    }
    *n0 = sqrtf(q0);
    *n1 = sqrtf(q1);
  }

  // one batch of R rows of `tile` (stride ld16 halves); folds the rows' bounds into ub_max / lb_max (per lane)
  static __device__ __forceinline__ void batch(const __half* tile, int ld16, int ngroups, int g, int nvalid,
                                               const float (&d0)[CH16][8], const float (&d1)[CH16][8], float e0, float e1,
                                               float& ub_max, float& lb_max) {
    float a0[R];
    float a1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { a0[r] = 0.f; a1[r] = 0.f; }
    const uint4* base = reinterpret_cast<const uint4*>(tile);
    const int rstride = ld16 >> 3;                               // uint4 stride between rows
#pragma unroll
    for (int j = 0; j < CH16; ++j) {
      const int c = g + 32 * j;
      if (c < ngroups) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 x = base[(size_t)r * rstride + c];
          const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&x.x));
          const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&x.y));
          const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&x.z));
          const float2 p3 = __half22float2(*reinterpret_cast<const __half2*>(&x.w));
          a0[r] = fmaf(p0.x, d0[j][0], a0[r]); a0[r] = fmaf(p0.y, d0[j][1], a0[r]);
          a0[r] = fmaf(p1.x, d0[j][2], a0[r]); a0[r] = fmaf(p1.y, d0[j][3], a0[r]);
          a0[r] = fmaf(p2.x, d0[j][4], a0[r]); a0[r] = fmaf(p2.y, d0[j][5], a0[r]);
          a0[r] = fmaf(p3.x, d0[j][6], a0[r]); a0[r] = fmaf(p3.y, d0[j][7], a0[r]);
          if (NDIR == 2) {
            a1[r] = fmaf(p0.x, d1[j][0], a1[r]); a1[r] = fmaf(p0.y, d1[j][1], a1[r]);
            a1[r] = fmaf(p1.x, d1[j][2], a1[r]); a1[r] = fmaf(p1.y, d1[j][3], a1[r]);
            a1[r] = fmaf(p2.x, d1[j][4], a1[r]); a1[r] = fmaf(p2.y, d1[j][5], a1[r]);
            a1[r] = fmaf(p3.x, d1[j][6], a1[r]); a1[r] = fmaf(p3.y, d1[j][7], a1[r]);
          }
        }
      }
    }
    int rid = 0;
    HalvingReduce<R, 16>::run(a0, g, rid);
    if (NDIR == 2) { int rid2 = 0; HalvingReduce<R, 16>::run(a1, g, rid2); }
#pragma unroll
    for (int t = 0; t < VREM; ++t) {
      const int rin = rid + t;
      float lb, ub;
      if (NDIR == 2) filter_bounds_giga(a0[t], a1[t], e0, e1, &lb, &ub);
      else filter_bounds_lin(a0[t], e0, &lb, &ub);
      if (rin < nvalid) {
        if (ub == ub) ub_max = fmaxf(ub_max, ub);                // a NaN score is ignored by the float32 scan as well
        if (lb == lb) lb_max = fmaxf(lb_max, lb);
      }
    }
  }
};

}  // namespace bcg
