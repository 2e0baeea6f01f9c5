// K1 as a stand-alone kernel: one pass of the scan core over the N x S matrix, one candidate per
// warp.  Used by the OMP selection step and by the launch-per-iteration engine (BCG_ENGINE=v1);
// GIGA / Frank-Wolfe builds normally run inside the persistent kernel of loop_kernel.cuh.
// HBM-bound: algorithmic traffic 4*N*S bytes per launch, 0.5-1 flop/byte.
#pragma once
#include "scan_core.cuh"
#include "step_logic.h"

namespace bcg {

template <int CH, int NDIR, int LPR, int R>
__global__ void __launch_bounds__(512, 1) scan_kernel(const ScanArgs a) {
  using Core = ScanCore<CH, NDIR, LPR, R>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
  const int64_t GW = (int64_t)gridDim.x * wpb;
  if (*a.skip0 != 0 || *a.skip1 != 0) return;

  const ScanGeom& q = a.g;
  const uint32_t stage_floats = (uint32_t)q.rps * (uint32_t)q.ld;
  float* wbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * q.stages * stage_floats;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem_raw + (size_t)wpb * q.stages * stage_floats * sizeof(float)) + warp * q.stages;
  if (lane == 0) {
    for (int s = 0; s < q.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t policy = l2_policy(q.evict_first);

  const int g = lane & (LPR - 1);
  const int grp = lane / LPR;
  const int nchunk = q.ld >> 2;
  float4 d0[CH];
  float4 d1[CH];
  Core::load_dirs(a.dir, q.ld, nchunk, g, d0, d1);

  const int64_t n_chunks = (q.n_rows + q.rps - 1) / q.rps;
  const int64_t n_my = (gw < n_chunks) ? (n_chunks - gw + GW - 1) / GW : 0;
  const int64_t row_step = GW * q.rps;
  // slot s of the ring always holds chunk k with k % stages == s, so a refill goes into the slot
  // that was just consumed; all ring bookkeeping is incremental (no 64-bit div/mod in the loop)
  const bool leader = lane == 0;
  auto issue = [&](int slot, int64_t row0) {        // executed by all lanes, issued by the leader only
    const int64_t left = q.n_rows - row0;
    const uint32_t nr = (uint32_t)(left < q.rps ? left : q.rps);
    tma_load_rows(&bars[slot], wbuf + (size_t)slot * stage_floats, q.An + (size_t)row0 * q.ld, nr * (uint32_t)q.ld * 4u,
                  policy, leader);
  };
  {
    const int pre = (int)(n_my < q.stages ? n_my : q.stages);
    for (int k = 0; k < pre; ++k) issue(k, (gw + k * GW) * q.rps);
  }

  float best = -INFINITY, lost = -INFINITY;
  uint32_t brow = kNoRowU;
  int slot = 0;
  uint32_t parity = 0;
  int64_t row0 = gw * q.rps;
  const int64_t ahead = (int64_t)q.stages * row_step;
  for (int64_t k = 0; k < n_my; ++k) {
    mbar_wait(&bars[slot], parity);
    const int64_t left = q.n_rows - row0;
    const int nr = (int)(left < q.rps ? left : q.rps);
    const float* tile = wbuf + (size_t)slot * stage_floats;
    for (int b0 = 0; b0 < nr; b0 += Core::RB)
      Core::batch(tile + (size_t)b0 * q.ld, q.ld, nchunk, g, grp, nr - b0, (uint32_t)(row0 + b0), d0, d1, best, brow, lost);
    __syncwarp();   // every lane's shared-memory reads of this stage are complete
    if (k + q.stages < n_my) issue(slot, row0 + ahead);
    row0 += row_step;
    if (++slot == q.stages) { slot = 0; parity ^= 1u; }
  }
  Core::warp_merge_lost(best, brow, lost);
  if (lane == 0) {
    ScanCand c;
    c.score = best;
    c.row = brow;
    a.cands[gw] = c;
    a.lost[gw] = lost;
  }
  // Exactness check of the candidate set, by the LAST CTA to finish: with top = the float32 maximum and the near-tie
  // window of pick_local below it, the float64 arg-max is guaranteed to be among the published candidates unless an
  // unpublished score lies inside the window, or more than kRescoreMax published ones do.  Then need_exact is raised
  // and exact_scan_kernel (launched right after this kernel; it returns at once otherwise) redoes the selection.
  __shared__ unsigned int s_last;
  __shared__ float s_top[32];
  __shared__ int s_cnt[32];
  __shared__ uint32_t s_row[32];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(a.done, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int n = (int)GW;
  float top = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(a.cands + i));
    if ((uint32_t)(raw >> 32) != kNoRowU) top = fmaxf(top, __uint_as_float((unsigned int)(raw & 0xffffffffull)));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, off));
  if (lane == 0) s_top[warp] = top;
  __syncthreads();
  top = -INFINITY;
  for (int w = 0; w < wpb; ++w) top = fmaxf(top, s_top[w]);
  __syncthreads();
  const float thr = top - (2e-5f + 1e-5f * fabsf(top));
  int cnt = 0;
  float lm = -INFINITY;
  uint32_t trow = kNoRowU;                                // lowest row among the candidates that attain the maximum
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(a.cands + i));
    const uint32_t rw = (uint32_t)(raw >> 32);
    const float sc = __uint_as_float((unsigned int)(raw & 0xffffffffull));
    if (rw != kNoRowU && sc >= thr) ++cnt;
    if (rw != kNoRowU && sc == top && rw < trow) trow = rw;
    lm = fmaxf(lm, __ldcg(a.lost + i));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, off));
    trow = min(trow, __shfl_xor_sync(0xffffffffu, trow, off));
  }
  if (lane == 0) { s_cnt[warp] = cnt; s_top[warp] = lm; s_row[warp] = trow; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt = 0; lm = -INFINITY; trow = kNoRowU;
    for (int w = 0; w < wpb; ++w) { cnt += s_cnt[w]; lm = fmaxf(lm, s_top[w]); trow = min(trow, s_row[w]); }
    const bool ambiguous = top > -INFINITY && (lm >= thr || cnt > kRescoreMax || *a.force_exact);
    if (ambiguous) *a.need_exact = 1;
    *a.top = top;
    *a.top_row = trow;
    *a.top_cnt = ambiguous ? 0 : cnt;
    *a.done = 0u;
  }
}

// Exact selection pass (rare): float64 inner products of EVERY local unit row with the float64 direction(s), the
// reference's score (giga.py:33-38 / frankwolfe.py:17), arg-max with the lowest row on ties -- what ndarray.argmax
// returns on float64 scores of the stored rows.  One candidate per CTA into st->exact_cands; the consumer
// (pick_local / the loop kernel's control warp) reduces them.  Returns immediately unless need_exact is set.
__global__ void __launch_bounds__(512) exact_scan_kernel(SolverState* st, int force) {
  if (!force && !st->need_exact) return;
  if (st->halted || st->select_failed) return;
  __shared__ double sd[2 * 1024];
  __shared__ double s_best[16];
  __shared__ long long s_row[16];
  const int S = st->S, ld = st->ld;
  const bool giga = st->alg == BCG_ALG_GIGA;
  for (int i = threadIdx.x; i < (giga ? 2 : 1) * S; i += blockDim.x) sd[i] = st->dir64[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * nw + warp, GW = (int64_t)gridDim.x * nw;
  double best = -INFINITY;
  long long brow = -1;
  for (int64_t row = gw; row < st->n_local; row += GW) {
    const float* x = st->An + (size_t)row * ld;
    double v0 = 0., v1 = 0.;
    for (int s = lane; s < S; s += 32) {
      const double xv = (double)__ldcs(x + s);
      v0 = fma(xv, sd[s], v0);
      if (giga) v1 = fma(xv, sd[S + s], v1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, off);
      v1 += __shfl_xor_sync(0xffffffffu, v1, off);
    }
    const double sc = giga ? giga_score64(v0, v1) : v0;
    if (sc > best) { best = sc; brow = row; }            // rows ascend per warp: the first maximum stays
  }
  if (lane == 0) { s_best[warp] = best; s_row[warp] = brow; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w)
      if (s_row[w] >= 0 && (brow < 0 || s_best[w] > best || (s_best[w] == best && s_row[w] < brow))) { best = s_best[w]; brow = s_row[w]; }
    ExactCand c;
    c.score = best;
    c.row = brow;
    st->exact_cands[blockIdx.x] = c;
  }
}

}  // namespace bcg
