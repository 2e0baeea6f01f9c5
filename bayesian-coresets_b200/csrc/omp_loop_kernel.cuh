// OrthoPursuit build(itrs) as ONE persistent cooperative kernel (snnls.py:31-79 around orthopursuit.py:17-42).
//
// Round 1 ran an OMP iteration as separate launches (scan_kernel, then the single-CTA omp_iteration_kernel): every
// iteration paid the kernel boundaries, the drained HBM pipe and the scan's ramp-up -- ~25 us of a 252 us iteration at
// N = 1e6, S = 256.  Here, as in greedy_loop_kernel:
//   * CTAs 1 .. G-1 are scanning CTAs (scan_cta_body of loop_kernel.cuh): private TMA ring per warp, the tile stream runs
//     ahead across iterations, per-CTA top-2 + lost score, arrive / go counters;
//   * CTA 0 is the CONTROL CTA: all of its 12 warps run the block-wide float64 logic of nnls_logic.h / step_logic.h
//     (selection with the negative direction, warm-started Lawson-Hanson NNLS with Givens removal, monotone check, event
//     log, next residual direction) -- the same code the launch-per-iteration engine runs, so semantics and results are
//     identical; it gives up one SM of scan bandwidth (0.7 %).
// The serial NNLS cannot overlap the next scan (the direction depends on it); what this kernel removes is everything
// around it.  Exactness of the candidate set: as for GIGA / FW, an ambiguous set stops the kernel before the iteration,
// the host runs exact_scan_kernel and relaunches (LoopArgs::cont / use_pre).
#pragma once
#include "loop_kernel.cuh"
#include "nnls_logic.h"
#include "mail_exchange.cuh"

namespace bcg {

constexpr int kOmpLoopThreads = 384;

// copy the scanning CTAs' candidates next to the control CTA (L2 loads: they were written by other SMs during this
// launch) and summarise them like the last CTA of scan_kernel does.  Returns true when the set is ambiguous.
__device__ __forceinline__ bool omp_gather_candidates(const Blk& B, SolverState* st, const LoopArgs& a, int n_ctas) {
  const int n = 2 * n_ctas;
  float top = -INFINITY, lm = -INFINITY;
  for (int i = B.tid; i < n; i += B.nthr) {
    const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(a.cta_cands + i));
    *reinterpret_cast<unsigned long long*>(st->cands + i) = raw;
    if ((uint32_t)(raw >> 32) != kNoRowU) top = fmaxf(top, __uint_as_float((unsigned int)(raw & 0xffffffffull)));
  }
  for (int i = B.tid; i < n_ctas; i += B.nthr) lm = fmaxf(lm, __ldcg(a.cta_lost + i));
  double v[2] = {(double)top, (double)lm};
  // block max through the sum helper's scratch: two rounds of (shuffle max, shared memory)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
    v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], off));
  }
  const int warp = B.tid >> 5, lane = B.tid & 31, nw = B.nthr >> 5;
  if (lane == 0) { B.sred[warp * 2] = v[0]; B.sred[warp * 2 + 1] = v[1]; }
  __syncthreads();
  v[0] = -INFINITY; v[1] = -INFINITY;
  for (int w = 0; w < nw; ++w) { v[0] = fmax(v[0], B.sred[w * 2]); v[1] = fmax(v[1], B.sred[w * 2 + 1]); }
  __syncthreads();
  top = (float)v[0]; lm = (float)v[1];
  const float thr = top - (2e-5f + 1e-5f * fabsf(top));
  double cnt = 0.;
  uint32_t trow = kNoRowU;
  for (int i = B.tid; i < n; i += B.nthr) {
    const ScanCand c = st->cands[i];
    if (c.row != kNoRowU && c.score >= thr) cnt += 1.;
    if (c.row != kNoRowU && c.score == top && c.row < trow) trow = c.row;
  }
  blk_sum<1>(B, &cnt);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) trow = min(trow, __shfl_xor_sync(0xffffffffu, trow, off));
  uint32_t* srow = reinterpret_cast<uint32_t*>(B.sred);
  if (lane == 0) srow[warp] = trow;
  __syncthreads();
  trow = kNoRowU;
  for (int w = 0; w < nw; ++w) trow = min(trow, srow[w]);
  __syncthreads();
  const bool ambiguous = top > -INFINITY && (lm >= thr || cnt > (double)kRescoreMax || st->force_exact);
  if (B.tid == 0) {
    st->scan_top = top;
    st->scan_top_row = trow;
    st->scan_cnt = ambiguous ? 0 : (int)cnt;
  }
  __syncthreads();
  return ambiguous;
}

__device__ void omp_control_cta(const LoopArgs& a, NnlsWork* W, int wide, unsigned char* smem_raw) {
  SolverState* st = a.st;
  LoopCtl* ctl = a.ctl;
  const int n_ctas = (int)gridDim.x - 1;
  double* sred = reinterpret_cast<double*>(smem_raw);
  double* swide = sred + 256;                                  // (nthr / 32) * kWideCols doubles
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, wide ? swide : nullptr};
  if (B.tid == 0 && !a.cont) st->retried = 0;                  // snnls.py:40: local to the build() call
  __syncthreads();
  auto publish = [&](unsigned int it_done, bool stop) {        // all threads' stores -> barrier -> release by thread 0
    __syncthreads();
    if (B.tid == 0) {
      if (stop) *reinterpret_cast<volatile unsigned int*>(&ctl->stop) = 1u;
      __threadfence();
      st_release_gpu_u32(&ctl->go, it_done);
    }
  };
  if (!st->halted) prepare_select(B, st);                      // first residual direction (orthopursuit.py:18)
  publish(1u, false);
  int it = 0, stop_exact = 0;
  for (; it < a.itrs; ++it) {
    if (B.tid < 32) spin_until_ge_gpu_u32(&ctl->arrive, (unsigned int)n_ctas * (unsigned int)(it + 1), 20);
    __syncthreads();
    if (st->halted) break;
    if (!(a.use_pre && it == 0)) {                             // (use_pre: the exact pass already holds the local winner)
      if (omp_gather_candidates(B, st, a, n_ctas)) {
        if (B.tid == 0) st->need_exact = 1;
        stop_exact = 1;
        break;
      }
    }
    omp_iteration(B, st, W, (it + 1 < a.itrs) ? 1 : 0);
    __syncthreads();
    if (st->halted || st->comm_error) break;
    if (it + 1 < a.itrs) publish((unsigned int)(it + 2), false);
  }
  publish((unsigned int)(a.itrs + 2), true);
  if (B.tid == 0) st->iters_done = stop_exact ? it : a.itrs;
}

template <int CH, int LPR, int R, int CH16>
__global__ void __launch_bounds__(kOmpLoopThreads, 1) omp_loop_kernel(const LoopArgs a, NnlsWork* W, int wide) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (blockIdx.x == 0) {
    omp_control_cta(a, W, wide, smem_raw);
    return;
  }
  if ((int)(threadIdx.x >> 5) >= a.wpb) return;                // the scanning CTAs use wpb warps
  scan_cta_body<CH, 1, LPR, R, CH16>(a, smem_raw, (int)blockIdx.x - 1, (int)gridDim.x - 1);
}

}  // namespace bcg
