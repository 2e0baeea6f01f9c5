// The never-materialising select: one greedy selection pass WITHOUT the N x S matrix.
//
// Every data row's centred log-likelihood vector is recomputed from Z and theta in float64 (projector.py:19-21 around
// model_lr.py:25-32 / model_gaussian.py:4-10 / model_poiss.py:25-38), normalised (giga.py:10-13) and scored against the
// current direction(s) (giga.py:31-38 / frankwolfe.py:17 / orthopursuit.py:19 / sparsevi.py:51); the arg-max (lowest row on
// ties, as ndarray.argmax) is reduced warp -> CTA -> grid, and the LAST CTA to finish re-evaluates the winning row and
// leaves its unit float32 form + norm in the solver state, where the step kernels pick it up exactly like a row of the
// resident matrix (step_logic.h: local_row / pick_local).
//
// Trade: the solver holds 8 N d_in bytes (the raw data) instead of 4 N S -- N = 1e7, d = 10, S = 512: 0.8 GB instead of
// 20.5 GB, so N = 1e8 fits one GPU -- and pays the float64 evaluation of all N S elements per iteration (float64-pipe
// bound, ~10x the HBM scan).  Not the default; `HilbertCoreset(..., materialize=False)`.
// Row evaluation is the audit scorer's (audit_kernel.cuh): the same float64 libdevice arithmetic as the reference.
#pragma once
#include "audit_kernel.cuh"
#include "bcg_state.h"

namespace bcg {

struct LazyArgs {
  AuditArgs a;          // Z, model, samples (scores / norms / colsum outputs unused)
  SolverState* st;
};

template <int J>
__global__ void __launch_bounds__(256) lazy_select_kernel(const LazyArgs L) {
  SolverState* st = L.st;
  if (st->halted || st->select_failed) return;
  __shared__ double sd[2 * 1024];
  __shared__ double s_best[8];
  __shared__ long long s_row[8];
  __shared__ unsigned int s_last;
  const AuditArgs& a = L.a;
  const int S = a.S, ld = st->ld;
  const bool giga = st->alg == BCG_ALG_GIGA;
  for (int i = threadIdx.x; i < (giga ? 2 : 1) * S; i += blockDim.x) sd[i] = st->dir64[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * nw + warp, GW = (int64_t)gridDim.x * nw;
  double best = -INFINITY;
  long long brow = -1;
  for (int64_t row = gw; row < a.n; row += GW) {
    double v[J];
    const double norm = audit_row<J>(a, row, lane, v);
    const double sc = audit_score<J>(v, norm, sd, S, giga, lane);
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
    __threadfence();
    s_last = (atomicAdd(&st->scan_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: grid winner, then its unit row ---------------------------------------------------------
  double ks = -INFINITY;
  long long kr = -1;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const ExactCand x = st->exact_cands[i];
    if (x.row >= 0 && (kr < 0 || x.score > ks || (x.score == ks && x.row < kr))) { ks = x.score; kr = x.row; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double s2 = __shfl_xor_sync(0xffffffffu, ks, off);
    const long long r2 = __shfl_xor_sync(0xffffffffu, kr, off);
    if (r2 >= 0 && (kr < 0 || s2 > ks || (s2 == ks && r2 < kr))) { ks = s2; kr = r2; }
  }
  __syncthreads();
  if (lane == 0) { s_best[warp] = ks; s_row[warp] = kr; }
  __syncthreads();
  if (warp == 0) {
    ks = s_best[0]; kr = s_row[0];
    for (int w = 1; w < nw; ++w)
      if (s_row[w] >= 0 && (kr < 0 || s_best[w] > ks || (s_best[w] == ks && s_row[w] < kr))) { ks = s_best[w]; kr = s_row[w]; }
    if (kr >= 0) {
      double v[J];
      const double norm = audit_row<J>(a, kr, lane, v);
      const double inv = norm > 0. ? 1. / norm : 0.;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int s = lane + 32 * j;
        if (s < ld) st->wrow[s] = (s < S) ? (float)(v[j] * inv) : 0.f;
      }
      if (lane == 0) st->wnorm = norm;
    }
    if (lane == 0) {
      st->fused_row = kr;
      st->fused_score = ks;
      st->scan_done = 0u;
    }
  }
}

}  // namespace bcg
