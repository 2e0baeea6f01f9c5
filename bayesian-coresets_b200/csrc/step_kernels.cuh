// K2 launch wrappers (single thread block) around step_logic.h, and the NVLink peer-memory
// candidate exchange used when the N axis is sharded over the GPUs of one node.
#pragma once
#include <cuda_runtime.h>
#include "step_logic.h"
#include "nnls_logic.h"
#include "sync_ptx.cuh"
#include "mail_exchange.cuh"

namespace bcg {

constexpr int kStepThreads = 512;

// do_finish: complete the iteration whose scan just ran; do_prep: produce the next scan
// direction; reset_retry: start of a build() call (snnls.py:40)
__global__ void __launch_bounds__(kStepThreads, 1)
step_kernel(SolverState* st, int do_finish, int do_prep, int reset_retry) {
  __shared__ double sred[256];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, nullptr};
  if (reset_retry) {
    if (threadIdx.x == 0) st->retried = 0;
    __syncthreads();
  }
  if (st->halted) return;
  if (do_finish) finish_iteration(B, st);
  __syncthreads();
  if (st->halted) return;
  if (do_prep) prepare_select(B, st);
}

__global__ void __launch_bounds__(kStepThreads, 1) omp_select_kernel(SolverState* st, int64_t* f_out) {
  __shared__ double sred[256];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, nullptr};
  const int64_t f = omp_select(B, st);
  if (threadIdx.x == 0) *f_out = f;
}

// local argmax of <unit row, dir> with float64 re-scoring; no state change (SparseVI selection)
__global__ void __launch_bounds__(kStepThreads, 1) probe_kernel(SolverState* st, int64_t* f_out, double* score_out) {
  __shared__ double sred[256];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, nullptr};
  uint32_t lrow; double sc;
  pick_local(B, st, true, &lrow, &sc);
  if (threadIdx.x == 0) {
    *f_out = (lrow == kNoRow) ? -1 : st->row_offset + (int64_t)lrow;
    *score_out = sc;
  }
}

// OMP iteration on the device (orthopursuit.py:17-42 inside snnls.py:41-78): selection (+ w[f] = 1), NNLS
// re-solve on the active set, monotone-error check with revert, event log, retry / latch, and -- when another
// iteration follows -- the residual direction for its scan (orthopursuit.py:18), so that an iteration is two
// launches (scan + this kernel).  act_w_new keeps the weights from before the iteration for the revert.
// wide: use the warp-split K x S products (blk_combine through 32 KB of shared scratch).
__global__ void __launch_bounds__(kStepThreads, 1) omp_iteration_kernel(SolverState* st, NnlsWork* W, int wide, int prep_next) {
  __shared__ double sred[256];
  __shared__ double swide[(kStepThreads / 32) * kWideCols];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, wide ? swide : nullptr};
  omp_iteration(B, st, W, prep_next);
}

__global__ void __launch_bounds__(kStepThreads, 1) nnls_kernel(SolverState* st, NnlsWork* W, int from_scratch, int wide) {
  __shared__ double sred[256];
  __shared__ double swide[(kStepThreads / 32) * kWideCols];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, wide ? swide : nullptr};
  nnls_solve(B, st, W, from_scratch);
}

__global__ void __launch_bounds__(kStepThreads, 1) refresh_kernel(SolverState* st) {
  __shared__ double sred[256];
  Blk B{(int)threadIdx.x, (int)blockDim.x, sred, nullptr};
  refresh_iterate(B, st);
}

}  // namespace bcg
