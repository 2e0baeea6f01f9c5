// K2 launch wrappers (single thread block) around step_logic.h, and the NVLink peer-memory
// candidate exchange used when the N axis is sharded over the GPUs of one node.
#pragma once
#include <cuda_runtime.h>
#include "step_logic.h"
#include "nnls_logic.h"
#include "sync_ptx.cuh"

namespace bcg {

constexpr int kStepThreads = 512;

// Every rank posts (float64 score, global index, norm, unit row) of its best local row into slot
// [parity][rank] of EVERY rank's mailbox with plain peer stores over NVLink, publishes it with a
// system-scope release of the sequence number, then waits (acquire) for the `world` slots of its
// own mailbox and picks the global winner: max score, ties -> lowest global index.  One fused
// all-gather per greedy iteration, no host involvement, no NCCL launch on the critical path.
// Slots are double-buffered on the parity of the sequence number: a rank can only be one
// exchange ahead of its slowest peer, so a slot is never overwritten while it is being read.
__device__ void mail_exchange(const Blk& B, SolverState* st, uint32_t lrow, double lscore, int64_t* f,
                              double* norm, const float** row) {
  const int W = st->world, me = st->rank, ld = st->ld;
  const unsigned long long seq = st->seq + 1ull;
  const int par = (int)(seq & 1ull);
  const int64_t sb = st->mail_slot_bytes;
  const bool have = lrow != kNoRow;
  const float* src = nullptr;
  double nrm = 0.;
  if (have) local_row(st, lrow, &src, &nrm);
  const int64_t gidx = have ? st->row_offset + (int64_t)lrow : -1;

  for (int p = 0; p < W; ++p) {
    unsigned char* slot = st->mail_peer[p] + (int64_t)(par * W + me) * sb;
    float* dst = reinterpret_cast<float*>(slot + sizeof(MailHeader));
    if (have)
      for (int s = B.tid; s < ld; s += B.nthr) dst[s] = src[s];
    if (B.tid == 0) {
      MailHeader* h = reinterpret_cast<MailHeader*>(slot);
      h->score = lscore; h->gidx = gidx; h->norm = nrm;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (B.tid < W) {
    MailHeader* h = reinterpret_cast<MailHeader*>(st->mail_peer[B.tid] + (int64_t)(par * W + me) * sb);
    st_release_sys_u64(&h->seq, seq);
    // wait for peer B.tid's slot in MY mailbox (bounded: a dead peer must not hang the GPU)
    const MailHeader* mine = reinterpret_cast<const MailHeader*>(st->mail_local + (int64_t)(par * W + B.tid) * sb);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(&mine->seq) != seq) {
      if (globaltimer_ns() - t0 > 10000000000ull) { st->comm_error = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (*reinterpret_cast<volatile int32_t*>(&st->comm_error)) {
    if (B.tid == 0) st->halted = 1;
    __syncthreads();
    return;
  }
  int win = -1; double best = -INFINITY; int64_t bidx = -1;
  for (int p = 0; p < W; ++p) {
    const MailHeader* h = reinterpret_cast<const MailHeader*>(st->mail_local + (int64_t)(par * W + p) * sb);
    const double sc = __ldcg(&h->score);
    const int64_t gi = __ldcg(reinterpret_cast<const long long*>(&h->gidx));
    if (gi < 0) continue;
    if (win < 0 || sc > best || (sc == best && gi < bidx)) { win = p; best = sc; bidx = gi; }
  }
  if (win < 0) {                             // no rank has a comparable row (non-finite matrix entries)
    if (B.tid == 0) { st->comm_error = 2; st->halted = 1; st->seq = seq; }
    __syncthreads();
    return;
  }
  const unsigned char* wslot = st->mail_local + (int64_t)(par * W + win) * sb;
  const float* wsrc = reinterpret_cast<const float*>(wslot + sizeof(MailHeader));
  for (int s = B.tid; s < ld; s += B.nthr) st->wrow[s] = __ldcg(wsrc + s);
  *f = bidx;
  *norm = __ldcg(&reinterpret_cast<const MailHeader*>(wslot)->norm);
  *row = st->wrow;
  if (B.tid == 0) st->seq = seq;
  __syncthreads();
}

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
