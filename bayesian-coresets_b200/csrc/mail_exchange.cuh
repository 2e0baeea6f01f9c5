// Block-wide NVLink peer-memory candidate exchange of the N-sharded solvers (used by the step kernels and by the control
// CTA of the persistent OrthoPursuit kernel).
#pragma once
#include <cuda_runtime.h>
#include "step_logic.h"
#include "sync_ptx.cuh"

namespace bcg {

// Every rank posts (float64 score, global index, norm, unit row) of its best local row into slot
// [parity][rank] of EVERY rank's mailbox with plain peer stores over NVLink, publishes it with a
// system-scope release of the sequence number, then waits (acquire) for the `world` slots of its
// own mailbox and picks the global winner: max score, ties -> lowest global index.  One fused
// all-gather per greedy iteration, no host involvement, no NCCL launch on the critical path.
// Slots are double-buffered on the parity of the sequence number: a rank can only be one
// exchange ahead of its slowest peer, so a slot is never overwritten while it is being read.
__device__ inline void mail_exchange(const Blk& B, SolverState* st, uint32_t lrow, double lscore, int64_t* f,
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

}  // namespace bcg
