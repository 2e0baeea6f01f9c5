// Device-resident solver state shared by the scan kernel (K1), the step kernel (K2) and the host
// API.  Plain C++ (no CUDA types) so that step_logic.h can also be compiled for the host by the
// logic check in tests/hostcheck (test-only; the product library never runs solver math on CPU).
#pragma once
#include <stdint.h>
#include "../../include/bcg.h"

namespace bcg {

// one candidate per scan warp: best fp32 score and the LOCAL row it belongs to
struct ScanCand {
  float score;
  uint32_t row;   // 0xffffffff = warp saw no rows
};

constexpr int kMaxWorld = 8;        // GPUs of one NVSwitch node
constexpr int kRescoreMax = 8;      // near-tie candidates re-scored in float64 per iteration

// result of one CTA of the exact float64 selection pass (exact_scan_kernel): best local row, lowest row on ties
struct ExactCand {
  double score;
  int64_t row;                      // -1: the CTA saw no comparable row
};

// mailbox slot written by one peer for one iteration parity (lives in IPC-shared device memory)
struct MailHeader {
  unsigned long long seq;           // iteration sequence number, written last (release)
  double score;                     // float64 re-scored best local score
  int64_t gidx;                     // global row index (-1: rank has no rows)
  double norm;                      // row norm
};

struct SolverState {
  // ---- problem -------------------------------------------------------------------------
  int32_t alg, S, ld;
  int32_t world, rank;
  int64_t n_local, row_offset, n_global;
  double tol, bnorm, nsum;
  const float* An;                  // n_local x ld unit rows
  const double* norms;              // n_local
  double* b;                        // S
  double* bn;                       // S   b / ||b||  (GIGA)
  // ---- iterate -------------------------------------------------------------------------
  double* xw;                       // S   A w
  double* xw_new;                   // S   candidate
  double* xf;                       // S   unnormalised selected row
  double* dir64;                    // 2 x S   float64 scan directions (re-scoring)
  float* dir32;                     // 2 x ld  float32 scan directions (K1 input)
  float* wrow;                      // ld      winner's unit row (copied out of the mailbox)
  double err;                       // ||A w - b||
  // ---- active set (replicated on every rank), selection order ----------------------------
  int32_t nact, cap;
  int64_t* act_idx;                 // cap       global indices
  double* act_w;                    // cap
  double* act_w_new;                // cap
  double* act_norm;                 // cap
  double* act_tmp;                  // cap       scratch (per-row values of the active set)
  float* act_rows;                  // cap x ld  copies of the unit rows
  // ---- loop control ----------------------------------------------------------------------
  int32_t retried, halted, select_failed, n_events;
  double sel_aux;
  bcg_iter_event* events;
  unsigned long long seq;           // iteration sequence number (mailbox protocol)
  // ---- scan output -------------------------------------------------------------------------
  ScanCand* cands;
  int32_t n_cands;
  int32_t comm_error;
  // ---- exact-selection fallback ----------------------------------------------------------------
  // The float32 scan publishes a bounded candidate set (one per warp / two per CTA).  Every scan warp also reports
  // the best float32 score it saw but did NOT publish (cand_lost).  When an unpublished score, or more than
  // kRescoreMax published ones, lie inside the near-tie window of the maximum, the candidate set may miss the
  // float64 arg-max: need_exact is raised and the selection is redone by exact_scan_kernel (float64 over all local
  // rows, lowest index on ties), whose per-CTA results land in exact_cands.
  float* cand_lost;                 // n_cands (scan_kernel) or null
  ExactCand* exact_cands;           // n_exact_cands entries
  int32_t n_exact_cands;
  int32_t need_exact;               // set by the scan's last CTA / the loop's control warp, cleared by the consumer
  int32_t n_exact;                  // selections resolved by the exact pass so far (diagnostics)
  int32_t iters_done;               // iterations consumed by the last persistent-kernel launch
  unsigned int scan_done;           // CTA completion counter of scan_kernel (last CTA resets it)
  // summary of the last scan by its last CTA: float32 maximum, the lowest row that attains it, published candidates inside
  // the near-tie window (0 = no summary).  With exactly one candidate in the window pick_local takes it without any reduction.
  float scan_top;
  uint32_t scan_top_row;
  int32_t scan_cnt;
  int32_t kkt_valid;                // OMP: the weights are the NNLS optimum of the active set (gradient ~ 0 on it)
  int32_t check_monotone;           // snnls.py:9 check_error_monotone
  int32_t force_exact;              // testing: treat every candidate set as ambiguous
  // ---- never-materialising select (lazy_select_kernel.cuh): An == null ---------------------------
  int32_t lazy;                     // 1: rows are recomputed from the raw data; the selection pass leaves its result below
  int64_t fused_row;                // local winner of the last selection pass (-1: none)
  double fused_score;               // its float64 score
  double wnorm;                     // its norm (its unit row is in wrow)
  // ---- N-sharding mailboxes ----------------------------------------------------------------
  unsigned char* mail_local;        // this rank's mailbox: [2][world] slots
  unsigned char* mail_peer[kMaxWorld];  // mapped mailboxes of all ranks (self included)
  int64_t mail_slot_bytes;
  // ---- float16 pre-filter of the persistent scan (filter_bounds.h) --------------------------------
  unsigned long long filt_rows;     // rows re-scanned in float32 so far (diagnostics; all other rows were excluded by bounds)
  int32_t filt_overflow;            // a scan warp ran out of re-scan slots (the iteration was resolved by the exact pass)
  // ---- diagnostics ---------------------------------------------------------------------------
  unsigned long long* omp_trace;    // null, or 16 device timestamps per OMP iteration (BCG_OMP_TRACE=1)
};

inline int64_t mail_slot_bytes_for(int ld) {
  int64_t raw = (int64_t)sizeof(MailHeader) + (int64_t)ld * 4;
  return (raw + 127) / 128 * 128;
}

}  // namespace bcg
