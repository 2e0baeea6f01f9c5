// The whole greedy build(itrs) loop of GIGA / Frank-Wolfe as ONE persistent cooperative kernel.
//
// Why: at N = 1e6, S = 256 the HBM roofline of one iteration is ~157 us; two kernel launches per
// iteration plus a block-wide step kernel whose ~20 dependent global round trips cost ~30 us
// throw 15 % away.  Here nothing is launched inside the loop:
//   * one CTA per SM; warps 0..WPB-1 of every CTA are the streaming scan engines of
//     scan_kernel.cuh (private TMA ring per warp), and their TMA pipeline runs ACROSS iterations
//     -- the first tiles of iteration t+1 are already landing in shared memory while iteration t
//     is being resolved, because the matrix does not change, only the direction does;
//   * the extra warp of CTA 0 is the control warp: it keeps A w (float64) in registers for the
//     whole build, waits for the grid to arrive, reduces the 2 candidates per CTA, fetches the
//     winning row, does the geodesic / line-search reweight with warp shuffles only (no block
//     barrier), publishes the next direction with a release store, and does the O(K)
//     bookkeeping (weights rescale, active-set append, event log) while the grid is already
//     scanning again;
//   * grid-wide ordering is two monotonic counters in global memory (arrive / go).
//   * for 128 < S <= 512 the scan warps stream a float16 COPY of the rows (half the bytes) and bound every row's float32
//     score from it; only ring stages whose bound reaches the near-tie window of the maximum are re-scanned from the
//     float32 rows, so the selection is bit-identical to the float32 stream (scan_cta_body<..., CH16 > 0>,
//     filter_bounds.h: the error bound and its proof obligation).
// Semantics are those of step_logic.h (snnls.py:41-78, giga.py:20-64, frankwolfe.py:15-40);
// A w is updated incrementally (A w' = alpha A w + (w_f' - alpha w_f) a_f) and re-summed exactly
// from the active set every kRefreshEvery iterations, off the critical path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bcg_state.h"
#include "scan_core.cuh"
#include "step_logic.h"
#include "sync_ptx.cuh"

namespace bcg {

constexpr int kRefreshEvery = 16;

// ---------------------------------------------------------------------------------------------
// control warp helpers (lane l owns columns l, l+32, ... of every S-vector)
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void wsum(double (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
}

// float64 score of one unit row against the current direction(s) held in shared memory (sdir: 2 x S)
template <int J>
__device__ __forceinline__ double row_score_regs(const double (&x)[J], const double* sdir, int S, bool giga, int lane) {
  double v[2] = {0., 0.};
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) {
      v[0] += x[j] * sdir[s];
      if (giga) v[1] += x[j] * sdir[S + s];
    }
  }
  wsum<2>(v);
  return giga ? giga_score64(v[0], v[1]) : v[0];
}

template <int J>
__device__ __forceinline__ void load_row(const float* row, int S, int lane, double (&x)[J]) {
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    x[j] = (s < S) ? (double)__ldcg(row + s) : 0.;
  }
}

// warp-level arg-best over the per-CTA candidates, skipping rows listed in `skip`
__device__ __forceinline__ void warp_pick(const ScanCand* c, int n, const uint32_t* skip, int nskip, int lane,
                                          float* s_out, uint32_t* r_out) {
  float bs = -INFINITY;
  uint32_t br = kNoRow;
  for (int i = lane; i < n; i += 32) {
    const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(c + i));
    const float sc = __uint_as_float((unsigned int)(raw & 0xffffffffull));
    const uint32_t rw = (uint32_t)(raw >> 32);
    bool skipped = false;
    for (int q = 0; q < nskip; ++q) skipped |= (skip[q] == rw);
    if (!skipped && cand_better(sc, rw, bs, br)) { bs = sc; br = rw; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float s2 = __shfl_xor_sync(0xffffffffu, bs, off);
    const uint32_t r2 = __shfl_xor_sync(0xffffffffu, br, off);
    if (cand_better(s2, r2, bs, br)) { bs = s2; br = r2; }
  }
  *s_out = bs;
  *r_out = br;
}

// N-sharding exchange, warp-level (protocol of mail_exchange in step_kernels.cuh).  Everything the
// exchange needs is hoisted into MailCtx / shared memory once per build call; the candidate row is
// already in registers.  Critical path: payload stores to all peers -> release of W flags -> acquire of
// the W flags in my mailbox -> W headers read in parallel (one lane each) -> winner's row.
struct MailCtx {
  int W, me, ld;
  int64_t sb;                 // slot bytes
  unsigned char* local;       // my mailbox
  unsigned char** peers;      // shared-memory array of the W mapped mailboxes
  unsigned long long seq;     // running sequence number (kept in registers, stored back at the end)
};

template <int J>
__device__ __forceinline__ bool mail_exchange_warp(MailCtx& mc, int lane, int S, const double (&xrow)[J], bool have,
                                                   double lscore, int64_t gidx, double nrm, int64_t* f, double* norm,
                                                   const float** row) {
  const int W = mc.W, me = mc.me;
  const unsigned long long seq = ++mc.seq;
  const int par = (int)(seq & 1ull);
  const int64_t my_off = (int64_t)(par * W + me) * mc.sb;
  for (int p = 0; p < W; ++p) {
    unsigned char* slot = mc.peers[p] + my_off;
    float* dst = reinterpret_cast<float*>(slot + sizeof(MailHeader));
    if (have) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int s = lane + 32 * j;
        if (s < S) dst[s] = (float)xrow[j];
      }
    }
    if (lane == 0) {
      MailHeader* h = reinterpret_cast<MailHeader*>(slot);
      h->score = lscore; h->gidx = have ? gidx : -1; h->norm = nrm;
    }
  }
  // payload stores of all lanes -> __syncwarp -> system-scope release by the W flag-writing lanes (the
  // release is cumulative over the warp's earlier stores)
  __syncwarp();
  if (lane < W) st_release_sys_u64(&reinterpret_cast<MailHeader*>(mc.peers[lane] + my_off)->seq, seq);
  __syncwarp();
  // warp-uniform wait: lane p watches peer p's slot in MY mailbox; bounded (a dead peer must not hang the GPU)
  const MailHeader* mine =
      reinterpret_cast<const MailHeader*>(mc.local + (int64_t)(par * W + (lane < W ? lane : 0)) * mc.sb);
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; ++spins) {
    const bool done = ld_acquire_sys_u64(&mine->seq) == seq;
    if (__all_sync(0xffffffffu, done)) break;
    if ((spins & 1023u) == 1023u) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (__any_sync(0xffffffffu, now - t0 > 10000000000ull)) return false;
    }
  }
  // lane p reads header p; warp arg-max: max score, ties -> lowest global index
  double sc = -INFINITY; long long gi = -1; double nm = 0.;
  if (lane < W) {
    sc = __ldcg(&mine->score);
    gi = __ldcg(reinterpret_cast<const long long*>(&mine->gidx));
    nm = __ldcg(&mine->norm);
  }
  int win = (gi >= 0) ? lane : -1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double s2 = __shfl_xor_sync(0xffffffffu, sc, off);
    const long long g2 = __shfl_xor_sync(0xffffffffu, gi, off);
    const double n2 = __shfl_xor_sync(0xffffffffu, nm, off);
    const int w2 = __shfl_xor_sync(0xffffffffu, win, off);
    const bool take = (w2 >= 0) && (win < 0 || s2 > sc || (s2 == sc && g2 < gi));
    if (take) { sc = s2; gi = g2; nm = n2; win = w2; }
  }
  if (win < 0) return false;
  *f = gi;
  *norm = nm;
  *row = reinterpret_cast<const float*>(mc.local + (int64_t)(par * W + win) * mc.sb + sizeof(MailHeader));
  return true;
}

// direction(s) for the next scan from the iterate held in registers; returns false when GIGA's
// cdirnrm < TOL (giga.py:28-29).  n2 = sum xw^2, bx = sum bn*xw, e2 = sum (xw-b)^2 precomputed.
template <int J>
__device__ __forceinline__ bool publish_direction(float* dir32, double* dir64, int S, int ld, bool giga, double tol,
                                                  const double (&xw)[J], const double* sb, const double* sbn, int lane,
                                                  double n2, double bx, double e2, double* cdirnrm_out) {
  if (giga) {
    double nw = sqrt(n2);
    if (nw == 0.) nw = 1.;
    const double inw = 1. / nw;
    const double bxw = bx * inw;
    double c2[1] = {0.};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < S) { const double cd = sbn[s] - bxw * (xw[j] * inw); c2[0] += cd * cd; }
    }
    wsum<1>(c2);
    const double cn = sqrt(c2[0]);
    *cdirnrm_out = cn;
    if (cn < tol) return false;
    const double icn = 1. / cn;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < ld) {
        const double x = (s < S) ? xw[j] * inw : 0.;
        const double c = (s < S) ? (sbn[s] - bxw * x) * icn : 0.;
        dir32[s] = (float)c;
        dir32[ld + s] = (float)x;
        if (s < S) { dir64[s] = c; dir64[S + s] = x; }
      }
    }
  } else {
    const double rn = sqrt(e2);
    const double inv = rn > 0. ? 1. / rn : 1.;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < ld) {
        const double r = (s < S) ? (sb[s] - xw[j]) * inv : 0.;
        dir32[s] = (float)r;
        if (s < S) dir64[s] = r;
      }
    }
  }
  return true;
}

// one pass over the per-CTA candidates: best and runner-up (score descending, row ascending)
__device__ __forceinline__ void warp_top2(const ScanCand* c, int n, int lane, float* s1o, uint32_t* r1o, float* s2o,
                                          uint32_t* r2o) {
  float s1 = -INFINITY, s2 = -INFINITY;
  uint32_t r1 = kNoRow, r2 = kNoRow;
  constexpr int kPerLane = 12;                       // 2 * 148 CTAs / 32 lanes, rounded up; looped beyond that
  for (int base = 0; base < n; base += 32 * kPerLane) {
    unsigned long long raw[kPerLane];
#pragma unroll
    for (int q = 0; q < kPerLane; ++q) {             // all loads in flight together: one L2 round trip
      const int i = base + lane + 32 * q;
      raw[q] = (i < n) ? __ldcg(reinterpret_cast<const unsigned long long*>(c + i)) : 0xffffffff00000000ull;
    }
#pragma unroll
    for (int q = 0; q < kPerLane; ++q) {
      const float sc = __uint_as_float((unsigned int)(raw[q] & 0xffffffffull));
      const uint32_t rw = (uint32_t)(raw[q] >> 32);
      if (cand_better(sc, rw, s1, r1)) { s2 = s1; r2 = r1; s1 = sc; r1 = rw; }
      else if (cand_better(sc, rw, s2, r2)) { s2 = sc; r2 = rw; }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float a1 = __shfl_xor_sync(0xffffffffu, s1, off);
    const uint32_t b1 = __shfl_xor_sync(0xffffffffu, r1, off);
    const float a2 = __shfl_xor_sync(0xffffffffu, s2, off);
    const uint32_t b2 = __shfl_xor_sync(0xffffffffu, r2, off);
    // merge two (best, runner-up) pairs; the same row never appears in both (rows are unique)
    if (cand_better(a1, b1, s1, r1)) {
      if (cand_better(s1, r1, a2, b2)) { s2 = s1; r2 = r1; } else { s2 = a2; r2 = b2; }
      s1 = a1; r1 = b1;
    } else if (cand_better(a1, b1, s2, r2)) {
      s2 = a1; r2 = b1;
    }
  }
  *s1o = s1; *r1o = r1; *s2o = s2; *r2o = r2;
}

template <int J>
__device__ void control_loop(const LoopArgs& a, double* sb, double* sbn, double* sdir, unsigned char** speers) {
  SolverState* st = a.st;
  LoopCtl* ctl = a.ctl;
  const int lane = threadIdx.x & 31;
  const int S = st->S, ld = st->ld;
  const unsigned int G = gridDim.x;
  const bool giga = st->alg == BCG_ALG_GIGA;
  const int world = st->world;
  const int64_t row_offset = st->row_offset;
  const double bnorm = st->bnorm, nsum = st->nsum, tol = st->tol;
  float* dir32 = st->dir32;
  MailCtx mc;
  mc.W = world; mc.me = st->rank; mc.ld = ld; mc.sb = st->mail_slot_bytes; mc.local = st->mail_local;
  mc.peers = speers; mc.seq = st->seq;
  if (lane < world && world > 1) speers[lane] = st->mail_peer[lane];
  const float* An = st->An;
  const double* norms = st->norms;
  int64_t* act_idx = st->act_idx;
  double* act_w = st->act_w;
  double* act_norm = st->act_norm;
  float* act_rows = st->act_rows;
  bcg_iter_event* events = st->events;

  double xw[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    xw[j] = 0.;
    if (s < S) { xw[j] = st->xw[s]; sb[s] = st->b[s]; sbn[s] = st->bn[s]; }
  }
  __syncwarp();

  int nact = st->nact;
  int retried = a.cont ? st->retried : 0;   // snnls.py:40: local to the build() call (kept across an exact-selection stop)
  int halted = 0;
  int stop_exact = 0;
  int n_events = st->n_events;
  const bool check_mono = st->check_monotone != 0;
  const bool force_exact = st->force_exact != 0;
  double err = st->err;
  double sel_aux = 0.;
  int npos = 0;
  for (int k = lane; k < nact; k += 32) npos += (act_w[k] > 0.) ? 1 : 0;
  npos = __reduce_add_sync(0xffffffffu, npos);

  auto iterate_sums = [&](double& n2, double& bx, double& e2) {
    double v[3] = {0., 0., 0.};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int s = lane + 32 * j;
      if (s < S) { v[0] += xw[j] * xw[j]; v[1] += sbn[s] * xw[j]; const double r = xw[j] - sb[s]; v[2] += r * r; }
    }
    wsum<3>(v);
    n2 = v[0]; bx = v[1]; e2 = v[2];
  };
  auto push = [&](int code, int64_t f, double e, double a0, double a1) {
    if (lane == 0) {
      bcg_iter_event* ev = events + n_events;
      ev->code = code; ev->nact = nact; ev->f = f; ev->error = e; ev->aux0 = a0; ev->aux1 = a1;
    }
    __syncwarp();
    n_events += 1;
  };
  // direction stores of all lanes -> __syncwarp -> release by lane 0 (cumulative over the warp's stores)
  auto publish = [&](unsigned int it_done, bool stop) {
    __syncwarp();
    if (lane == 0) {
      if (stop) *reinterpret_cast<volatile unsigned int*>(&ctl->stop) = 1u;
      st_release_gpu_u32(&ctl->go, it_done);
    }
    __syncwarp();
  };

  double n2, bx, e2;
  iterate_sums(n2, bx, e2);
  bool sel_ok = publish_direction<J>(dir32, sdir, S, ld, giga, tol, xw, sb, sbn, lane, n2, bx, e2, &sel_aux);
  publish(1u, false);

  int it = 0;
  for (; it < a.itrs; ++it) {
    spin_until_ge_gpu_u32(&ctl->arrive, G * (unsigned int)(it + 1), 20);
    if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 0] = globaltimer_ns();
    __syncwarp();

    bool failed = false;
    int fcode = 0; int64_t f = -1; double fa0 = 0., fa1 = 0.;
    double alpha = 0., beta = 0., nf_stored = 0., lscore = 0.;
    const float* frow = nullptr;
    double xf[J];
#pragma unroll
    for (int j = 0; j < J; ++j) xf[j] = 0.;
    const bool nonempty = npos > 0;

    if (!sel_ok) {
      // giga.py:28-29: the selection itself failed (the grid scanned a stale direction; ignored)
      failed = true; fcode = BCG_IT_FAIL_CDIR; fa0 = sel_aux;
    } else {
      // ---- winner: best fp32 candidate; near ties re-scored in float64 --------------------------
      const int nc = 2 * (int)G;
      float top = -INFINITY, second = -INFINITY; uint32_t lrow = kNoRow, row2 = kNoRow;
      bool ambiguous = false;
      const bool pre = a.use_pre && it == 0;
      if (pre) {
        // this launch continues a build() call that stopped on an ambiguous candidate set: the local winner of its
        // first iteration is the result of the exact float64 pass (exact_scan_kernel), one candidate per CTA
        double ks = -INFINITY; long long kr = -1;
        for (int i = lane; i < st->n_exact_cands; i += 32) {
          const ExactCand x = st->exact_cands[i];
          if (x.row >= 0 && (kr < 0 || x.score > ks || (x.score == ks && x.row < kr))) { ks = x.score; kr = x.row; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double s2 = __shfl_xor_sync(0xffffffffu, ks, off);
          const long long r2 = __shfl_xor_sync(0xffffffffu, kr, off);
          if (r2 >= 0 && (kr < 0 || s2 > ks || (s2 == ks && r2 < kr))) { ks = s2; kr = r2; }
        }
        lrow = kr < 0 ? kNoRow : (uint32_t)kr;
        lscore = ks;
        if (lane == 0) { st->need_exact = 0; st->n_exact += 1; }
        __syncwarp();
      } else {
        float lostmax = -INFINITY;
        for (int i = lane; i < (int)G; i += 32) lostmax = fmaxf(lostmax, __ldcg(a.cta_lost + i));
        warp_top2(a.cta_cands, nc, lane, &top, &lrow, &second, &row2);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lostmax = fmaxf(lostmax, __shfl_xor_sync(0xffffffffu, lostmax, off));
        lscore = (double)top;
        // exactness of the candidate set (SolverState::cand_lost): an unpublished score inside the near-tie window
        ambiguous = (lrow != kNoRow) && (lostmax >= top - (2e-5f + 1e-5f * fabsf(top)) || force_exact);
      }
      if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 4] = globaltimer_ns();
      __syncwarp();
      const float thr = top - (2e-5f + 1e-5f * fabsf(top));
      const bool near_tie = !pre && !ambiguous && (lrow != kNoRow) && (row2 != kNoRow) && (second >= thr);
      if (near_tie) {
        // rare path: gather up to kRescoreMax candidates within the threshold, re-score in float64
        uint32_t chosen[kRescoreMax];
        int nch = 0;
        chosen[nch++] = lrow;
        while (nch < kRescoreMax) {
          float s2; uint32_t r2;
          warp_pick(a.cta_cands, nc, chosen, nch, lane, &s2, &r2);
          if (r2 == kNoRow || !(s2 >= thr)) break;
          chosen[nch++] = r2;
        }
        if (nch == kRescoreMax) {               // more published candidates inside the window than are re-scored?
          float s2; uint32_t r2;
          warp_pick(a.cta_cands, nc, chosen, nch, lane, &s2, &r2);
          ambiguous = (r2 != kNoRow) && (s2 >= thr);
        }
        uint32_t brow = kNoRow; double best = -INFINITY;
        for (int r = 0; r < nch; ++r) {
          double xr[J];
          load_row<J>(An + (size_t)chosen[r] * ld, S, lane, xr);
          const double sc = row_score_regs<J>(xr, sdir, S, giga, lane);
          if (brow == kNoRow || sc > best || (sc == best && chosen[r] < brow)) { best = sc; brow = chosen[r]; }
        }
        lrow = brow; lscore = best;
      }
      if (ambiguous) {
        // stop BEFORE this iteration: the host runs the exact float64 pass and relaunches (rare; bcg_solver_build)
        if (lane == 0) st->need_exact = 1;
        __syncwarp();
        stop_exact = 1;
        break;
      }
      const bool have = lrow != kNoRow;
      if (world == 1 && !have) {                 // no comparable score at all (non-finite matrix)
        if (lane == 0) { st->comm_error = 2; }
        __syncwarp();
        halted = 1;
        break;
      }
      if (have) {
        nf_stored = norms[lrow];
        load_row<J>(An + (size_t)lrow * ld, S, lane, xf);        // unit row; scaled by the norm below
        f = row_offset + (int64_t)lrow;
        frow = An + (size_t)lrow * ld;
      }
      if (world > 1) {
        // cross-rank comparison needs the float64 score of the local best
        if (have) lscore = row_score_regs<J>(xf, sdir, S, giga, lane);
        if (!mail_exchange_warp<J>(mc, lane, S, xf, have, lscore, f, nf_stored, &f, &nf_stored, &frow)) {
          if (lane == 0) { st->comm_error = 1; }
          __syncwarp();
          halted = 1;
          break;
        }
        load_row<J>(frow, S, lane, xf);                          // the winner's unit row from my mailbox
      }
#pragma unroll
      for (int j = 0; j < J; ++j) xf[j] *= nf_stored;
      if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 5] = globaltimer_ns() + (unsigned long long)(xf[0] == 12345.678);
      __syncwarp();

      // ---- line search ---------------------------------------------------------------------------
      if (giga) {
        double v[3] = {0., 0., 0.};            // xf.xf, bn.xf, xw.xf   (xw.xw = n2, bn.xw = bx known)
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int s = lane + 32 * j;
          if (s < S) { v[0] += xf[j] * xf[j]; v[1] += sbn[s] * xf[j]; v[2] += xw[j] * xf[j]; }
        }
        wsum<3>(v);
        double nw = sqrt(n2);
        if (nw == 0.) nw = 1.;
        const double nf = sqrt(v[0]);
        const double bxf = v[1] / nf, bxw = bx / nw, xwxf = v[2] / (nw * nf);
        const double gA = bxf - bxw * xwxf;
        const double gB = bxw - bxf * xwxf;
        if (gA <= 0. || gB < 0.) {
          failed = true; fcode = BCG_IT_FAIL_GEODESIC; fa0 = gA; fa1 = gB;
        } else {
          const double ca = gB / (gA + gB) / nw;
          const double cb = gA / (gA + gB) / nf;
          double u[2] = {0., 0.};
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int s = lane + 32 * j;
            if (s < S) { const double x = ca * xw[j] + cb * xf[j]; u[0] += x * x; u[1] += x * sbn[s]; }
          }
          wsum<2>(u);
          const double nx = sqrt(u[0]);
          const double scale = bnorm / nx * (u[1] / nx);
          alpha = ca * scale;
          beta = cb * scale;
        }
      } else {
        if (!nonempty) {
          alpha = 0.;
          beta = nsum / nf_stored;
        } else {
          const double r = nsum / nf_stored;
          double v[2] = {0., 0.};
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int s = lane + 32 * j;
            if (s < S) { const double t = r * xf[j] - xw[j]; v[0] += t * (sb[s] - xw[j]); v[1] += t * t; }
          }
          wsum<2>(v);
          if (v[0] < 0. || v[1] == 0. || v[0] > v[1]) {
            failed = true; fcode = BCG_IT_FAIL_GAMMA; fa0 = v[0]; fa1 = v[1];
          } else {
            alpha = 1. - v[0] / v[1];
            beta = r * v[0] / v[1];
          }
        }
      }
    }

    // ---- apply: w <- alpha w ; w[f] <- max(0, w[f] + beta) ; A w incrementally ---------------------
    // A w' = alpha A w + delta a_f with delta = w_f' - alpha w_f.  When alpha, beta >= 0 the clamp
    // cannot engage and delta = beta, so the slot search stays off the critical path.
    if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 6] = globaltimer_ns() + (unsigned long long)(alpha == 12345.678);
    __syncwarp();
    int slot = -2;                               // -2: not searched yet
    double wf_new = 0., delta = beta;
    double n2n = 0., bxn = 0., e2n = 0.;
    auto find_slot = [&]() {
      int sl = 0x7fffffff;
      for (int k = lane; k < nact; k += 32)
        if (act_idx[k] == f && k < sl) sl = k;
      sl = __reduce_min_sync(0xffffffffu, sl);
      slot = (sl == 0x7fffffff) ? -1 : sl;
      const double wf_old = slot >= 0 ? act_w[slot] : 0.;
      wf_new = fmax(0., alpha * wf_old + beta);
      return wf_new - alpha * wf_old;
    };
    if (!failed) {
      if (!(alpha >= 0. && beta >= 0.)) delta = find_slot();
      double v[3] = {0., 0., 0.};
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int s = lane + 32 * j;
        if (s < S) {
          const double x = alpha * xw[j] + delta * xf[j];
          v[0] += x * x; v[1] += sbn[s] * x;
          const double r = x - sb[s]; v[2] += r * r;
        }
      }
      wsum<3>(v);
      n2n = v[0]; bxn = v[1]; e2n = v[2];
      const double err_new = sqrt(e2n);
      if (check_mono && nonempty && err_new > err) {         // snnls.py:56-61
        failed = true; fcode = BCG_IT_FAIL_MONOTONE; fa0 = err_new; fa1 = err;
      }
    }

    if (failed) {
      // snnls.py:63-72.  The state is unchanged, so the retry fails identically: when another
      // iteration is available it is consumed here (second consecutive failure -> latch).
      push(fcode, f, err, fa0, fa1);
      if (retried || it + 1 < a.itrs) {
        if (!retried) { push(fcode, f, err, fa0, fa1); }
        halted = 1;
      } else {
        retried = 1;
      }
      if (halted) break;
      continue;                                 // it was the last iteration of this call
    }

    // commit the iterate, publish the next direction, THEN do the O(K) bookkeeping
#pragma unroll
    for (int j = 0; j < J; ++j) xw[j] = alpha * xw[j] + delta * xf[j];
    err = sqrt(e2n);
    n2 = n2n; bx = bxn; e2 = e2n;
    if (check_mono && nonempty) retried = 0;    // snnls.py:62 (inside the monotone branch)
    const bool last = (it + 1 == a.itrs);
    if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 7] = globaltimer_ns() + (unsigned long long)(err == 12345.678);
    __syncwarp();
    if (!last) {
      sel_ok = publish_direction<J>(dir32, sdir, S, ld, giga, tol, xw, sb, sbn, lane, n2, bx, e2, &sel_aux);
      publish((unsigned int)(it + 2), false);
    }
    if (a.trace && lane == 0) a.trace[(size_t)it * 8 + 1] = globaltimer_ns();
    __syncwarp();

    // ---- bookkeeping (overlaps the next scan) --------------------------------------------------
    if (slot == -2) (void)find_slot();
    int np = 0;
    for (int k = lane; k < nact; k += 32) {
      double wk = alpha * act_w[k];
      if (k == slot) wk = wf_new;
      act_w[k] = wk;
      np += (wk > 0.) ? 1 : 0;
    }
    if (slot < 0) {
      slot = nact;
      for (int s = lane; s < ld; s += 32) act_rows[(size_t)slot * ld + s] = __ldcg(frow + s);
      if (lane == 0) { act_idx[slot] = f; act_norm[slot] = nf_stored; act_w[slot] = wf_new; }
      __syncwarp();
      np += (lane == 0 && wf_new > 0.) ? 1 : 0;
      nact += 1;
    }
    npos = __reduce_add_sync(0xffffffffu, np);
    __syncwarp();
    push(BCG_IT_OK, f, err, lscore, 0.);

    if (((it + 1) % kRefreshEvery) == 0 && !last) {
      // exact re-summation of A w from the active set (bounds the drift of the incremental update)
      double xr[J];
#pragma unroll
      for (int j = 0; j < J; ++j) xr[j] = 0.;
      for (int k = 0; k < nact; ++k) {
        const double c = act_w[k] * act_norm[k];
        if (c == 0.) continue;
        const float* row = act_rows + (size_t)k * ld;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int s = lane + 32 * j;
          if (s < S) xr[j] += c * (double)row[s];
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) xw[j] = xr[j];
      iterate_sums(n2, bx, e2);
      err = sqrt(e2);
    }
  }

  // ---- end of the build call: stop the grid, write the state back -------------------------------
  publish((unsigned int)(a.itrs + 2), true);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int s = lane + 32 * j;
    if (s < S) st->xw[s] = xw[j];
  }
  if (lane == 0) {
    st->nact = nact;
    st->err = err;
    st->retried = retried;
    st->halted = halted ? 1 : st->halted;
    st->n_events = n_events;
    st->select_failed = sel_ok ? 0 : 1;
    st->sel_aux = sel_aux;
    st->seq = mc.seq;
    st->iters_done = stop_exact ? it : a.itrs;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// the scan warps of one CTA: stream this CTA's share of the matrix once per iteration, publish the CTA's two best
// candidates + its lost score, arrive.  cta_rank / n_ctas: position among the SCANNING CTAs (the OrthoPursuit kernel
// reserves CTA 0 for the control path).  Shared memory: rings | barriers | 64 candidate slots | ... | slot records.
// ---------------------------------------------------------------------------------------------
constexpr int kFiltCap = 32;       // re-scan slots per scan warp and iteration (one per lane)

// CH16 > 0: float16 pre-filter (filter_bounds.h).  The ring streams the float16 copy (half the bytes); every ring stage
// yields an upper bound of the float32 scores of its rows and raises the warp's lower bound L of the float32 maximum;
// stages whose upper bound reaches filter_threshold(L) are remembered (kFiltCap slots, compacted against the rising L)
// and re-scanned afterwards from the float32 rows in global memory with the unchanged Core::batch.  Rows that are not
// re-scanned provably score below (float32 maximum - near-tie window), so best / runner-up / lost -- and with them the
// selection, the near-tie re-scoring and the ambiguity test -- are exactly those of the plain float32 scan.  A warp
// that runs out of slots reports lost = +inf, which sends the iteration to the exact float64 pass.
template <int CH, int NDIR, int LPR, int R, int CH16>
__device__ __forceinline__ void scan_cta_body(const LoopArgs& a, unsigned char* smem_raw, int cta_rank, int n_ctas) {
  using Core = ScanCore<CH, NDIR, LPR, R>;
  constexpr bool F16 = CH16 > 0;
  using FCore = Filter16Core<(F16 ? CH16 : 1), NDIR, (CH16 == 1 ? 8 : 4)>;   // short rows: 8 per batch amortise the reduce + bounds
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = a.wpb;
  const ScanGeom& q = a.g;
  const uint32_t stage_floats = (uint32_t)q.rps * (uint32_t)q.ld;
  const size_t ring_bytes = (size_t)wpb * q.stages * stage_floats * sizeof(float);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes);
  ScanCand* cta_c = reinterpret_cast<ScanCand*>(bars_all + wpb * q.stages);   // 64 slots (see the end of the scan loop)
  unsigned char** speers = reinterpret_cast<unsigned char**>(reinterpret_cast<double*>(cta_c + 64) + 4 * a.st->S);
  LoopCtl* ctl = a.ctl;
  const int64_t gw = (int64_t)cta_rank * wpb + warp;
  const int64_t GW = (int64_t)n_ctas * wpb;
  float* wbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * q.stages * stage_floats;
  uint64_t* bars = bars_all + warp * q.stages;
  if (lane == 0) {
    for (int s = 0; s < q.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t policy = l2_policy(q.evict_first);

  const int g = lane & (LPR - 1);
  const int grp = lane / LPR;
  const int nchunk = q.ld >> 2;
  // what the ring streams: the float32 rows, or their float16 copy
  const int rps = F16 ? q.rps16 : q.rps;
  const uint32_t row_bytes = F16 ? (uint32_t)q.ld16 * 2u : (uint32_t)q.ld * 4u;
  const unsigned char* src = F16 ? reinterpret_cast<const unsigned char*>(q.An16) : reinterpret_cast<const unsigned char*>(q.An);
  const int64_t n_chunks = (q.n_rows + rps - 1) / rps;
  // Work split per iteration: the first n_s chunks of every warp are STATIC (chunk gw + k GW, interleaved so
  // that neighbouring warps stream neighbouring rows); the remaining chunks [n_static, n_chunks) are claimed
  // DYNAMICALLY with one atomic per chunk from a per-iteration counter, so SMs that the memory system serves
  // faster take more of the tail and the whole grid arrives together (measured: with a purely static split
  // the first CTA finished ~30 us before the last one at N = 1e6, S = 256).
  const int64_t n_s = (int64_t)((double)(n_chunks / GW) * a.static_frac);
  const int64_t n_static = n_s * GW;

  // Ring bookkeeping, all warp-uniform (every lane tracks the same counters; the leader lane issues copies
  // and claims through PTX predicates -- no divergent branch anywhere in the scan warps).  Chunks are
  // issued in consumption order, iteration after iteration, so the tile stream runs ahead ACROSS
  // iterations.  sm_row0 / sm_it describe what each ring slot holds.
  const bool leader = lane == 0;
  int64_t* sm_row0 = reinterpret_cast<int64_t*>(speers + kMaxWorld) + warp * q.stages;
  int* sm_it = reinterpret_cast<int*>(reinterpret_cast<int64_t*>(speers + kMaxWorld) + wpb * q.stages) + warp * q.stages;
  // re-scan slots of this warp (float16 pre-filter): first row and upper bound of a remembered stage
  uint32_t* fl_base = reinterpret_cast<uint32_t*>(sm_it - warp * q.stages + wpb * q.stages);
  uint32_t* fl_row = fl_base + warp * kFiltCap;
  float* fl_ub = reinterpret_cast<float*>(fl_base + wpb * kFiltCap) + warp * kFiltCap;
  int issue_slot = 0;
  int64_t issue_k = 0;         // static chunks of iteration issue_it issued so far
  int issue_it = 0;            // iteration the next issue belongs to
  int outstanding = 0;         // issued, not yet consumed
  unsigned int pending = 0;    // pre-claimed dynamic chunk (valid in lane 0 when have_pending)
  bool have_pending = false;
  auto claim = [&](int it_claim) {                  // asynchronous: the result is only read at the next issue
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p atom.global.add.u32 %0, [%1], 1;\n\t}"
        : "=r"(pending)
        : "l"(a.claims + it_claim), "r"((uint32_t)leader)
        : "memory");
    have_pending = true;
  };
  auto issue_next = [&]() {
    for (;;) {
      if (issue_it >= a.itrs || n_chunks == 0) return;
      int64_t c;
      if (issue_k < n_s) {
        c = gw + issue_k * GW;
        ++issue_k;
      } else {
        if (!have_pending) claim(issue_it);
        c = n_static + (int64_t)__shfl_sync(0xffffffffu, pending, 0);
        have_pending = false;
        if (c >= n_chunks) {                         // this iteration's chunks are all taken: next iteration
          ++issue_it;
          issue_k = 0;
          continue;
        }
      }
      const int64_t row0 = c * rps;
      const int64_t left = q.n_rows - row0;
      const uint32_t nr = (uint32_t)(left < rps ? left : rps);
      tma_load_rows(&bars[issue_slot], wbuf + (size_t)issue_slot * stage_floats,
                    reinterpret_cast<const float*>(src + (size_t)row0 * row_bytes), nr * row_bytes, policy, leader);
      if (leader) { sm_row0[issue_slot] = row0; sm_it[issue_slot] = issue_it; }
      if (++issue_slot == q.stages) issue_slot = 0;
      ++outstanding;
      if (issue_k >= n_s) claim(issue_it);           // pre-claim: its latency hides behind the next tile's math
      return;
    }
  };
  for (int s = 0; s < q.stages; ++s) issue_next();   // tiles do not depend on the direction
  __syncwarp();

  int slot = 0;
  uint32_t parity = 0;
  for (int it = 0; it < a.itrs; ++it) {
    spin_until_ge_gpu_u32(&ctl->go, (unsigned int)(it + 1), 32);
    if (__shfl_sync(0xffffffffu, *reinterpret_cast<volatile unsigned int*>(&ctl->stop), 0)) break;
    if (a.trace && gw == 0 && lane == 0) a.trace[(size_t)it * 8 + 2] = globaltimer_ns();
    __syncwarp();

    float best = -INFINITY, lost = -INFINITY;
    uint32_t brow = kNoRowU;
    if constexpr (!F16) {
      float4 d0[CH];
      float4 d1[CH];
      Core::load_dirs(a.st->dir32, q.ld, nchunk, g, d0, d1);
      while (outstanding > 0 && sm_it[slot] == it) {   // uniform: shared-memory broadcast reads
        mbar_wait(&bars[slot], parity);
        const int64_t row0 = sm_row0[slot];
        const int64_t left = q.n_rows - row0;
        const int nr = (int)(left < rps ? left : rps);
        const float* tile = wbuf + (size_t)slot * stage_floats;
        for (int b0 = 0; b0 < nr; b0 += Core::RB)
          Core::batch(tile + (size_t)b0 * q.ld, q.ld, nchunk, g, grp, nr - b0, (uint32_t)(row0 + b0), d0, d1, best, brow, lost);
        __syncwarp();   // every lane's shared-memory reads of this stage (and of its slot record) are complete
        --outstanding;
        issue_next();
        __syncwarp();   // the leader's slot record is visible to all lanes
        if (++slot == q.stages) { slot = 0; parity ^= 1u; }
      }
    } else {
      // ---- pass A: bounds from the float16 copy ----------------------------------------------------------------
      // lower bound L of the float32 maximum (warp-uniform): the rows this warp has seen, and -- through one word of global
      // memory per iteration, raised with atomicMax and polled every 8th stage one stage ahead of its use -- the rows every
      // other warp of the GPU has seen.  Any valid lower bound may be used at any time, so the exchange needs no ordering.
      float Lw = -INFINITY;
      int* Lslot = a.filt_L + it;
      float Lpub = -INFINITY;                          // what the global word is known to hold at least
      int Lpend = (int)0x80808080;
      int nstage = 0;
      int nlist = 0;
      bool overflow = false;
      {
        float dh0[F16 ? CH16 : 1][8];
        float dh1[F16 ? CH16 : 1][8];
        float dn0, dn1;
        FCore::load_dirs(a.st->dir32, q.ld, a.st->S, lane, dh0, dh1, &dn0, &dn1);
        const float e0 = q.eps16 * dn0, e1 = q.eps16 * dn1;
        const int ngroups = q.ld16 >> 3;
        while (outstanding > 0 && sm_it[slot] == it) {
          mbar_wait(&bars[slot], parity);
          const int64_t row0 = sm_row0[slot];
          const int64_t left = q.n_rows - row0;
          const int nr = (int)(left < rps ? left : rps);
          const __half* tile = reinterpret_cast<const __half*>(wbuf + (size_t)slot * stage_floats);
          float ubm = -INFINITY, lbm = -INFINITY;
          for (int b0 = 0; b0 < nr; b0 += FCore::RB)
            FCore::batch(tile + (size_t)b0 * q.ld16, q.ld16, ngroups, lane, nr - b0, dh0, dh1, e0, e1, ubm, lbm);
          ubm = warp_max_f(ubm);
          Lw = fmaxf(Lw, warp_max_f(lbm));
          if ((nstage & 7) == 1) { const float Lg = sortable_float(Lpend); Lw = fmaxf(Lw, Lg); Lpub = fmaxf(Lpub, Lg); }
          if ((nstage & 7) == 0) Lpend = __ldcg(Lslot);
          ++nstage;
          if (Lw > Lpub) {                             // this warp raises the GPU-wide bound
            if (lane == 0) atomicMax(Lslot, float_sortable(Lw));
            Lpub = Lw;
          }
          if (ubm >= filter_threshold(Lw)) {           // warp-uniform
            if (nlist == kFiltCap) {                   // compact the slots against the risen bound
              const uint32_t r_i = fl_row[lane];
              const float u_i = fl_ub[lane];
              const bool keep = u_i >= filter_threshold(Lw);
              const unsigned int m = __ballot_sync(0xffffffffu, keep);
              __syncwarp();
              if (keep) { const int pos = __popc(m & ((1u << lane) - 1u)); fl_row[pos] = r_i; fl_ub[pos] = u_i; }
              nlist = __popc(m);
              __syncwarp();
            }
            if (nlist < kFiltCap) {
              if (lane == 0) { fl_row[nlist] = (uint32_t)row0; fl_ub[nlist] = ubm; }
              ++nlist;
            } else {
              overflow = true;
            }
          }
          __syncwarp();
          --outstanding;
          issue_next();
          __syncwarp();
          if (++slot == q.stages) { slot = 0; parity ^= 1u; }
        }
      }
      // ---- pass B: the remembered stages again, float32 rows straight from global memory -------------------------------
      {
        float4 d0[CH];
        float4 d1[CH];
        Core::load_dirs(a.st->dir32, q.ld, nchunk, g, d0, d1);
        if (nlist > 0) Lw = fmaxf(Lw, sortable_float(__ldcg(Lslot)));
        const float thr = filter_threshold(Lw);
        unsigned int nres = 0;
        for (int i = 0; i < nlist; ++i) {
          if (!(fl_ub[i] >= thr)) continue;            // uniform (shared-memory broadcast)
          const int64_t row0 = (int64_t)fl_row[i];
          const int64_t left = q.n_rows - row0;
          const int nr = (int)(left < rps ? left : rps);
          for (int b0 = 0; b0 < nr; b0 += Core::RB)    // (the last batch may read up to RB - 1 rows of padding: kRowPad)
            Core::batch(q.An + (size_t)(row0 + b0) * q.ld, q.ld, nchunk, g, grp, nr - b0, (uint32_t)(row0 + b0), d0, d1, best,
                        brow, lost);
          nres += (unsigned int)nr;
        }
        if (overflow) lost = INFINITY;
        if (lane == 0) {
          if (nres) atomicAdd(&a.st->filt_rows, (unsigned long long)nres);
          if (overflow) a.st->filt_overflow = 1;
        }
        __syncwarp();
      }
    }

    if (a.trace && gw == 0 && lane == 0) a.trace[(size_t)it * 8 + 3] = globaltimer_ns();
    __syncwarp();
    // (Pulling the next chunks into L2 while the control warp resolves the iteration -- the rings are full and HBM idles for
    // ~5 us -- was measured slightly SLOWER, 155.5-157.7 vs 154.4 us per iteration at N = 1e6, S = 256, and was removed.)
    Core::warp_merge_lost(best, brow, lost);
    // per-warp results, double-buffered on the iteration parity: a warp that runs ahead into iteration it + 1 writes
    // the other half, and its write of iteration it + 2 is ordered behind the barrier of it + 1, which warp 0 only
    // joins after it has read iteration it.  Slots [par*16 + w]: candidate; [32 + par*16 + w].score: lost score.
    ScanCand* cc = cta_c + (it & 1) * 16;
    if (lane == 0) { cc[warp].score = best; cc[warp].row = brow; cc[32 + warp].score = lost; }
    named_bar_sync(1, wpb * 32);
    if (warp == 0) {
      // best and runner-up of the CTA's warps -> global, then arrive (release)
      float s1 = -INFINITY, l1 = -INFINITY; uint32_t r1 = kNoRowU;
      if (lane < wpb) { s1 = cc[lane].score; r1 = cc[lane].row; l1 = cc[32 + lane].score; }
      const float s_mine = s1; const uint32_t r_mine = r1;
      float bs = s1; uint32_t br = r1;
      Core::warp_merge(bs, br);
      if (r1 == br) { s1 = -INFINITY; r1 = kNoRowU; }     // exclude the winner, then second best
      float ss = s1; uint32_t sr = r1;
      Core::warp_merge(ss, sr);
      // what the CTA does not publish: the warps' lost scores and every warp best other than the two above
      if (r_mine != kNoRowU && r_mine != br && r_mine != sr) l1 = fmaxf(l1, s_mine);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) l1 = fmaxf(l1, __shfl_xor_sync(0xffffffffu, l1, off));
      if (lane == 0) {
        ScanCand c0; c0.score = bs; c0.row = br;
        ScanCand c1; c1.score = ss; c1.row = sr;
        a.cta_cands[2 * cta_rank] = c0;
        a.cta_cands[2 * cta_rank + 1] = c1;
        a.cta_lost[cta_rank] = l1;
        __threadfence();
        red_release_gpu_add(&ctl->arrive, 1u);
      }
    }
  }
  // drain tiles that were prefetched for an iteration that will not run (early stop)
  {
    for (int left = outstanding; left > 0; --left) {
      mbar_wait(&bars[slot], parity);
      if (++slot == q.stages) { slot = 0; parity ^= 1u; }
    }
  }
}

template <int CH, int NDIR, int LPR, int R, int J, int CH16>
__global__ void __launch_bounds__(384, 1) greedy_loop_kernel(const LoopArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = a.wpb;
  const ScanGeom& q = a.g;
  const uint32_t stage_floats = (uint32_t)q.rps * (uint32_t)q.ld;
  const size_t ring_bytes = (size_t)wpb * q.stages * stage_floats * sizeof(float);
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes);
  ScanCand* cta_c = reinterpret_cast<ScanCand*>(bars_all + wpb * q.stages);   // 64 slots (see the end of the scan loop)
  double* sb = reinterpret_cast<double*>(cta_c + 64);
  double* sbn = sb + a.st->S;

  // shared memory after the rings: barriers | per-warp candidates | b, bn (2 S) | float64 directions (2 S) |
  // peer mailbox pointers (8) | ring slot records
  double* sdir = sbn + a.st->S;
  unsigned char** speers = reinterpret_cast<unsigned char**>(sdir + 2 * a.st->S);
  if (warp == wpb) {                    // control warp (only CTA 0's does anything)
    if (blockIdx.x == 0) control_loop<J>(a, sb, sbn, sdir, speers);
    return;
  }

  scan_cta_body<CH, NDIR, LPR, R, CH16>(a, smem_raw, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace bcg
