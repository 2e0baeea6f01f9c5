// K2 "step" logic: everything a greedy iteration does besides the N x S scan -- winner
// resolution with float64 re-scoring of near ties, the GIGA geodesic / Frank-Wolfe line-search
// reweight on the active set, the monotone-error check with the reference's retry / latch
// semantics, and the next scan direction.  Executed by ONE thread block; all branches are
// block-uniform.  Written against a tiny Blk abstraction so the same source also compiles for
// the host with nthr == 1 (tests/hostcheck: logic check against the oracle, test-only).
//
// Reference semantics followed (paths relative to the reference repository root):
//   snnls/snnls.py:41-78        iteration skeleton, monotone check, retry, numeric-limit latch
//   snnls/giga.py:20-38, 40-64  GIGA select direction / reweight
//   snnls/frankwolfe.py:15-40   Frank-Wolfe select direction / reweight
//   snnls/orthopursuit.py:17-38 OMP select (+ negative direction over the active set)
#pragma once
#include <math.h>
#include <stdint.h>
#include "bcg_state.h"

#if defined(__CUDACC__)
#define BCG_HD __host__ __device__ __forceinline__
#else
#define BCG_HD inline
#endif

namespace bcg {

constexpr uint32_t kNoRow = 0xffffffffu;
constexpr int kWideCols = 256;       // output columns per blk_combine round (8 per lane)

struct Blk {
  int tid, nthr;
  double* sred;      // shared scratch, >= 32 * 8 doubles
  double* wide;      // device only, optional: (nthr / 32) * kWideCols doubles of shared scratch for blk_combine;
                     // null selects the one-thread-per-output forms (host build, BCG_OMP_WIDE=0)
  BCG_HD void sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
};

// phase timestamp i of the current OMP iteration (diagnostics: BCG_OMP_TRACE=1); no-op on the host build
BCG_HD void omp_mark(const Blk& B, SolverState* st, int i) {
#ifdef __CUDA_ARCH__
  if (st->omp_trace && B.tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    st->omp_trace[(size_t)st->n_events * 16 + i] = t;
  }
#else
  (void)B; (void)st; (void)i;
#endif
}

// sum N per-thread values over the block; every thread receives the totals (fixed order)
template <int N>
BCG_HD void blk_sum(const Blk& B, double* v) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  const int warp = B.tid >> 5, lane = B.tid & 31, nw = B.nthr >> 5;
  if (lane == 0)
    for (int i = 0; i < N; ++i) B.sred[warp * N + i] = v[i];
  __syncthreads();
  for (int i = 0; i < N; ++i) {
    double t = 0.;
    for (int w = 0; w < nw; ++w) t += B.sred[w * N + i];
    v[i] = t;
  }
  __syncthreads();
#else
  (void)B; (void)v;
#endif
}

// block arg-best over (key descending, id ascending); payload carried along.  id < 0 = empty.
BCG_HD void blk_argbest(const Blk& B, double* key, int64_t* id, int* payload) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double k2 = __shfl_xor_sync(0xffffffffu, *key, off);
    int64_t i2 = __shfl_xor_sync(0xffffffffu, *id, off);
    int p2 = __shfl_xor_sync(0xffffffffu, *payload, off);
    bool take = (i2 >= 0) && ((*id < 0) || (k2 > *key) || (k2 == *key && i2 < *id));
    if (take) { *key = k2; *id = i2; *payload = p2; }
  }
  const int warp = B.tid >> 5, lane = B.tid & 31, nw = B.nthr >> 5;
  int64_t* sid = reinterpret_cast<int64_t*>(B.sred + 32);
  int* spl = reinterpret_cast<int*>(B.sred + 64);
  if (lane == 0) { B.sred[warp] = *key; sid[warp] = *id; spl[warp] = *payload; }
  __syncthreads();
  double bk = B.sred[0]; int64_t bi = sid[0]; int bp = spl[0];
  for (int w = 1; w < nw; ++w) {
    double k2 = B.sred[w]; int64_t i2 = sid[w]; int p2 = spl[w];
    bool take = (i2 >= 0) && ((bi < 0) || (k2 > bk) || (k2 == bk && i2 < bi));
    if (take) { bk = k2; bi = i2; bp = p2; }
  }
  *key = bk; *id = bi; *payload = bp;
  __syncthreads();
#else
  (void)B; (void)key; (void)id; (void)payload;
#endif
}

// ---------------------------------------------------------------------------------------------
// Two block-wide building blocks for the K x S algebra of the OMP / NNLS iteration.  With one thread per
// output and a K-step loop behind it every such product was a chain of ~K dependent-latency L2 round trips
// (ncu: the single-CTA kernel spent its time in long-scoreboard stalls); the wide forms split the K terms
// over the warps, keep 8-16 independent loads in flight per lane and reduce across warps through shared memory.
// All threads of the block must call them (they contain barriers); arguments are block-uniform.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct CombTerm { double coef; const T* row; int len; };

// fin(s, sum_{i < n_terms, s < len_i} coef_i * row_i[s]) for every s < n_out, called by exactly one thread per s
template <typename T, typename TermF, typename FinF>
BCG_HD void blk_combine(const Blk& B, int n_out, int n_terms, TermF term, FinF fin) {
#ifdef __CUDA_ARCH__
  if (B.wide) {
    const int lane = B.tid & 31, warp = B.tid >> 5, nw = B.nthr >> 5;
    for (int c0 = 0; c0 < n_out; c0 += kWideCols) {
      double acc[kWideCols / 32];
#pragma unroll
      for (int j = 0; j < kWideCols / 32; ++j) acc[j] = 0.;
      constexpr int G = 4;                                 // terms per warp pass: G * 8 independent loads per lane
      for (int i0 = warp * G; i0 < n_terms; i0 += nw * G) {
        CombTerm<T> t[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (i0 + g < n_terms) t[g] = term(i0 + g);
          else { t[g].coef = 0.; t[g].row = nullptr; t[g].len = 0; }
          if (t[g].coef == 0.) t[g].len = 0;               // warp-uniform
        }
        T v[G][kWideCols / 32];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < kWideCols / 32; ++j) {
            const int s = c0 + lane + 32 * j;
            v[g][j] = (s < t[g].len) ? t[g].row[s] : (T)0;
          }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int j = 0; j < kWideCols / 32; ++j) acc[j] += t[g].coef * (double)v[g][j];
      }
#pragma unroll
      for (int j = 0; j < kWideCols / 32; ++j) B.wide[warp * kWideCols + lane + 32 * j] = acc[j];
      __syncthreads();
      if (B.tid < kWideCols && c0 + B.tid < n_out) {
        double tot = 0.;
        for (int w = 0; w < nw; ++w) tot += B.wide[w * kWideCols + B.tid];
        fin(c0 + B.tid, tot);
      }
      __syncthreads();
    }
    return;
  }
#endif
  for (int s = B.tid; s < n_out; s += B.nthr) {
    double acc = 0.;
    for (int i = 0; i < n_terms; ++i) {
      const CombTerm<T> t = term(i);
      if (t.coef != 0. && s < t.len) acc += t.coef * (double)t.row[s];
    }
    fin(s, acc);
  }
  B.sync();
}

// emit(i, sum_{s < S} row_i[s] * x[s]) for every i < n whose rowf(i) is not null, called by exactly one thread per i
template <typename T, typename RowF, typename EmitF>
BCG_HD void blk_dots(const Blk& B, int n, int S, RowF rowf, const double* x, EmitF emit) {
#ifdef __CUDA_ARCH__
  {
    const int lane = B.tid & 31, warp = B.tid >> 5, nw = B.nthr >> 5;
    constexpr int G = 4;                                   // rows per warp pass: G * S/32 independent loads per lane
    for (int i0 = warp * G; i0 < n; i0 += nw * G) {
      const T* r[G];
      double p[G];
#pragma unroll
      for (int k = 0; k < G; ++k) { r[k] = (i0 + k < n) ? rowf(i0 + k) : nullptr; p[k] = 0.; }
      for (int s = lane; s < S; s += 32) {
        const double xs = x[s];
#pragma unroll
        for (int k = 0; k < G; ++k)
          if (r[k]) p[k] += (double)r[k][s] * xs;            // warp-uniform predicate
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < G; ++k) p[k] += __shfl_xor_sync(0xffffffffu, p[k], off);
      if (lane == 0)
#pragma unroll
        for (int k = 0; k < G; ++k)
          if (r[k]) emit(i0 + k, p[k]);
    }
    __syncthreads();
    return;
  }
#else
  for (int i = B.tid; i < n; i += B.nthr) {
    const T* r = rowf(i);
    if (!r) continue;
    double p = 0.;
    for (int s = 0; s < S; ++s) p += (double)r[s] * x[s];
    emit(i, p);
  }
  B.sync();
#endif
}

// giga.py:33-38 evaluated in float64 on the two row/direction inner products
BCG_HD double giga_score64(double s0, double s1) {
  const bool ok = (s1 > -1. + 1e-14) && (1. - s1 * s1 > 0.);
  const double den = ok ? sqrt(1. - s1 * s1) : INFINITY;
  return s0 / den;
}

BCG_HD void push_event(SolverState* st, int code, int64_t f, int nact, double err, double a0, double a1) {
  bcg_iter_event* e = st->events + st->n_events;
  e->code = code; e->nact = nact; e->f = f; e->error = err; e->aux0 = a0; e->aux1 = a1;
  st->n_events += 1;
}

// snnls.py:63-72 -- first failure: retry; second consecutive failure: latch the numeric limit
BCG_HD void fail_event(SolverState* st, int code, int64_t f, double a0, double a1) {
  push_event(st, code, f, st->nact, st->err, a0, a1);
  if (st->retried) st->halted = 1; else st->retried = 1;
}

BCG_HD double row_score64(const Blk& B, const SolverState* st, const float* row) {
  const int S = st->S;
  double v[2] = {0., 0.};
  const double* d0 = st->dir64;
  const double* d1 = st->dir64 + S;
  if (st->alg == BCG_ALG_GIGA) {
    for (int s = B.tid; s < S; s += B.nthr) { double x = (double)row[s]; v[0] += x * d0[s]; v[1] += x * d1[s]; }
    blk_sum<2>(B, v);
    return giga_score64(v[0], v[1]);
  }
  for (int s = B.tid; s < S; s += B.nthr) v[0] += (double)row[s] * d0[s];
  blk_sum<1>(B, v);
  return v[0];
}

// Resolve the best LOCAL row from the per-warp scan candidates.  Candidates whose float32 score
// is within delta of the maximum are re-scored in float64 (ties -> lowest row index, as
// ndarray.argmax).  force_rescore: always produce the float64 score (needed for cross-rank and
// OMP comparisons).
// unit row and norm of LOCAL row lrow: a row of the resident matrix, or -- never-materialising solver -- the row the
// selection pass just re-evaluated (it can only be the winner)
BCG_HD void local_row(const SolverState* st, uint32_t lrow, const float** row, double* norm) {
  if (st->lazy) { *row = st->wrow; *norm = st->wnorm; return; }
  *row = st->An + (size_t)lrow * st->ld;
  *norm = st->norms[lrow];
}

BCG_HD void pick_local(const Blk& B, SolverState* st, bool force_rescore, uint32_t* row_out, double* score_out) {
  if (st->lazy) {                     // lazy_select_kernel already reduced the grid in float64
    *row_out = st->fused_row < 0 ? kNoRow : (uint32_t)st->fused_row;
    *score_out = st->fused_score;
    return;
  }
  if (st->need_exact) {
    // the float32 candidate set was ambiguous (SolverState::cand_lost): the local winner is the result of the exact
    // float64 pass over all local rows (exact_scan_kernel), one candidate per CTA -> best score, lowest row on ties
    double key = -INFINITY; int64_t id = -1; int pl = 0;
    for (int i = B.tid; i < st->n_exact_cands; i += B.nthr) {
      const ExactCand x = st->exact_cands[i];
      if (x.row < 0) continue;
      if (id < 0 || x.score > key || (x.score == key && x.row < id)) { key = x.score; id = x.row; }
    }
    blk_argbest(B, &key, &id, &pl);
    B.sync();
    if (B.tid == 0) { st->need_exact = 0; st->n_exact += 1; }
    B.sync();
    *row_out = id < 0 ? kNoRow : (uint32_t)id;
    *score_out = key;
    return;
  }
  if (st->scan_cnt == 1) {
    // the scan's last CTA found exactly ONE published candidate inside the near-tie window (and no unpublished one):
    // it is the arg-max -- no block-wide reduction over the candidates (~7 us of an OMP iteration)
    const uint32_t r = st->scan_top_row;
    const double sc = force_rescore ? row_score64(B, st, st->An + (size_t)r * st->ld) : (double)st->scan_top;
    B.sync();
    if (B.tid == 0) st->scan_cnt = 0;                     // consumed
    B.sync();
    *row_out = r; *score_out = sc;
    return;
  }
  ScanCand* c = st->cands;
  const int n = st->n_cands;
  uint32_t chosen[kRescoreMax];
  float chosen_s[kRescoreMax];
  int nch = 0;
  float top = 0.f;
  for (int r = 0; r < kRescoreMax; ++r) {
    double bs = -INFINITY; int64_t br = -1; int bi = -1;
    for (int i = B.tid; i < n; i += B.nthr) {
      const ScanCand x = c[i];
      if (x.row == kNoRow) continue;
      if (br < 0 || (double)x.score > bs || ((double)x.score == bs && (int64_t)x.row < br)) {
        bs = (double)x.score; br = (int64_t)x.row; bi = i;
      }
    }
    blk_argbest(B, &bs, &br, &bi);
    if (br < 0) break;
    if (r == 0) top = (float)bs;
    else if (!((float)bs >= top - (2e-5f + 1e-5f * fabsf(top)))) break;
    chosen[nch] = (uint32_t)br; chosen_s[nch] = (float)bs; ++nch;
    if (B.tid == 0) c[bi].row = kNoRow;     // exclude from the next round
    B.sync();
  }
  if (nch == 0) { *row_out = kNoRow; *score_out = -INFINITY; return; }
  if (nch == 1 && !force_rescore) { *row_out = chosen[0]; *score_out = (double)chosen_s[0]; return; }
  uint32_t best_row = kNoRow; double best = -INFINITY;
  for (int r = 0; r < nch; ++r) {
    const double sc = row_score64(B, st, st->An + (size_t)chosen[r] * st->ld);
    if (best_row == kNoRow || sc > best || (sc == best && chosen[r] < best_row)) { best = sc; best_row = chosen[r]; }
  }
  *row_out = best_row; *score_out = best;
}

#ifdef __CUDA_ARCH__
// all-to-all candidate exchange over NVLink peer memory (defined in kernels.cuh)
__device__ inline void mail_exchange(const Blk& B, SolverState* st, uint32_t lrow, double lscore, int64_t* f,
                              double* norm, const float** row);
#endif

BCG_HD void count_positive(const Blk& B, const SolverState* st, bool* nonempty) {
  double cnt = 0.;
  for (int k = B.tid; k < st->nact; k += B.nthr) cnt += (st->act_w[k] > 0.) ? 1. : 0.;
  blk_sum<1>(B, &cnt);
  *nonempty = cnt > 0.;
}

// find f in the stored active set; returns its slot or -1
BCG_HD int find_slot(const Blk& B, const SolverState* st, int64_t f) {
  double key = -INFINITY; int64_t id = -1; int pl = -1;
  for (int k = B.tid; k < st->nact; k += B.nthr)
    if (st->act_idx[k] == f && (id < 0 || k < id)) { key = 0.; id = k; pl = k; }
  blk_argbest(B, &key, &id, &pl);
  return id < 0 ? -1 : (int)id;
}

// xw_new = sum_k w_new[k] * norm[k] * row_k ; returns ||xw_new - b||
BCG_HD double recompute_iterate(const Blk& B, SolverState* st, const double* w, int nact, double* xw_out) {
  const int S = st->S, ld = st->ld;
  double e = 0.;
  for (int s = B.tid; s < S; s += B.nthr) {
    double acc = 0.;
    for (int k = 0; k < nact; ++k) acc += (w[k] * st->act_norm[k]) * (double)st->act_rows[(size_t)k * ld + s];
    xw_out[s] = acc;
    const double r = acc - st->b[s];
    e += r * r;
  }
  blk_sum<1>(B, &e);
  return sqrt(e);
}

// ---------------------------------------------------------------------------------------------
// finish one GIGA / Frank-Wolfe iteration: snnls.py:44-62 around giga.py:40-64 / frankwolfe.py:19-40
// ---------------------------------------------------------------------------------------------
BCG_HD void finish_iteration(const Blk& B, SolverState* st) {
  const int S = st->S, ld = st->ld;
  bool nonempty;
  count_positive(B, st, &nonempty);          // sampled BEFORE the step (snnls.py:44)
  const double prev_err = st->err;

  if (st->select_failed) {                   // giga.py:28-29 raised inside _select
    B.sync();
    if (B.tid == 0) fail_event(st, BCG_IT_FAIL_CDIR, -1, st->sel_aux, 0.);
    B.sync();
    return;
  }

  uint32_t lrow; double lscore;
  pick_local(B, st, st->world > 1, &lrow, &lscore);
  if (st->world == 1 && lrow == kNoRow) {    // no comparable score at all (non-finite matrix entries)
    B.sync();
    if (B.tid == 0) { st->comm_error = 2; st->halted = 1; }
    B.sync();
    return;
  }

  int64_t f; double nf_stored; const float* frow;
  if (st->world > 1) {
#ifdef __CUDA_ARCH__
    mail_exchange(B, st, lrow, lscore, &f, &nf_stored, &frow);
    if (st->comm_error) return;
#else
    return;
#endif
  } else {
    f = st->row_offset + (int64_t)lrow;
    local_row(st, lrow, &frow, &nf_stored);
  }

  for (int s = B.tid; s < S; s += B.nthr) st->xf[s] = nf_stored * (double)frow[s];
  B.sync();

  double alpha, beta;
  if (st->alg == BCG_ALG_GIGA) {
    double v[5] = {0., 0., 0., 0., 0.};
    for (int s = B.tid; s < S; s += B.nthr) {
      const double xw = st->xw[s], xf = st->xf[s], bn = st->bn[s];
      v[0] += xw * xw; v[1] += xf * xf; v[2] += bn * xw; v[3] += bn * xf; v[4] += xw * xf;
    }
    blk_sum<5>(B, v);
    double nw = sqrt(v[0]);
    if (nw == 0.) nw = 1.;                   // giga.py:43-44
    const double nf = sqrt(v[1]);
    const double bxf = v[3] / nf, bxw = v[2] / nw, xwxf = v[4] / (nw * nf);
    const double gA = bxf - bxw * xwxf;
    const double gB = bxw - bxf * xwxf;
    if (gA <= 0. || gB < 0.) {               // giga.py:50-51
      if (B.tid == 0) fail_event(st, BCG_IT_FAIL_GEODESIC, f, gA, gB);
      B.sync();
      return;
    }
    const double a = gB / (gA + gB) / nw;
    const double bb = gA / (gA + gB) / nf;
    double u[2] = {0., 0.};
    for (int s = B.tid; s < S; s += B.nthr) {
      const double x = a * st->xw[s] + bb * st->xf[s];
      u[0] += x * x; u[1] += x * st->bn[s];
    }
    blk_sum<2>(B, u);
    const double nx = sqrt(u[0]);
    const double scale = st->bnorm / nx * (u[1] / nx);   // giga.py:58
    alpha = a * scale;
    beta = bb * scale;
  } else {  // Frank-Wolfe
    if (!nonempty) {                         // frankwolfe.py:20-23
      alpha = 0.;
      beta = st->nsum / nf_stored;
    } else {
      const double r = st->nsum / nf_stored;
      double v[2] = {0., 0.};
      for (int s = B.tid; s < S; s += B.nthr) {
        const double t = r * st->xf[s] - st->xw[s];
        v[0] += t * (st->b[s] - st->xw[s]); v[1] += t * t;
      }
      blk_sum<2>(B, v);
      const double gnum = v[0], gden = v[1];
      if (gnum < 0. || gden == 0. || gnum > gden) {      // frankwolfe.py:33-34
        if (B.tid == 0) fail_event(st, BCG_IT_FAIL_GAMMA, f, gnum, gden);
        B.sync();
        return;
      }
      alpha = 1. - gnum / gden;
      beta = r * gnum / gden;
    }
  }

  // w <- alpha w ; w[f] <- max(0, w[f] + beta)      (giga.py:63-64, frankwolfe.py:39-40)
  int slot = find_slot(B, st, f);
  int nact_new = st->nact;
  if (slot < 0) {
    slot = st->nact;
    nact_new = st->nact + 1;
    for (int s = B.tid; s < ld; s += B.nthr) st->act_rows[(size_t)slot * ld + s] = frow[s];
    if (B.tid == 0) { st->act_idx[slot] = f; st->act_norm[slot] = nf_stored; }
  }
  for (int k = B.tid; k < nact_new; k += B.nthr) {
    const double wk = (k < st->nact) ? alpha * st->act_w[k] : 0.;
    st->act_w_new[k] = (k == slot) ? fmax(0., wk + beta) : wk;
  }
  B.sync();

  const double err_new = recompute_iterate(B, st, st->act_w_new, nact_new, st->xw_new);

  if (st->check_monotone && nonempty && err_new > prev_err) {      // snnls.py:56-61: revert (nothing was committed)
    if (B.tid == 0) fail_event(st, BCG_IT_FAIL_MONOTONE, f, err_new, prev_err);
    B.sync();
    return;
  }
  for (int k = B.tid; k < nact_new; k += B.nthr) st->act_w[k] = st->act_w_new[k];
  for (int s = B.tid; s < S; s += B.nthr) st->xw[s] = st->xw_new[s];
  B.sync();
  if (B.tid == 0) {
    st->nact = nact_new;
    st->err = err_new;
    if (st->check_monotone && nonempty) st->retried = 0;           // snnls.py:62 (inside the monotone branch)
    push_event(st, BCG_IT_OK, f, nact_new, err_new, lscore, 0.);
  }
  B.sync();
}

// ---------------------------------------------------------------------------------------------
// direction(s) for the next scan: giga.py:21-30 / frankwolfe.py:16 / orthopursuit.py:18
// ---------------------------------------------------------------------------------------------
BCG_HD void prepare_select(const Blk& B, SolverState* st) {
  const int S = st->S, ld = st->ld;
  if (st->alg == BCG_ALG_GIGA) {
    double v[2] = {0., 0.};
    for (int s = B.tid; s < S; s += B.nthr) { const double x = st->xw[s]; v[0] += x * x; v[1] += st->bn[s] * x; }
    blk_sum<2>(B, v);
    double nw = sqrt(v[0]);
    if (nw == 0.) nw = 1.;
    const double bxw = v[1] / nw;            // bn . (xw/nw)
    double c2 = 0.;
    for (int s = B.tid; s < S; s += B.nthr) {
      const double xn = st->xw[s] / nw;
      const double cd = st->bn[s] - bxw * xn;
      st->dir64[s] = cd; st->dir64[S + s] = xn;
      c2 += cd * cd;
    }
    blk_sum<1>(B, &c2);
    const double cdirnrm = sqrt(c2);
    if (cdirnrm < st->tol) {                 // giga.py:28-29
      if (B.tid == 0) { st->select_failed = 1; st->sel_aux = cdirnrm; }
      B.sync();
      return;
    }
    for (int s = B.tid; s < ld; s += B.nthr) {
      double cd = 0., xn = 0.;
      if (s < S) { cd = st->dir64[s] / cdirnrm; xn = st->dir64[S + s]; st->dir64[s] = cd; }
      st->dir32[s] = (float)cd; st->dir32[ld + s] = (float)xn;
    }
  } else {
    double r2 = 0.;
    for (int s = B.tid; s < S; s += B.nthr) { const double r = st->b[s] - st->xw[s]; st->dir64[s] = r; r2 += r * r; }
    blk_sum<1>(B, &r2);
    const double rn = sqrt(r2);
    const double inv = rn > 0. ? 1. / rn : 1.;     // positive rescale: argmax / sign tests unchanged
    for (int s = B.tid; s < ld; s += B.nthr) {
      double r = 0.;
      if (s < S) { r = st->dir64[s] * inv; st->dir64[s] = r; }
      st->dir32[s] = (float)r;
    }
  }
  if (B.tid == 0) st->select_failed = 0;
  B.sync();
}

// ---------------------------------------------------------------------------------------------
// OMP selection: orthopursuit.py:17-35, then w[f] = 1 (orthopursuit.py:38).  Result in st->seq_f.
// ---------------------------------------------------------------------------------------------
BCG_HD int64_t omp_select(const Blk& B, SolverState* st) {
  const int S = st->S, ld = st->ld;
  bool nonempty;
  count_positive(B, st, &nonempty);
  uint32_t lrow; double pos;
  pick_local(B, st, true, &lrow, &pos);
  omp_mark(B, st, 2);
  if (st->world == 1 && lrow == kNoRow) {    // no comparable score at all (non-finite matrix entries)
    B.sync();
    if (B.tid == 0) { st->comm_error = 2; st->halted = 1; }
    B.sync();
    return -1;
  }
  int64_t f; double nf_stored; const float* frow;
  if (st->world > 1) {
#ifdef __CUDA_ARCH__
    mail_exchange(B, st, lrow, pos, &f, &nf_stored, &frow);
    if (st->comm_error) return -1;
    pos = row_score64(B, st, frow);
#else
    return -1;
#endif
  } else {
    f = st->row_offset + (int64_t)lrow;
    local_row(st, lrow, &frow, &nf_stored);
  }
  // orthopursuit.py:26-35 compares pos with neg = max over the active set of -<a_k, residual>.  Right after an NNLS solve the
  // weights are the optimum of the active set, so those inner products are zero up to the backward error of the solve
  // (<= ~1e-9 of the unit residual while ||r|| > 1e-6 ||b||): with pos > 1e-3 the comparison is decided and the K x S pass is
  // skipped (7-15 us of the iteration).  Any other state (host-written weights, reverted step, near convergence) takes it.
  const bool neg_decided = st->kkt_valid && pos > 1e-3 && st->err > 1e-6 * st->bnorm;
  if (nonempty && !neg_decided) {
    // negative direction over the active set (w > 0), lowest global index wins ties:
    // -dots over the active rows (coalesced, several rows in flight per warp), then the block arg-max
    blk_dots<float>(B, st->nact, S,
                    [&](int k) { return st->act_w[k] > 0. ? st->act_rows + (size_t)k * ld : (const float*)nullptr; },
                    st->dir64, [&](int k, double d) { st->act_tmp[k] = -d; });
    double key = -INFINITY; int64_t id = -1; int pl = -1;
    for (int k = B.tid; k < st->nact; k += B.nthr) {
      if (!(st->act_w[k] > 0.)) continue;
      const double ng = st->act_tmp[k];
      const int64_t gi = st->act_idx[k];
      if (id < 0 || ng > key || (ng == key && gi < id)) { key = ng; id = gi; pl = k; }
    }
    blk_argbest(B, &key, &id, &pl);
    omp_mark(B, st, 3);
    if (id >= 0 && !(pos >= key)) {          // orthopursuit.py:32-35
      f = id;
      if (B.tid == 0) st->act_w[pl] = 1.;
      B.sync();
      return f;
    }
  }
  if (!nonempty || neg_decided) omp_mark(B, st, 3);
  const int nact0 = st->nact;              // read by every thread BEFORE the barriers of find_slot: thread 0 writes st->nact below
  int slot = find_slot(B, st, f);
  if (slot < 0) {
    slot = nact0;
    for (int s = B.tid; s < ld; s += B.nthr) st->act_rows[(size_t)slot * ld + s] = frow[s];
    if (B.tid == 0) { st->act_idx[slot] = f; st->act_norm[slot] = nf_stored; st->nact = slot + 1; }
  }
  B.sync();
  if (B.tid == 0) st->act_w[slot] = 1.;
  B.sync();
  return f;
}

// weights were overwritten by the host (NNLS write-back): recompute A w and the error
BCG_HD void refresh_iterate(const Blk& B, SolverState* st) {
  const double e = recompute_iterate(B, st, st->act_w, st->nact, st->xw);
  B.sync();
  if (B.tid == 0) st->err = e;
  B.sync();
}

}  // namespace bcg
