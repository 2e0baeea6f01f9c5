// TEST-ONLY logic check (never part of the product library): compiles the single-block step logic
// of bayesian-coresets_b200/csrc/step_logic.h for the host (one "thread", Blk{0,1}) and drives it
// with a plain-loop float32 stand-in for the scan kernel, so the iteration state machine (winner
// resolution, reweight, monotone check, retry / latch) can be compared with the oracle on a box
// without a GPU.  The CUDA kernels themselves are only ever validated on the GPU (tests -m gpu).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../bayesian-coresets_b200/csrc/step_logic.h"

using namespace bcg;

namespace {
struct Host {
  SolverState st;
  std::vector<float> An, dir32, wrow, act_rows;
  std::vector<double> norms, b, bn, xw, xw_new, xf, dir64, act_w, act_w_new, act_norm, act_tmp;
  std::vector<int64_t> act_idx;
  std::vector<ScanCand> cands;
  std::vector<float> lost;
  ExactCand exact;
  std::vector<bcg_iter_event> events;
  double sred[256];
};

// float32 scan stand-in: same per-row arithmetic as scan_kernel.cuh (fmaf accumulation, float32
// GIGA epilogue), candidates bucketed over `nb` interleaved row groups, first maximum wins
void scan(Host& H, int nb) {
  SolverState& st = H.st;
  H.cands.assign(nb, ScanCand{-INFINITY, kNoRow});
  H.lost.assign(nb, -INFINITY);
  st.check_monotone = 1;
  st.exact_cands = &H.exact; st.n_exact_cands = 1;
  if (st.halted || st.select_failed) return;
  for (int64_t r = 0; r < st.n_local; ++r) {
    const float* row = &H.An[(size_t)r * st.ld];
    float a0 = 0.f, a1 = 0.f;
    for (int s = 0; s < st.ld; ++s) {
      a0 = fmaf(row[s], H.dir32[s], a0);
      if (st.alg == BCG_ALG_GIGA) a1 = fmaf(row[s], H.dir32[st.ld + s], a1);
    }
    float score = a0;
    if (st.alg == BCG_ALG_GIGA) {
      const float den = 1.f - a1 * a1;
      score = (a1 > -1.f && den > 0.f) ? a0 / sqrtf(den) : 0.f;
    }
    ScanCand& c = H.cands[r % nb];
    float& lost = H.lost[r % nb];
    if (c.row == kNoRow || score > c.score) { if (c.row != kNoRow) lost = fmaxf(lost, c.score); c.score = score; c.row = (uint32_t)r; }
    else lost = fmaxf(lost, score);
  }
  // the exactness check of scan_kernel's last CTA and, when it fires, the exact float64 pass (exact_scan_kernel)
  float top = -INFINITY, lm = -INFINITY;
  for (int i = 0; i < nb; ++i) { if (H.cands[i].row != kNoRow) top = fmaxf(top, H.cands[i].score); lm = fmaxf(lm, H.lost[i]); }
  const float thr = top - (2e-5f + 1e-5f * fabsf(top));
  int cnt = 0;
  for (int i = 0; i < nb; ++i) cnt += (H.cands[i].row != kNoRow && H.cands[i].score >= thr) ? 1 : 0;
  st.scan_top = top; st.scan_cnt = 0; st.scan_top_row = kNoRow;
  for (int i = 0; i < nb; ++i)
    if (H.cands[i].row != kNoRow && H.cands[i].score == top && H.cands[i].row < st.scan_top_row) st.scan_top_row = H.cands[i].row;
  if (!(top > -INFINITY && (lm >= thr || cnt > kRescoreMax || st.force_exact))) st.scan_cnt = cnt;
  if (top > -INFINITY && (lm >= thr || cnt > kRescoreMax || st.force_exact)) {
    st.need_exact = 1;
    H.exact.score = -INFINITY; H.exact.row = -1;
    for (int64_t r = 0; r < st.n_local; ++r) {
      const float* row = &H.An[(size_t)r * st.ld];
      double v0 = 0., v1 = 0.;
      for (int s = 0; s < st.S; ++s) {
        v0 += (double)row[s] * H.dir64[s];
        if (st.alg == BCG_ALG_GIGA) v1 += (double)row[s] * H.dir64[st.S + s];
      }
      const double sc = st.alg == BCG_ALG_GIGA ? giga_score64(v0, v1) : v0;
      if (sc > H.exact.score) { H.exact.score = sc; H.exact.row = r; }
    }
  }
}
}  // namespace

extern "C" int hostcheck_run(int alg, const float* An, const double* norms, const double* b, int S, int ld,
                             int64_t N, double nsum, int itrs, int builds, double tol, bcg_iter_event* events_out,
                             int* n_events_out, int64_t* idx_out, double* w_out, int* k_out, double* err_out,
                             int* halted_out) {
  Host H;
  memset(&H.st, 0, sizeof(SolverState));
  SolverState& st = H.st;
  const int cap = itrs * builds + 8;
  H.An.assign(An, An + (size_t)N * ld);
  H.norms.assign(norms, norms + N);
  H.b.assign(b, b + S);
  double bnorm = 0.;
  for (int i = 0; i < S; ++i) bnorm += b[i] * b[i];
  bnorm = sqrt(bnorm);
  H.bn.resize(S);
  for (int i = 0; i < S; ++i) H.bn[i] = bnorm > 0 ? b[i] / bnorm : 0.;
  H.xw.assign(S, 0.); H.xw_new.assign(S, 0.); H.xf.assign(S, 0.); H.dir64.assign(2 * S, 0.);
  H.dir32.assign(2 * ld, 0.f); H.wrow.assign(ld, 0.f);
  H.act_rows.assign((size_t)cap * ld, 0.f);
  H.act_w.assign(cap, 0.); H.act_w_new.assign(cap, 0.); H.act_norm.assign(cap, 0.); H.act_tmp.assign(cap, 0.); H.act_idx.assign(cap, -1);
  H.events.resize((size_t)itrs * builds);
  st.alg = alg; st.S = S; st.ld = ld; st.world = 1; st.rank = 0;
  st.n_local = N; st.row_offset = 0; st.n_global = N; st.tol = tol; st.bnorm = bnorm; st.nsum = nsum;
  st.An = H.An.data(); st.norms = H.norms.data(); st.b = H.b.data(); st.bn = H.bn.data();
  st.xw = H.xw.data(); st.xw_new = H.xw_new.data(); st.xf = H.xf.data(); st.dir64 = H.dir64.data();
  st.dir32 = H.dir32.data(); st.wrow = H.wrow.data(); st.err = bnorm;
  st.cap = cap; st.act_idx = H.act_idx.data(); st.act_w = H.act_w.data(); st.act_w_new = H.act_w_new.data();
  st.act_norm = H.act_norm.data(); st.act_tmp = H.act_tmp.data(); st.act_rows = H.act_rows.data();
  st.events = H.events.data();
  const int nb = 37;
  Blk B{0, 1, H.sred, nullptr};
  for (int bld = 0; bld < builds && !st.halted; ++bld) {
    st.retried = 0;                       // snnls.py:40: local to each build() call
    prepare_select(B, &st);
    for (int i = 0; i < itrs; ++i) {
      scan(H, nb);
      st.cands = H.cands.data(); st.n_cands = nb;
      if (st.halted) break;
      finish_iteration(B, &st);
      if (st.halted) break;
      if (i + 1 < itrs) prepare_select(B, &st);
    }
  }
  memcpy(events_out, H.events.data(), sizeof(bcg_iter_event) * st.n_events);
  *n_events_out = st.n_events;
  for (int k = 0; k < st.nact; ++k) { idx_out[k] = st.act_idx[k]; w_out[k] = st.act_w[k]; }
  *k_out = st.nact;
  *err_out = st.err;
  *halted_out = st.halted;
  return 0;
}

// OMP selection stand-in: returns the selected index sequence driven by host-provided weights
extern "C" int hostcheck_omp_select(const float* An, const double* norms, const double* b, int S, int ld, int64_t N,
                                    const int64_t* act_idx, const double* act_w, int nact, int64_t* f_out) {
  Host H;
  memset(&H.st, 0, sizeof(SolverState));
  SolverState& st = H.st;
  const int cap = nact + 2;
  H.An.assign(An, An + (size_t)N * ld);
  H.norms.assign(norms, norms + N);
  H.b.assign(b, b + S);
  H.bn.assign(S, 0.);
  H.xw.assign(S, 0.); H.xw_new.assign(S, 0.); H.xf.assign(S, 0.); H.dir64.assign(2 * S, 0.);
  H.dir32.assign(2 * ld, 0.f); H.wrow.assign(ld, 0.f);
  H.act_rows.assign((size_t)cap * ld, 0.f);
  H.act_w.assign(cap, 0.); H.act_w_new.assign(cap, 0.); H.act_norm.assign(cap, 0.); H.act_tmp.assign(cap, 0.); H.act_idx.assign(cap, -1);
  st.alg = BCG_ALG_OMP; st.S = S; st.ld = ld; st.world = 1;
  st.n_local = N; st.n_global = N; st.tol = 1e-12;
  st.An = H.An.data(); st.norms = H.norms.data(); st.b = H.b.data(); st.bn = H.bn.data();
  st.xw = H.xw.data(); st.xw_new = H.xw_new.data(); st.xf = H.xf.data(); st.dir64 = H.dir64.data();
  st.dir32 = H.dir32.data(); st.wrow = H.wrow.data();
  st.cap = cap; st.act_idx = H.act_idx.data(); st.act_w = H.act_w.data(); st.act_w_new = H.act_w_new.data();
  st.act_norm = H.act_norm.data(); st.act_tmp = H.act_tmp.data(); st.act_rows = H.act_rows.data();
  for (int k = 0; k < nact; ++k) {
    st.act_idx[k] = act_idx[k]; st.act_w[k] = act_w[k]; st.act_norm[k] = norms[act_idx[k]];
    memcpy(&H.act_rows[(size_t)k * ld], &An[(size_t)act_idx[k] * ld], sizeof(float) * ld);
  }
  st.nact = nact;
  Blk B{0, 1, H.sred, nullptr};
  refresh_iterate(B, &st);
  prepare_select(B, &st);
  scan(H, 37);
  st.cands = H.cands.data(); st.n_cands = 37;
  *f_out = omp_select(B, &st);
  return 0;
}

// ---- NNLS logic check ------------------------------------------------------------------------
#include "../../bayesian-coresets_b200/csrc/nnls_logic.h"

namespace {
struct NnlsHost {
  Host H;
  NnlsWork W;
  std::vector<double> Q, R, R2, rot, c, z, wP, h, v;
  std::vector<int32_t> P, Z, inP, rem;
};
}  // namespace

// sequence of warm-started solves: at step t the columns cols[0..counts[t]) are in the problem; columns whose
// weight is positive after step t-1 keep it, new columns enter with weight 1 (what omp_select does).
// w_out: steps x ncols solutions.
extern "C" int hostcheck_nnls_sequence(const float* An, const double* norms, const double* b, int S, int ld, int64_t N,
                                       const int64_t* cols, int ncols, const int* counts, int steps, int from_scratch,
                                       int downdate, double* w_out, int* rebuilds_out) {
  NnlsHost X;
  Host& H = X.H;
  memset(&H.st, 0, sizeof(SolverState));
  SolverState& st = H.st;
  const int cap = ncols + 2;
  H.An.assign(An, An + (size_t)N * ld);
  H.norms.assign(norms, norms + N);
  H.b.assign(b, b + S);
  H.bn.assign(S, 0.); H.xw.assign(S, 0.); H.xw_new.assign(S, 0.); H.xf.assign(S, 0.); H.dir64.assign(2 * S, 0.);
  H.dir32.assign(2 * ld, 0.f); H.wrow.assign(ld, 0.f);
  H.act_rows.assign((size_t)cap * ld, 0.f);
  H.act_w.assign(cap, 0.); H.act_w_new.assign(cap, 0.); H.act_norm.assign(cap, 0.); H.act_tmp.assign(cap, 0.); H.act_idx.assign(cap, -1);
  st.alg = BCG_ALG_OMP; st.S = S; st.ld = ld; st.world = 1; st.n_local = N; st.n_global = N;
  for (int i = 0; i < S; ++i) st.bnorm += b[i] * b[i];
  st.bnorm = sqrt(st.bnorm);
  st.An = H.An.data(); st.norms = H.norms.data(); st.b = H.b.data(); st.bn = H.bn.data();
  st.xw = H.xw.data(); st.xw_new = H.xw_new.data(); st.xf = H.xf.data(); st.dir64 = H.dir64.data();
  st.dir32 = H.dir32.data(); st.wrow = H.wrow.data();
  st.cap = cap; st.act_idx = H.act_idx.data(); st.act_w = H.act_w.data(); st.act_w_new = H.act_w_new.data();
  st.act_norm = H.act_norm.data(); st.act_tmp = H.act_tmp.data(); st.act_rows = H.act_rows.data();
  X.Q.assign((size_t)cap * S, 0.); X.R.assign((size_t)cap * cap, 0.); X.c.assign(cap, 0.); X.z.assign(2 * cap, 0.);
  X.wP.assign(cap, 0.); X.h.assign(cap, 0.); X.v.assign(S, 0.); X.P.assign(cap, 0); X.Z.assign(cap, 0); X.inP.assign(cap, 0);
  NnlsWork& W = X.W;
  memset(&W, 0, sizeof(W));
  W.Q = X.Q.data(); W.R = X.R.data(); W.c = X.c.data(); W.z = X.z.data(); W.wP = X.wP.data(); W.h = X.h.data();
  W.v = X.v.data(); W.P = X.P.data(); W.Z = X.Z.data(); W.inP = X.inP.data(); W.cap = cap; W.tld = cap;
  X.R2.assign((size_t)cap * cap, 0.); X.rot.assign((size_t)3 * cap, 0.); X.rem.assign(cap, 0);
  W.R2 = X.R2.data(); W.rot = X.rot.data(); W.rem = X.rem.data(); W.downdate = downdate;
  Blk B{0, 1, H.sred, nullptr};
  int have = 0, reb = 0;
  for (int t = 0; t < steps; ++t) {
    for (; have < counts[t]; ++have) {
      st.act_idx[have] = cols[have]; st.act_norm[have] = norms[cols[have]]; st.act_w[have] = 1.;
      memcpy(&H.act_rows[(size_t)have * ld], &An[(size_t)cols[have] * ld], sizeof(float) * ld);
    }
    st.nact = have;
    nnls_solve(B, &st, &W, from_scratch);
    reb += W.rebuilds;
    for (int k = 0; k < ncols; ++k) w_out[(size_t)t * ncols + k] = k < have ? st.act_w[k] : 0.;
  }
  *rebuilds_out = reb;
  return 0;
}

// full OrthoPursuit loop: scan stand-in + omp_iteration (selection, warm-started NNLS, monotone check, events,
// next direction) exactly as bcg_solver_build drives omp_iteration_kernel
extern "C" int hostcheck_run_omp(const float* An, const double* norms, const double* b, int S, int ld, int64_t N,
                                 int itrs, int builds, int downdate, bcg_iter_event* events_out, int* n_events_out, int64_t* idx_out,
                                 double* w_out, int* k_out, double* err_out, int* halted_out) {
  NnlsHost X;
  Host& H = X.H;
  memset(&H.st, 0, sizeof(SolverState));
  SolverState& st = H.st;
  const int cap = itrs * builds + 8;
  H.An.assign(An, An + (size_t)N * ld);
  H.norms.assign(norms, norms + N);
  H.b.assign(b, b + S);
  double bnorm = 0.;
  for (int i = 0; i < S; ++i) bnorm += b[i] * b[i];
  bnorm = sqrt(bnorm);
  H.bn.assign(S, 0.); H.xw.assign(S, 0.); H.xw_new.assign(S, 0.); H.xf.assign(S, 0.); H.dir64.assign(2 * S, 0.);
  H.dir32.assign(2 * ld, 0.f); H.wrow.assign(ld, 0.f);
  H.act_rows.assign((size_t)cap * ld, 0.f);
  H.act_w.assign(cap, 0.); H.act_w_new.assign(cap, 0.); H.act_norm.assign(cap, 0.); H.act_tmp.assign(cap, 0.); H.act_idx.assign(cap, -1);
  H.events.resize((size_t)itrs * builds);
  st.alg = BCG_ALG_OMP; st.S = S; st.ld = ld; st.world = 1; st.n_local = N; st.n_global = N; st.tol = 1e-12;
  st.bnorm = bnorm; st.err = bnorm;
  st.An = H.An.data(); st.norms = H.norms.data(); st.b = H.b.data(); st.bn = H.bn.data();
  st.xw = H.xw.data(); st.xw_new = H.xw_new.data(); st.xf = H.xf.data(); st.dir64 = H.dir64.data();
  st.dir32 = H.dir32.data(); st.wrow = H.wrow.data();
  st.cap = cap; st.act_idx = H.act_idx.data(); st.act_w = H.act_w.data(); st.act_w_new = H.act_w_new.data();
  st.act_norm = H.act_norm.data(); st.act_tmp = H.act_tmp.data(); st.act_rows = H.act_rows.data();
  st.events = H.events.data();
  X.Q.assign((size_t)cap * S, 0.); X.R.assign((size_t)cap * cap, 0.); X.c.assign(cap, 0.); X.z.assign(2 * cap, 0.);
  X.wP.assign(cap, 0.); X.h.assign(cap, 0.); X.v.assign(S, 0.); X.P.assign(cap, 0); X.Z.assign(cap, 0); X.inP.assign(cap, 0);
  NnlsWork& W = X.W;
  memset(&W, 0, sizeof(W));
  W.Q = X.Q.data(); W.R = X.R.data(); W.c = X.c.data(); W.z = X.z.data(); W.wP = X.wP.data(); W.h = X.h.data();
  W.v = X.v.data(); W.P = X.P.data(); W.Z = X.Z.data(); W.inP = X.inP.data(); W.cap = cap; W.tld = cap;
  X.R2.assign((size_t)cap * cap, 0.); X.rot.assign((size_t)3 * cap, 0.); X.rem.assign(cap, 0);
  W.R2 = X.R2.data(); W.rot = X.rot.data(); W.rem = X.rem.data(); W.downdate = downdate;
  Blk B{0, 1, H.sred, nullptr};
  const int nb = 37;
  for (int bld = 0; bld < builds && !st.halted; ++bld) {
    st.retried = 0;
    prepare_select(B, &st);
    for (int i = 0; i < itrs; ++i) {
      scan(H, nb);
      st.cands = H.cands.data(); st.n_cands = nb;
      omp_iteration(B, &st, &W, (i + 1 < itrs) ? 1 : 0);
    }
  }
  memcpy(events_out, H.events.data(), sizeof(bcg_iter_event) * st.n_events);
  *n_events_out = st.n_events;
  for (int k = 0; k < st.nact; ++k) { idx_out[k] = st.act_idx[k]; w_out[k] = st.act_w[k]; }
  *k_out = st.nact;
  *err_out = st.err;
  *halted_out = st.halted;
  return 0;
}

// ---- table-driven link functions of the projection kernels (csrc/softplus_table.h), host build ----------------
#include "../../bayesian-coresets_b200/csrc/softplus_table.h"

// model: 0 softplus_neg(t = lin), 1 LR link, 2 Poisson link, 3 / 4 the branch-free LR / Poisson forms (|lin| <= 37)
extern "C" int hostcheck_link(int model, const double* lin, const double* y, int64_t n, double* out) {
  static std::vector<double> tab;
  if (tab.empty()) { tab.resize(kSpTableDoubles); softplus_table_build(tab.data()); }
  for (int64_t i = 0; i < n; ++i) {
    if (model == 0) out[i] = softplus_neg(tab.data(), lin[i]);
    else if (model == 1) out[i] = lr_link_fast(tab.data(), lin[i]);
    else if (model == 3) out[i] = lr_link_nb(tab.data(), lin[i]);
    else if (model == 4) out[i] = poisson_link_nb(tab.data(), lin[i], y[i]);
    else out[i] = poisson_link_fast(tab.data(), lin[i], y[i]);
  }
  return 0;
}

// ---- bounds of the float16 pre-filter (csrc/filter_bounds.h), host build ------------------------------------------
#include "../../bayesian-coresets_b200/csrc/filter_bounds.h"

// For every row: the float32 score exactly as scan_core.cuh evaluates it (two summation orders: `order` 0 = left to
// right, 1 = 32 interleaved partial sums added pairwise -- the kernel's lane layout), and [lb, ub] from the float16-rounded
// row.  1.f / sqrtf stands in for rsqrtf (2 ulp on the device; covered by kScoreRel).
extern "C" int hostcheck_filter_bounds(const float* An, int64_t n, int ld, int S, const float* d0, const float* d1, int giga,
                                       int order, float* score32, float* lb, float* ub) {
  auto dot = [&](const float* x, const float* d, bool half) {
    float part[32];
    for (int i = 0; i < 32; ++i) part[i] = 0.f;
    float seq = 0.f;
    for (int s = 0; s < S; ++s) {
      float v = x[s];
      if (half) v = (float)(_Float16)v;                   // round to nearest even, as __floats2half2_rn
      if (order == 0) seq = fmaf(v, d[s], seq);
      else part[(s / 4) % 32] = fmaf(v, d[s], part[(s / 4) % 32]);
    }
    if (order == 0) return seq;
    for (int off = 16; off > 0; off >>= 1)
      for (int i = 0; i < off; ++i) part[i] += part[i + off];
    return part[0];
  };
  double q0 = 0., q1 = 0.;
  for (int s = 0; s < S; ++s) { q0 += (double)d0[s] * d0[s]; if (giga) q1 += (double)d1[s] * d1[s]; }
  const float eu = bcg::filter_eps_unit(S);
  const float e0 = eu * (float)sqrt(q0), e1 = eu * (float)sqrt(q1);
  for (int64_t r = 0; r < n; ++r) {
    const float* x = An + r * ld;
    const float s0 = dot(x, d0, false), t0 = dot(x, d0, true);
    if (giga) {
      const float s1 = dot(x, d1, false), t1 = dot(x, d1, true);
      const float den = 1.f - s1 * s1;
      score32[r] = (s1 > -1.f && den > 0.f) ? s0 * (1.f / sqrtf(den)) : 0.f;
      bcg::filter_bounds_giga(t0, t1, e0, e1, &lb[r], &ub[r]);
    } else {
      score32[r] = s0;
      bcg::filter_bounds_lin(t0, e0, &lb[r], &ub[r]);
    }
  }
  return 0;
}

extern "C" float hostcheck_filter_threshold(float L) { return bcg::filter_threshold(L); }
