// Non-negative least squares on the K active columns, on the device, in float64 -- the reweight of
// OrthoPursuit and the re-solve of SparseNNLS.optimize():
//   reference  snnls/orthopursuit.py:37-42   res = nnls(A[:, w>0], b); w[w>0] = res[0]
//              snnls/snnls.py:82-97          optimize()
//   (scipy.optimize.nnls = Lawson & Hanson's active-set method)
//
// Lawson-Hanson with a WARM START: the NNLS minimiser over linearly independent columns is unique,
// so instead of rebuilding the passive set from nothing at every OMP iteration (what the SciPy call
// does) the passive set P and its QR factorisation are kept between iterations.  In the common
// case the new column enters, one triangular solve gives an all-positive solution, and the
// iteration costs O(S K) (Gram-Schmidt append) + O(K^2) (back substitution) instead of O(S K^2).
// When a weight would turn negative the standard step-back / removal loop runs and the dropped column is
// removed from the factorisation by a sweep of Givens rotations (nnls_qr_remove, O((S + K) K)); the first
// version rebuilt the factorisation behind the removed position with one Gram-Schmidt append per column,
// which cost 0.3 - 4 ms on exactly the iterations that drop a column (ncu launch list, N = 1e6, S = 256: 79 of
// the 97 ms that 205 OMP iterations spent in this kernel).
//
// QR: classical Gram-Schmidt applied twice (CGS2, orthogonal to working precision), Q stored as
// K rows of length S, c = Q^T b maintained incrementally.  Instead of R the factorisation keeps
// T = R^{-1} (upper triangular, column-major; appending a column [r; rho] to R appends
// [-T r / rho; 1 / rho] to T), so the least-squares solve z = T c is a fully parallel
// triangular mat-vec rather than a K-step back substitution with a block barrier per step
// (measured: 1-1.5 us per step, ~200 us per OMP iteration at K ~ 150).
// Single thread block; the same Blk abstraction as step_logic.h, so the file also compiles for the
// host (nthr == 1) and tests/test_hostcheck_logic.py checks it against scipy.optimize.nnls.
#pragma once
#include <math.h>
#include <stdint.h>
#include "step_logic.h"

namespace bcg {

struct NnlsWork {
  double* Q;      // cap x S, row p = orthonormal vector q_p
  double* R;      // cap x cap column-major, T = R^{-1}: T[i + j*cap], i <= j
  double* c;      // cap   q_p . b
  double* z;      // 2 cap least-squares solution on P (upper half: Gram-Schmidt coefficients of the append)
  double* wP;     // cap   current feasible weights on P
  double* h;      // cap   scratch
  double* v;      // S     scratch column
  int32_t* P;     // cap   P position -> active-set slot
  int32_t* Z;     // cap   zero set (slots of the problem that are not in P)
  int32_t* inP;   // cap   slot -> 1 when the slot is in P
  double* R2;     // cap x cap second buffer of T (a column removal writes the new T there, then the two swap)
  double* rot;    // 3 cap  Givens rotations of a column removal: c_j, s_j, staged row of T
  int32_t* rem;   // cap   P positions dropped by the current step-back
  int32_t nrem, downdate;   // downdate = 1: remove columns by Givens rotations; 0: rebuild behind the first removed
  int32_t nP, nZ, cap, valid;
  int32_t tld;                     // leading dimension of T and row capacity of Q: min(cap, S + 1) -- P holds independent columns, nP <= S
  int32_t outer_iters, rebuilds;   // diagnostics of the last solve
  int32_t first_removed;           // first P position dropped by the last step-back (block-uniform hand-over)
};

BCG_HD double act_col(const SolverState* st, int slot, int s) {
  return st->act_norm[slot] * (double)st->act_rows[(size_t)slot * st->ld + s];
}

// append the column of `slot` to the QR of P; returns false when it is numerically dependent
BCG_HD bool nnls_qr_append(const Blk& B, SolverState* st, NnlsWork* W, int slot) {
  const int S = st->S, p = W->nP, cap = W->cap, tld = W->tld;
  double n0 = 0.;
  for (int s = B.tid; s < S; s += B.nthr) { const double x = act_col(st, slot, s); W->v[s] = x; n0 += x * x; }
  blk_sum<1>(B, &n0);                                   // (also a barrier: v is complete)
  double* const Q = W->Q;
  double* const T = W->R;
  double* const v = W->v;
  double* const h = W->h;
  double* const r = W->z + cap;                         // Gram-Schmidt coefficients r = h1 + h2
  for (int pass = 0; pass < 2; ++pass) {
    // h_i = q_i . v
    blk_dots<double>(B, p, S, [&](int i) { return (const double*)(Q + (size_t)i * S); }, v,
                     [&](int i, double d) { h[i] = d; r[i] = (pass == 0 ? 0. : r[i]) + d; });
    // v -= sum_i h_i q_i
    blk_combine<double>(B, S, p, [&](int i) { return CombTerm<double>{h[i], Q + (size_t)i * S, S}; },
                        [&](int s, double acc) { v[s] -= acc; });
    omp_mark(B, st, 12 + pass);
  }
  double u[2] = {0., 0.};
  for (int s = B.tid; s < S; s += B.nthr) { const double x = v[s]; u[0] += x * x; u[1] += x * st->b[s]; }
  blk_sum<2>(B, u);
  const double rpp = sqrt(u[0]);
  if (!(rpp > 1e-13 * sqrt(n0)) || p >= tld) return false;   // (numerically) in the span of P (P holds at most S columns)
  double* q = Q + (size_t)p * S;
  for (int s = B.tid; s < S; s += B.nthr) q[s] = v[s] / rpp;
  // new column of T = R^{-1}:  -(T r) / rho on top of 1 / rho   (column j of T holds rows 0..j)
  omp_mark(B, st, 14);
  blk_combine<double>(B, p, p, [&](int j) { return CombTerm<double>{r[j], T + (size_t)j * tld, j + 1}; },
                      [&](int i, double acc) { T[(size_t)i + (size_t)p * tld] = -acc / rpp; });
  if (B.tid == 0) {
    T[(size_t)p + (size_t)p * tld] = 1. / rpp;
    W->c[p] = u[1] / rpp;
    W->P[p] = slot;
    W->inP[slot] = 1;
    W->wP[p] = 0.;
    W->nP = p + 1;
  }
  B.sync();
  return true;
}

// z = R^{-1} c = T c : no sequential dependency (column j of T is contiguous over its rows 0..j)
BCG_HD void nnls_solve_R(const Blk& B, NnlsWork* W) {
  const int n = W->nP, tld = W->tld;
  const double* const T = W->R;
  const double* const c = W->c;
  double* const z = W->z;
  blk_combine<double>(B, n, n, [&](int j) { return CombTerm<double>{c[j], T + (size_t)j * tld, j + 1}; },
                      [&](int i, double acc) { z[i] = acc; });
}

// Remove position k (of p) from the factorisation; positions behind it move down by one.  With A_P = Q R and
// E = the identity without column k, A_P E = (Q G^T) (G R E) for any orthogonal G; G R E is upper triangular with a
// zero last row exactly when the last row of G is the normalised row k of T = R^{-1} (that row is orthogonal to
// every remaining column of R).  Such a G is a sweep of plane rotations (j, j+1), j = k .. p-2, that moves the mass
// of t = T[k, k:p] into its last component, so the rotations come from ONE row of T by a scalar recurrence; they are
// then applied to the q vectors (Q G^T), to the columns of T with row k deleted (T_new = (E^T T G^T)[:, :p-1]) and
// to c = Q^T b.  R itself is never needed.  P / wP are compacted by the caller.
BCG_HD void nnls_qr_remove(const Blk& B, SolverState* st, NnlsWork* W, int k, int p) {
  const int S = st->S, cap = W->cap, tld = W->tld;
  double* const T = W->R;
  double* const Tn = W->R2;
  double* const Q = W->Q;
  double* const c = W->c;
  double* const rc = W->rot;
  double* const rs = W->rot + cap;
  double* const trow = W->rot + 2 * (size_t)cap;
  for (int j = k + B.tid; j < p; j += B.nthr) trow[j] = T[(size_t)k + (size_t)j * tld];   // strided row: stage it
  B.sync();
  if (B.tid == 0) {
    double carry = trow[k], ccur = c[k];
    for (int j = k; j < p - 1; ++j) {
      const double a = carry, b = trow[j + 1];
      const double h = sqrt(a * a + b * b);                    // >= |T[k,k]| > 0
      const double ih = 1. / h;
      const double cj = b * ih, sj = a * ih;                   // (c a - s b, s a + c b) = (0, h)
      rc[j] = cj; rs[j] = sj; carry = h;
      const double y = c[j + 1];
      c[j] = cj * ccur - sj * y;                               // c[j]'s old value was consumed one step earlier
      ccur = sj * ccur + cj * y;
    }
  }
  B.sync();
  // q vectors: thread per component s, the rotation sweep runs down the rows k .. p-1 (in place)
  for (int s = B.tid; s < S; s += B.nthr) {
    double cur = Q[(size_t)k * S + s];
#pragma unroll 4
    for (int j = k; j < p - 1; ++j) {
      const double y = Q[(size_t)(j + 1) * S + s];
      Q[(size_t)j * S + s] = rc[j] * cur - rs[j] * y;
      cur = rs[j] * cur + rc[j] * y;
    }
  }
  // T: thread per new row i (old row r = i, or i + 1 behind the deleted row), sweep over the columns k .. p-1 into the
  // second buffer; entries below the diagonal are never stored, so they read as zero
  for (int i = B.tid; i < p - 1; i += B.nthr) {
    const int r = i + (i >= k ? 1 : 0);
    for (int j = i; j < k; ++j) Tn[(size_t)i + (size_t)j * tld] = T[(size_t)i + (size_t)j * tld];   // (only rows above k)
    const int j0 = (i >= k) ? i : k;                           // first column with a non-zero result
    double cur = (r <= j0) ? T[(size_t)r + (size_t)j0 * tld] : 0.;
#pragma unroll 4
    for (int j = j0; j < p - 1; ++j) {
      const double y = T[(size_t)r + (size_t)(j + 1) * tld];   // r <= j + 1 always holds here
      Tn[(size_t)i + (size_t)j * tld] = rc[j] * cur - rs[j] * y;
      cur = rs[j] * cur + rc[j] * y;
    }
  }
  B.sync();
  if (B.tid == 0) { W->R = Tn; W->R2 = T; }
  B.sync();
}

// Rebuild the factorisation for the slots currently listed in P (after removals), keeping their weights.
// Positions < from were not touched by the removal: their q vectors, T columns and c entries stay valid;
// only the columns behind the first removed position are re-appended.
BCG_HD void nnls_rebuild(const Blk& B, SolverState* st, NnlsWork* W, int from) {
  const int n = W->nP;
  B.sync();
  if (B.tid == 0) W->nP = from;
  B.sync();
  int kept = from;
  for (int i = from; i < n; ++i) {
    // the append only writes positions <= i of P / wP while re-inserting entry i: read in place first
    const int slot = W->P[i];
    const double w = W->wP[i];
    B.sync();
    if (nnls_qr_append(B, st, W, slot)) {
      if (B.tid == 0) W->wP[kept] = w;
      ++kept;
    } else {
      if (B.tid == 0) { W->inP[slot] = 0; }
    }
    B.sync();
  }
  if (B.tid == 0) W->rebuilds += 1;
  B.sync();
}

// Solve the NNLS over the slots with act_w > 0.  from_scratch: forget the factorisation (optimize()).
// On return act_w holds the solution (zeros for the columns left out), st->xw / st->err are refreshed.
// xw_current: st->xw equals A w for the weights the solver was left with by its previous call (true inside the
// OMP loop) -- then a valid warm start needs no residual pass before the first dual test.
BCG_HD void nnls_solve(const Blk& B, SolverState* st, NnlsWork* W, int from_scratch, int xw_current = 0) {
  const int S = st->S, nact = st->nact, ld = st->ld;
  // a passive column whose weight the host zeroed invalidates the warm start
  double bad = 0.;
  for (int p = B.tid; p < W->nP; p += B.nthr) bad += (st->act_w[W->P[p]] > 0.) ? 0. : 1.;
  blk_sum<1>(B, &bad);
  const bool warm = !(from_scratch || !W->valid || bad > 0.);
  B.sync();
  if (!warm) {
    for (int k = B.tid; k < W->cap; k += B.nthr) W->inP[k] = 0;
    B.sync();
    if (B.tid == 0) { W->nP = 0; }
    B.sync();
  }
  // zero set = problem columns (act_w > 0) not in P, in slot order
  if (B.tid == 0) { W->nZ = 0; W->outer_iters = 0; W->rebuilds = 0; }
  B.sync();
  for (int k = B.tid; k < nact; k += B.nthr) {
    if (st->act_w[k] > 0. && !W->inP[k]) {
#ifdef __CUDA_ARCH__
      const int pos = atomicAdd(&W->nZ, 1);
#else
      const int pos = W->nZ++;
#endif
      W->Z[pos] = k;                                          // order is irrelevant: ties go to the lowest slot
    }
  }
  B.sync();
  omp_mark(B, st, 5);
  const double tolscale = st->bnorm;                     // scale of the dual tolerance, ||b||
  const int maxit = 3 * (W->nP + W->nZ) + 10;
  double* const ax = st->xw_new;                              // A_P wP
  double* const resid = st->xf;                               // b - A_P wP (xf is otherwise unused by OMP)
  bool ax_is_final = false;                                   // ax == A_P wP for the P / wP the loop ends with
  for (int outer = 0; outer < maxit; ++outer) {
    if (outer == 0 && warm && xw_current) {
      for (int s = B.tid; s < S; s += B.nthr) { const double a = st->xw[s]; ax[s] = a; resid[s] = st->b[s] - a; }
      B.sync();
    } else {
      const int nP = W->nP;
      blk_combine<float>(B, S, nP,
                         [&](int p) {
                           const int slot = W->P[p];
                           return CombTerm<float>{W->wP[p] * st->act_norm[slot], st->act_rows + (size_t)slot * ld, S};
                         },
                         [&](int s, double acc) { ax[s] = acc; resid[s] = st->b[s] - acc; });
    }
    ax_is_final = true;
    if (W->nZ == 0) break;
    // duals of the zero set, scale-free (cosine-like): d_j = (a_j / ||a_j||) . (b - A w)
    blk_dots<float>(B, W->nZ, S, [&](int j) { return (const float*)(st->act_rows + (size_t)W->Z[j] * ld); }, resid,
                    [&](int j, double d) { W->h[j] = d; });
    double key = -INFINITY; int64_t id = -1; int pl = -1;
    for (int j = B.tid; j < W->nZ; j += B.nthr) {
      const double d = W->h[j];
      if (id < 0 || d > key || (d == key && W->Z[j] < id)) { key = d; id = W->Z[j]; pl = j; }
    }
    blk_argbest(B, &key, &id, &pl);
    if (outer == 0) omp_mark(B, st, 6);
    if (id < 0 || !(key > 1e-13 * tolscale)) break;           // KKT: no zero column has a positive dual
    ax_is_final = false;
    // move it into P
    if (B.tid == 0) { W->Z[pl] = W->Z[W->nZ - 1]; W->nZ -= 1; W->outer_iters += 1; }
    B.sync();
    if (!nnls_qr_append(B, st, W, (int)id)) continue;         // dependent column: stays at zero, dropped
    if (outer == 0) omp_mark(B, st, 7);
    // inner loop: move towards the unconstrained solution on P, dropping columns that hit zero
    for (int inner = 0; inner < maxit; ++inner) {
      nnls_solve_R(B, W);
      double mn = INFINITY;
      for (int p = B.tid; p < W->nP; p += B.nthr) mn = fmin(mn, W->z[p]);
      double neg = -mn;                                       // block max of -z = -(block min of z)
      { int64_t i2 = 0; int p2 = 0; double k2 = neg; blk_argbest(B, &k2, &i2, &p2); neg = k2; }
      if (-neg > 0.) {
        for (int p = B.tid; p < W->nP; p += B.nthr) W->wP[p] = W->z[p];
        B.sync();
        if (outer == 0 && inner == 0) omp_mark(B, st, 8);
        break;
      }
      // step length alpha = min_{z_p <= 0} w_p / (w_p - z_p)
      double a = -INFINITY;
      for (int p = B.tid; p < W->nP; p += B.nthr)
        if (W->z[p] <= 0.) a = fmax(a, -(W->wP[p] / (W->wP[p] - W->z[p])));
      { int64_t i2 = 0; int p2 = 0; double k2 = a; blk_argbest(B, &k2, &i2, &p2); a = -k2; }
      if (!(a >= 0.) || !(a <= 1.)) a = 0.;
      double wmax = 0.;
      for (int p = B.tid; p < W->nP; p += B.nthr) {
        const double w = W->wP[p] + a * (W->z[p] - W->wP[p]);
        W->wP[p] = w;
        wmax = fmax(wmax, w);
      }
      { int64_t i2 = 0; int p2 = 0; double k2 = wmax; blk_argbest(B, &k2, &i2, &p2); wmax = k2; }
      const int p_before = W->nP;          // read by every thread BEFORE the barrier: thread 0 rewrites W->nP right behind it
      B.sync();
      // drop the columns that reached zero (to the zero set) and compact P
      if (B.tid == 0) {
        int keep = 0, first = -1, nrem = 0;
        for (int p = 0; p < W->nP; ++p) {
          if (W->wP[p] > 1e-15 * wmax) { W->P[keep] = W->P[p]; W->wP[keep] = W->wP[p]; ++keep; }
          else { if (first < 0) first = p; W->rem[nrem++] = p; W->inP[W->P[p]] = 0; W->Z[W->nZ++] = W->P[p]; }
        }
        W->nP = keep;
        W->nrem = nrem;
        W->first_removed = (first < 0) ? keep : first;
      }
      B.sync();
      if (W->downdate) {
        // Givens removal, highest position first (the positions below it keep their meaning)
        const int nrem = W->nrem;
        for (int q = nrem - 1; q >= 0; --q) nnls_qr_remove(B, st, W, W->rem[q], p_before - (nrem - 1 - q));
        if (B.tid == 0 && nrem > 0) W->rebuilds += 1;
        B.sync();
      } else {
        const int first_removed = W->first_removed;
        nnls_rebuild(B, st, W, first_removed < W->nP ? first_removed : W->nP);
      }
      if (W->nP == 0) break;
    }
  }
  // write the solution back
  omp_mark(B, st, 9);
  for (int k = B.tid; k < nact; k += B.nthr)
    if (st->act_w[k] > 0.) st->act_w[k] = 0.;
  B.sync();
  for (int p = B.tid; p < W->nP; p += B.nthr) st->act_w[W->P[p]] = W->wP[p];
  if (B.tid == 0) { W->valid = 1; st->kkt_valid = 1; }
  B.sync();
  if (ax_is_final) {
    // the last residual pass already is A w for the final weights: error() without another K x S pass
    double e = 0.;
    for (int s = B.tid; s < S; s += B.nthr) { st->xw[s] = ax[s]; const double rr = resid[s]; e += rr * rr; }
    blk_sum<1>(B, &e);
    if (B.tid == 0) st->err = sqrt(e);
    B.sync();
  } else {
    refresh_iterate(B, st);
  }
}

// One OMP iteration behind the selection scan (orthopursuit.py:17-42 inside snnls.py:41-78): selection (+ w[f] = 1),
// NNLS re-solve on the active set, monotone-error check with revert, event log, retry / latch, and -- when another
// iteration follows -- the residual direction for its scan (orthopursuit.py:18).
// act_w_new keeps the weights from before the iteration for the revert.
BCG_HD void omp_iteration(const Blk& B, SolverState* st, NnlsWork* W, int prep_next) {
  if (st->halted) return;
  omp_mark(B, st, 0);
  bool nonempty;
  count_positive(B, st, &nonempty);
  const double prev_err = st->err;
  const int nact0 = st->nact;
  for (int k = B.tid; k < nact0; k += B.nthr) st->act_w_new[k] = st->act_w[k];
  B.sync();
  omp_mark(B, st, 1);
  const int64_t f = omp_select(B, st);
  if (st->comm_error) return;
  omp_mark(B, st, 4);
  nnls_solve(B, st, W, 0, 1);
  omp_mark(B, st, 10);
  const double err = st->err;
  if (st->check_monotone && nonempty && err > prev_err) {  // snnls.py:56-61: revert
    for (int k = B.tid; k < st->nact; k += B.nthr) st->act_w[k] = (k < nact0) ? st->act_w_new[k] : 0.;
    B.sync();
    if (B.tid == 0) { W->valid = 0; st->kkt_valid = 0; }
    B.sync();
    refresh_iterate(B, st);
    if (B.tid == 0) fail_event(st, BCG_IT_FAIL_MONOTONE, f, err, prev_err);
  } else if (B.tid == 0) {
    if (st->check_monotone && nonempty) st->retried = 0;
    push_event(st, BCG_IT_OK, f, st->nact, err, 0., 0.);
  }
  B.sync();
  if (prep_next && !st->halted) prepare_select(B, st);
#ifdef __CUDA_ARCH__
  if (st->omp_trace && B.tid == 0) {                     // n_events already counts this iteration
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    st->omp_trace[(size_t)(st->n_events - 1) * 16 + 11] = t;
  }
#endif
}

}  // namespace bcg
