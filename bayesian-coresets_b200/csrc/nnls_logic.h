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
// When a weight would turn negative the standard step-back / removal loop runs and the QR of the
// reduced set is rebuilt.
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
  int32_t nP, nZ, cap, valid;
  int32_t outer_iters, rebuilds;   // diagnostics of the last solve
  int32_t first_removed;           // first P position dropped by the last step-back (block-uniform hand-over)
};

BCG_HD int blk_lane(const Blk& B) {
#ifdef __CUDA_ARCH__
  return B.tid & 31;
#else
  (void)B; return 0;
#endif
}
BCG_HD int blk_warp(const Blk& B) {
#ifdef __CUDA_ARCH__
  return B.tid >> 5;
#else
  (void)B; return 0;
#endif
}
BCG_HD int blk_nwarps(const Blk& B) {
#ifdef __CUDA_ARCH__
  return B.nthr >> 5;
#else
  (void)B; return 1;
#endif
}
BCG_HD int blk_lanes(const Blk& B) {
#ifdef __CUDA_ARCH__
  (void)B; return 32;
#else
  (void)B; return 1;
#endif
}
BCG_HD double blk_warp_sum(const Blk& B, double v) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
#endif
  (void)B;
  return v;
}

BCG_HD double act_col(const SolverState* st, int slot, int s) {
  return st->act_norm[slot] * (double)st->act_rows[(size_t)slot * st->ld + s];
}

// append the column of `slot` to the QR of P; returns false when it is numerically dependent
BCG_HD bool nnls_qr_append(const Blk& B, SolverState* st, NnlsWork* W, int slot) {
  const int S = st->S, p = W->nP, cap = W->cap;
  const int lane = blk_lane(B), warp = blk_warp(B), nw = blk_nwarps(B), lanes = blk_lanes(B);
  double n0 = 0.;
  for (int s = B.tid; s < S; s += B.nthr) { const double x = act_col(st, slot, s); W->v[s] = x; n0 += x * x; }
  blk_sum<1>(B, &n0);                                   // (also a barrier: v is complete)
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = warp; i < p; i += nw) {                // h_i = q_i . v, one warp per i
      double part = 0.;
      const double* q = W->Q + (size_t)i * S;
      for (int s = lane; s < S; s += lanes) part += q[s] * W->v[s];
      part = blk_warp_sum(B, part);
      if (lane == 0) W->h[i] = part;
    }
    B.sync();
    for (int s = B.tid; s < S; s += B.nthr) {           // v -= sum_i h_i q_i
      double acc = W->v[s];
      for (int i = 0; i < p; ++i) acc -= W->h[i] * W->Q[(size_t)i * S + s];
      W->v[s] = acc;
    }
    for (int i = B.tid; i < p; i += B.nthr) W->z[cap + i] = (pass == 0 ? 0. : W->z[cap + i]) + W->h[i];   // r = h1 + h2
    B.sync();
  }
  double u[2] = {0., 0.};
  for (int s = B.tid; s < S; s += B.nthr) { const double x = W->v[s]; u[0] += x * x; u[1] += x * st->b[s]; }
  blk_sum<2>(B, u);
  const double rpp = sqrt(u[0]);
  if (!(rpp > 1e-13 * sqrt(n0))) return false;          // (numerically) in the span of P
  double* q = W->Q + (size_t)p * S;
  for (int s = B.tid; s < S; s += B.nthr) q[s] = W->v[s] / rpp;
  // new column of T = R^{-1}:  -(T r) / rho on top of 1 / rho
  for (int i = B.tid; i < p; i += B.nthr) {
    double acc = 0.;
    for (int j = i; j < p; ++j) acc += W->R[(size_t)i + (size_t)j * cap] * W->z[cap + j];
    W->h[i] = -acc / rpp;
  }
  B.sync();
  for (int i = B.tid; i < p; i += B.nthr) W->R[(size_t)i + (size_t)p * cap] = W->h[i];
  if (B.tid == 0) {
    W->R[(size_t)p + (size_t)p * cap] = 1. / rpp;
    W->c[p] = u[1] / rpp;
    W->P[p] = slot;
    W->inP[slot] = 1;
    W->wP[p] = 0.;
    W->nP = p + 1;
  }
  B.sync();
  return true;
}

// z = R^{-1} c = T c : thread i owns row i (column-major T: coalesced over i), no sequential dependency
BCG_HD void nnls_solve_R(const Blk& B, NnlsWork* W) {
  const int n = W->nP, cap = W->cap;
  for (int i = B.tid; i < n; i += B.nthr) {
    double acc = 0.;
    for (int j = i; j < n; ++j) acc += W->R[(size_t)i + (size_t)j * cap] * W->c[j];
    W->z[i] = acc;
  }
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
BCG_HD void nnls_solve(const Blk& B, SolverState* st, NnlsWork* W, int from_scratch) {
  const int S = st->S, nact = st->nact;
  const int lane = blk_lane(B), warp = blk_warp(B), nw = blk_nwarps(B), lanes = blk_lanes(B);
  // a passive column whose weight the host zeroed invalidates the warm start
  double bad = 0.;
  for (int p = B.tid; p < W->nP; p += B.nthr) bad += (st->act_w[W->P[p]] > 0.) ? 0. : 1.;
  blk_sum<1>(B, &bad);
  if (from_scratch || !W->valid || bad > 0.) {
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
  // scale of the dual tolerance
  double tolscale = 0.;
  {
    double m = 0.;
    for (int s = B.tid; s < S; s += B.nthr) m += st->b[s] * st->b[s];
    blk_sum<1>(B, &m);
    tolscale = sqrt(m);
  }
  const int maxit = 3 * (W->nP + W->nZ) + 10;
  for (int outer = 0; outer < maxit; ++outer) {
    // residual b - A_P wP  -> st->xw_new holds A_P wP
    for (int s = B.tid; s < S; s += B.nthr) {
      double acc = 0.;
      for (int p = 0; p < W->nP; ++p) acc += W->wP[p] * act_col(st, W->P[p], s);
      st->xw_new[s] = acc;
    }
    B.sync();
    if (W->nZ == 0) break;
    // duals of the zero set: d_j = a_j . (b - A w), one warp per column
    for (int j = warp; j < W->nZ; j += nw) {
      const int slot = W->Z[j];
      double part = 0.;
      for (int s = lane; s < S; s += lanes) part += act_col(st, slot, s) * (st->b[s] - st->xw_new[s]);
      part = blk_warp_sum(B, part);
      if (lane == 0) W->h[j] = part / st->act_norm[slot];     // scale-free (cosine-like) dual
    }
    B.sync();
    double key = -INFINITY; int64_t id = -1; int pl = -1;
    for (int j = B.tid; j < W->nZ; j += B.nthr) {
      const double d = W->h[j];
      if (id < 0 || d > key || (d == key && W->Z[j] < id)) { key = d; id = W->Z[j]; pl = j; }
    }
    blk_argbest(B, &key, &id, &pl);
    if (id < 0 || !(key > 1e-13 * tolscale)) break;           // KKT: no zero column has a positive dual
    // move it into P
    if (B.tid == 0) { W->Z[pl] = W->Z[W->nZ - 1]; W->nZ -= 1; W->outer_iters += 1; }
    B.sync();
    if (!nnls_qr_append(B, st, W, (int)id)) continue;         // dependent column: stays at zero, dropped
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
      B.sync();
      // drop the columns that reached zero (to the zero set), compact P, rebuild the factorisation
      if (B.tid == 0) {
        int keep = 0, first = -1;
        for (int p = 0; p < W->nP; ++p) {
          if (W->wP[p] > 1e-15 * wmax) { W->P[keep] = W->P[p]; W->wP[keep] = W->wP[p]; ++keep; }
          else { if (first < 0) first = p; W->inP[W->P[p]] = 0; W->Z[W->nZ++] = W->P[p]; }
        }
        W->nP = keep;
        W->first_removed = (first < 0) ? keep : first;
      }
      B.sync();
      const int first_removed = W->first_removed;
      nnls_rebuild(B, st, W, first_removed < W->nP ? first_removed : W->nP);
      if (W->nP == 0) break;
    }
  }
  // write the solution back
  for (int k = B.tid; k < nact; k += B.nthr)
    if (st->act_w[k] > 0.) st->act_w[k] = 0.;
  B.sync();
  for (int p = B.tid; p < W->nP; p += B.nthr) st->act_w[W->P[p]] = W->wP[p];
  if (B.tid == 0) W->valid = 1;
  B.sync();
  refresh_iterate(B, st);
}

}  // namespace bcg
