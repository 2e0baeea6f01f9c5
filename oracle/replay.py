"""Oracle (TEST INFRASTRUCTURE): the reference greedy solvers restated in SPARSE form for sizes at which the dense
float64 N x S matrix does not fit the test budget (N = 1e7, S = 512 needs 2 x 41 GB in the reference).

The arithmetic of every O(S) / O(K S) expression is the reference's (same formulas as oracle/greedy.py, which is
pinned bit-for-bit to the reference); the only thing that changes is WHERE the two O(N S) quantities come from:

  * the per-row scores of an iteration (giga.py:31-38, frankwolfe.py:17, orthopursuit.py:19) are supplied by a
    callback ``score_fn(kind, dirs) -> (N,) float64`` -- on the GPU box the independent float64 audit kernel
    (csrc/audit_kernel.cuh, which recomputes every row from the raw data), in CPU tests dense NumPy;
  * ``A.dot(w)`` is formed from the K active rows only (w is zero elsewhere), rows supplied by
    ``rows_fn(idx) -> (len(idx), S)`` (the oracle's float64 model evaluation of those datapoints).

tests/test_replay_oracle.py checks these classes against the dense oracle (identical selections, weights to 1e-12).
"""
import numpy as np
from scipy.optimize import nnls as _scipy_nnls

from .greedy import (IT_OK, IT_FAIL_CDIR, IT_FAIL_GEODESIC, IT_FAIL_GAMMA, IT_FAIL_MONOTONE, DEFAULT_TOL,
                     OracleNumericalPrecisionError)


class SparseReplay(object):
  """snnls/snnls.py:8-79 with w stored sparsely (selection order)."""
  kind = None

  def __init__(self, n, b, score_fn, rows_fn, tol=DEFAULT_TOL, check_error_monotone=True):
    self.n = int(n)
    self.b = np.asarray(b, dtype=np.float64)
    self.score_fn, self.rows_fn = score_fn, rows_fn
    self.tol = tol
    self.check_error_monotone = check_error_monotone
    self.idx = np.zeros(0, dtype=np.int64)          # stored rows, selection order
    self.wa = np.zeros(0)                           # their weights (may contain zeros)
    self.rows = np.zeros((0, self.b.shape[0]))      # their float64 vectors
    self.reached_numeric_limit = False
    self.events = []
    self.diag = []                                  # per selection: (top score, top-2 gap, rows within the fp32 window)

  # ---- sparse state ----------------------------------------------------------------------
  def size(self):
    return int((self.wa > 0).sum())

  @property
  def w(self):
    w = np.zeros(self.n)
    w[self.idx] = self.wa
    return w

  def xw(self):
    return self.wa.dot(self.rows) if self.idx.shape[0] else np.zeros(self.b.shape[0])

  def error(self):
    return np.sqrt(((self.xw() - self.b)**2).sum())               # snnls.py:28-29

  def _slot(self, f):
    hit = np.flatnonzero(self.idx == f)
    if hit.shape[0]:
      return int(hit[0])
    self.idx = np.append(self.idx, np.int64(f))
    self.wa = np.append(self.wa, 0.)
    self.rows = np.vstack((self.rows, self.rows_fn(np.array([f], dtype=np.int64))))
    return self.idx.shape[0] - 1

  def _record(self, sc, scale=1.):
    """scale: norm of the direction (the engine scans the UNIT residual, so its window is relative to that)"""
    sc = sc/scale if scale != 1. else sc
    top2 = np.partition(sc, -2)[-2:] if sc.shape[0] > 1 else np.array([-np.inf, sc[0]])
    top, second = float(top2[1]), float(top2[0])
    win = 2e-5 + 1e-5*abs(top)
    self.diag.append((top, top - second, int((sc >= top - win).sum())))

  # ---- snnls.py:31-79 -----------------------------------------------------------------------
  def build(self, itrs):
    first_event = len(self.events)
    if self.reached_numeric_limit or self.n == 0:
      return []
    retried = False
    for _ in range(itrs):
      f = -1
      saved = (self.idx.copy(), self.wa.copy(), self.rows)
      try:
        nonempty = self.size() > 0
        if self.check_error_monotone and nonempty:
          err_before = self.error()
        f = int(self.select())
        self.reweight(f)
        if self.check_error_monotone and nonempty:
          err_after = self.error()
          if err_after > err_before:
            raise OracleNumericalPrecisionError(IT_FAIL_MONOTONE)
          retried = False
        self.events.append((IT_OK, f, float(self.error())))
      except OracleNumericalPrecisionError as e:
        self.idx, self.wa = saved[0], saved[1]
        self.rows = self.rows[:self.idx.shape[0]]
        self.events.append((e.code, f, float(self.error())))
        if retried:
          self.reached_numeric_limit = True
          break
        retried = True
    return self.events[first_event:]


class GigaReplay(SparseReplay):
  """snnls/giga.py:6-64"""
  def __init__(self, n, b, score_fn, rows_fn, **kw):
    super().__init__(n, b, score_fn, rows_fn, **kw)
    self.bnorm = np.sqrt((self.b**2).sum())
    self.bn = self.b/self.bnorm

  def _unit_iterate(self):
    xw = self.xw()
    nw = np.sqrt((xw**2).sum())
    return xw, (1. if nw == 0. else nw)

  def select(self):
    xw, nw = self._unit_iterate()
    xw = xw/nw
    cdir = self.bn - self.bn.dot(xw)*xw
    cdirnrm = np.sqrt((cdir**2).sum())
    if cdirnrm < self.tol:
      raise OracleNumericalPrecisionError(IT_FAIL_CDIR)
    cdir /= cdirnrm
    sc = self.score_fn('giga', np.vstack((cdir, xw)))
    self._record(sc)
    return sc.argmax()

  def reweight(self, f):
    xw, nw = self._unit_iterate()
    k = self._slot(f)
    xf = self.rows[k]
    nf = np.sqrt((xf**2).sum())
    gA = self.bn.dot((xf/nf)) - self.bn.dot((xw/nw))*(xw/nw).dot((xf/nf))
    gB = self.bn.dot((xw/nw)) - self.bn.dot((xf/nf))*(xw/nw).dot((xf/nf))
    if gA <= 0. or gB < 0:
      raise OracleNumericalPrecisionError(IT_FAIL_GEODESIC)
    a = gB/(gA+gB)/nw
    b = gA/(gA+gB)/nf
    x = a*xw + b*xf
    nx = np.sqrt((x**2).sum())
    scale = self.bnorm/nx*(x/nx).dot(self.bn)
    alpha, beta = a*scale, b*scale
    self.wa = alpha*self.wa
    self.wa[k] = max(0., self.wa[k] + beta)


class FrankWolfeReplay(SparseReplay):
  """snnls/frankwolfe.py:5-40; norm_sum = Anorms.sum() over ALL rows (supplied by the caller)"""
  def __init__(self, n, b, score_fn, rows_fn, norm_sum, **kw):
    super().__init__(n, b, score_fn, rows_fn, **kw)
    self.nsum = float(norm_sum)

  def select(self):
    resid = self.b - self.xw()
    sc = self.score_fn('lin', resid[np.newaxis, :])
    self._record(sc, np.sqrt((resid**2).sum()))
    return sc.argmax()

  def reweight(self, f):
    empty = self.size() == 0
    xw = self.xw()
    k = self._slot(f)
    xf = self.rows[k]
    nf = np.sqrt((xf**2).sum())
    if empty:
      alpha, beta = 0., self.nsum/nf
    else:
      gammanum = (self.nsum/nf*xf - xw).dot(self.b - xw)
      gammadenom = ((self.nsum/nf*xf - xw)**2).sum()
      if gammanum < 0. or gammadenom == 0. or gammanum > gammadenom:
        raise OracleNumericalPrecisionError(IT_FAIL_GAMMA)
      alpha = 1. - gammanum/gammadenom
      beta = self.nsum/nf*gammanum/gammadenom
    self.wa = alpha*self.wa
    self.wa[k] = max(0., self.wa[k] + beta)


class OrthoPursuitReplay(SparseReplay):
  """snnls/orthopursuit.py:7-42"""
  def select(self):
    resid = self.b - self.xw()
    dots = self.score_fn('lin', resid[np.newaxis, :])
    self._record(dots, np.sqrt((resid**2).sum()))
    fpos = dots.argmax()
    if self.size() == 0:
      return fpos
    pos = dots[fpos]
    act = np.flatnonzero(self.wa > 0)
    act = act[np.argsort(self.idx[act], kind='stable')]          # ascending index, as w > 0 over the dense vector
    neg_all = -dots[self.idx[act]]
    fneg = neg_all.argmax()
    if pos >= neg_all[fneg]:
      return fpos
    return self.idx[act][fneg]

  def reweight(self, f):
    k = self._slot(f)
    self.wa[k] = 1.
    act = np.flatnonzero(self.wa > 0)
    act = act[np.argsort(self.idx[act], kind='stable')]
    sol = _scipy_nnls(np.ascontiguousarray(self.rows[act].T), self.b, maxiter=100*self.n)
    self.wa[act] = sol[0]


REPLAYS = {'giga': GigaReplay, 'fw': FrankWolfeReplay, 'omp': OrthoPursuitReplay}


def dense_score_fn(vecs):
  """score callback over a dense float64 (N, S) matrix (CPU tests; same expressions as oracle/greedy.py)"""
  norms = np.sqrt((vecs**2).sum(axis=1))
  An = vecs/norms[:, np.newaxis]

  def fn(kind, dirs):
    if kind == 'giga':
      sc = An.dot(dirs.T)
      ok = np.logical_and(sc[:, 1] > -1.+1e-14, 1.-sc[:, 1]**2 > 0.)
      sc[ok, 1] = np.sqrt(1.-sc[ok, 1]**2)
      sc[np.logical_not(ok), 1] = np.inf
      return sc[:, 0]/sc[:, 1]
    return An.dot(dirs[0])
  return fn
