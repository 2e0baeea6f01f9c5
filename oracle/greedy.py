"""Oracle (TEST INFRASTRUCTURE): float64 NumPy restatement of the reference greedy sparse-NNLS
solvers.  Each routine cites the reference lines it follows (paths relative to the reference
repository root).  The arithmetic is written with the same NumPy expressions, in the same
order, as the reference, so results are bit-identical to the reference in float64; the control
flow is restated as an explicit state machine that also records a per-iteration event log (the
device engine emits the same log, which is what the parity tests compare).

Layout convention (reference coreset/hilbert.py:24): ``A`` has shape (S, N) -- column n is the
S-dimensional projected vector of datapoint n -- and ``b`` has shape (S,).
"""
import numpy as np
from scipy.optimize import nnls as _scipy_nnls

# event codes written to the iteration log (shared with include/bcg.h: BCG_IT_*)
IT_OK = 0
IT_FAIL_CDIR = 1        # giga.py:28-29   cdirnrm < TOL
IT_FAIL_GEODESIC = 2    # giga.py:50-51   gA <= 0 or gB < 0
IT_FAIL_GAMMA = 3       # frankwolfe.py:33-34
IT_FAIL_MONOTONE = 4    # snnls.py:58-61
DEFAULT_TOL = 1e-12     # util/__init__.py:4


class OracleNumericalPrecisionError(Exception):
  """util/errors.py:1"""
  def __init__(self, code, msg=''):
    super().__init__(msg)
    self.code = code


class GreedyOracle(object):
  """snnls/snnls.py:8-106 -- state A (S,N), b (S,), dense w (N,); greedy build loop."""
  name = 'base'

  def __init__(self, A, b, tol=DEFAULT_TOL, check_error_monotone=True):
    self.A = A
    self.b = b
    self.tol = tol
    self.check_error_monotone = check_error_monotone
    self.w = np.zeros(A.shape[1])                       # snnls.py:15
    self.reached_numeric_limit = False                  # snnls.py:14
    self.events = []                                    # [(code, f or -1, error-after)]

  # snnls.py:18-29
  def reset(self):
    self.w = np.zeros(self.A.shape[1])
    self.reached_numeric_limit = False

  def size(self):
    return int((self.w > 0).sum())

  def weights(self):
    return self.w.copy()

  def error(self):
    return np.sqrt(((self.A.dot(self.w) - self.b)**2).sum())

  def build(self, itrs):
    """snnls.py:31-79.  Returns the list of events appended during this call."""
    first_event = len(self.events)
    if self.reached_numeric_limit or self.A.size == 0:  # snnls.py:32-38
      return []
    retried = False                                     # local to the call, snnls.py:40
    for _ in range(itrs):
      f = -1
      try:
        nonempty = self.size() > 0                      # sampled before the step, snnls.py:44
        if self.check_error_monotone and nonempty:
          err_before = self.error()
          w_before = self.w.copy()
        f = int(self.select())                          # snnls.py:50
        self.reweight(f)                                # snnls.py:53
        if self.check_error_monotone and nonempty:
          err_after = self.error()
          if err_after > err_before:                    # snnls.py:58-61
            self.w = w_before
            raise OracleNumericalPrecisionError(IT_FAIL_MONOTONE)
          retried = False                               # snnls.py:62
        self.events.append((IT_OK, f, float(self.error())))
      except OracleNumericalPrecisionError as e:        # snnls.py:63-72
        self.events.append((e.code, f, float(self.error())))
        if retried:
          self.reached_numeric_limit = True
          break
        retried = True
    return self.events[first_event:]

  def optimize(self):
    """snnls.py:82-97 -- NNLS re-solve restricted to the active set."""
    err_before = self.error()
    w_before = self.w.copy()
    active = self.w > 0
    sol = _scipy_nnls(self.A[:, active], self.b, maxiter=100*self.A.shape[1])
    self.w[active] = sol[0]
    if self.error() > err_before*(1. + self.tol):
      self.w = w_before
      self.reached_numeric_limit = True

  def select(self):
    raise NotImplementedError

  def reweight(self, f):
    raise NotImplementedError

  def _column_norms(self):
    norms = np.sqrt((self.A**2).sum(axis=0))
    if np.any(norms == 0):
      raise ValueError('A must not have any 0 columns')
    return norms


class GigaOracle(GreedyOracle):
  """snnls/giga.py:6-64"""
  name = 'giga'

  def __init__(self, A, b, **kw):
    super().__init__(A, b, **kw)
    norms = self._column_norms()                        # giga.py:10-12
    self.An = self.A / norms                            # giga.py:13
    self.bnorm = np.sqrt(((self.b)**2).sum())           # giga.py:15
    if self.bnorm == 0.:
      raise OracleNumericalPrecisionError(-1, 'norm of b must be > 0')
    self.bn = self.b / self.bnorm

  def _unit_iterate(self):
    xw = self.A.dot(self.w)
    nw = np.sqrt(((xw)**2).sum())
    nw = 1. if nw == 0. else nw                         # giga.py:22-23, 43-44
    return xw, nw

  def scores(self):
    """giga.py:20-38 up to (not including) the argmax; exposed so tests can measure top-2 gaps."""
    xw, nw = self._unit_iterate()
    xw /= nw
    cdir = self.bn - self.bn.dot(xw)*xw
    cdirnrm = np.sqrt((cdir**2).sum())
    if cdirnrm < self.tol:
      raise OracleNumericalPrecisionError(IT_FAIL_CDIR, 'cdirnrm < TOL: cdirnrm = ' + str(cdirnrm))
    cdir /= cdirnrm
    sc = self.An.T.dot(np.hstack((cdir[:, np.newaxis], xw[:, np.newaxis])))
    ok = np.logical_and(sc[:, 1] > -1.+1e-14, 1.-sc[:, 1]**2 > 0.)
    sc[ok, 1] = np.sqrt(1.-sc[ok, 1]**2)
    sc[np.logical_not(ok), 1] = np.inf
    return sc[:, 0]/sc[:, 1]

  def select(self):
    return self.scores().argmax()

  def reweight(self, f):
    """giga.py:40-64 -- closed-form geodesic line search."""
    xw, nw = self._unit_iterate()
    xf = self.A[:, f]
    nf = np.sqrt((xf**2).sum())
    gA = self.bn.dot((xf/nf)) - self.bn.dot((xw/nw)) * (xw/nw).dot((xf/nf))
    gB = self.bn.dot((xw/nw)) - self.bn.dot((xf/nf)) * (xw/nw).dot((xf/nf))
    if gA <= 0. or gB < 0:
      raise OracleNumericalPrecisionError(IT_FAIL_GEODESIC)
    a = gB/(gA+gB)/nw
    b = gA/(gA+gB)/nf
    x = a*xw + b*xf
    nx = np.sqrt((x**2).sum())
    scale = self.bnorm/nx*(x/nx).dot(self.bn)
    alpha = a*scale
    beta = b*scale
    self.w = alpha*self.w
    self.w[f] = max(0., self.w[f]+beta)


class FrankWolfeOracle(GreedyOracle):
  """snnls/frankwolfe.py:5-40"""
  name = 'fw'

  def __init__(self, A, b, **kw):
    super().__init__(A, b, **kw)
    self.Anorms = self._column_norms()                  # frankwolfe.py:10-12
    self.An = self.A / self.Anorms

  def scores(self):
    residual = self.b - self.A.dot(self.w)              # frankwolfe.py:16
    return self.An.T.dot(residual)

  def select(self):
    return self.scores().argmax()                       # frankwolfe.py:17

  def reweight(self, f):
    if self.size() == 0:                                # frankwolfe.py:20-23
      alpha = 0.
      beta = self.Anorms.sum() / self.Anorms[f]
    else:
      nsum = self.Anorms.sum()
      nf = self.Anorms[f]
      xw = self.A.dot(self.w)
      xf = self.A[:, f]
      gammanum = (nsum/nf*xf - xw).dot(self.b-xw)
      gammadenom = ((nsum/nf*xf-xw)**2).sum()
      if gammanum < 0. or gammadenom == 0. or gammanum > gammadenom:
        raise OracleNumericalPrecisionError(IT_FAIL_GAMMA)
      alpha = 1. - gammanum/gammadenom
      beta = nsum/nf*gammanum/gammadenom
    self.w = alpha*self.w
    self.w[f] = max(0., self.w[f]+beta)


class OrthoPursuitOracle(GreedyOracle):
  """snnls/orthopursuit.py:7-42"""
  name = 'omp'

  def __init__(self, A, b, **kw):
    super().__init__(A, b, **kw)
    self.An = self.A / self._column_norms()             # orthopursuit.py:12-15

  def scores(self):
    residual = self.b - self.A.dot(self.w)
    return self.An.T.dot(residual)

  def select(self):
    dots = self.scores()
    if self.size() == 0:                                # orthopursuit.py:22-23
      return dots.argmax()
    fpos = dots.argmax()                                # orthopursuit.py:26-35
    pos = dots[fpos]
    active = self.w > 0
    fneg = (-dots[active]).argmax()
    neg = (-dots[active])[fneg]
    if pos >= neg:
      return fpos
    return np.arange(self.w.shape[0])[active][fneg]

  def reweight(self, f):
    self.w[f] = 1.                                      # orthopursuit.py:38
    active = self.w > 0
    sol = _scipy_nnls(self.A[:, active], self.b, maxiter=100*self.A.shape[1])
    self.w[active] = sol[0]


ORACLES = {'giga': GigaOracle, 'fw': FrankWolfeOracle, 'omp': OrthoPursuitOracle}
