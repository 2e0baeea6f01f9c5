import numpy as np
from scipy.optimize import nnls

from .base import SparseNNLS
from .. import util
from .. import _native as nat


class OrthoPursuit(SparseNNLS):
  """Orthogonal matching pursuit (reference: snnls/orthopursuit.py:7-42).

  Selection (the N x S residual-correlation scan, plus the negative direction over the active
  set) runs on the device; the reweight is the reference's own `scipy.optimize.nnls` call on the
  K active columns (K x S, gathered from the device's replicated active set)."""
  _alg = nat.ALG_OMP

  def _run(self, itrs):
    events = []
    retried = False
    for _ in range(itrs):
      nonempty = self.size() > 0
      prev_error = self.error()
      _, prev_w = self._native.active()
      f = self._native.omp_select()                       # orthopursuit.py:17-38 (w[f] = 1 on device)
      idx, w, pos, Aact = self._active_problem()
      res = nnls(Aact, self.b, maxiter=100*self.n_global)  # orthopursuit.py:40
      w_new = w.copy()
      w_new[pos] = res[0]
      self._native.set_weights(w_new)
      err = self.error()
      ev = nat.IterEvent(nat.IT_OK, idx.shape[0], f, err, 0., 0.)
      if nonempty and err > prev_error:                   # snnls.py:58-61
        revert = np.zeros(idx.shape[0])
        revert[:prev_w.shape[0]] = prev_w
        self._native.set_weights(revert)
        ev = nat.IterEvent(nat.IT_FAIL_MONOTONE, idx.shape[0], f, self.error(), err, prev_error)
      elif nonempty:
        retried = False
      events.append(ev)
      if ev.code != nat.IT_OK:
        if retried:
          self.reached_numeric_limit = True
          break
        retried = True
    return events
