import os
import numpy as np
from scipy.optimize import nnls

from .base import SparseNNLS
from .. import util
from .. import _native as nat


class OrthoPursuit(SparseNNLS):
  """Orthogonal matching pursuit (reference: snnls/orthopursuit.py:7-42).

  Everything runs on the device, `build(itrs)` without a host round trip per iteration: the
  N x S residual-correlation scan, the negative direction over the active set, and the NNLS
  re-solve on the K active columns (float64 Lawson-Hanson, warm-started from the previous passive
  set -- csrc/nnls_logic.h; the NNLS minimiser is unique, so it equals `scipy.optimize.nnls`).
  With BCG_NNLS=scipy in the environment the reweight is the reference's own SciPy call on the
  host instead (kept for cross-checking)."""
  _alg = nat.ALG_OMP

  def _run(self, itrs):
    if os.environ.get('BCG_NNLS', 'device') == 'scipy':
      return self._run_scipy(itrs)
    return super()._run(itrs)

  def _active_problem(self):
    """as the base class, but the float64 active rows are mirrored on the host incrementally: one new
    row (S float64) crosses PCIe per iteration instead of the whole K x S active set"""
    idx, w = self._native.active()
    have = 0 if getattr(self, '_rows', None) is None else self._rows.shape[0]
    if idx.shape[0] > have:
      new = self._native.active_rows(have, idx.shape[0] - have)
      self._rows = new if have == 0 else np.vstack((self._rows, new))
    pos = np.flatnonzero(w > 0)
    pos = pos[np.argsort(idx[pos], kind='stable')]
    return idx, w, pos, np.ascontiguousarray(self._rows[pos].T)

  def reset(self):
    super().reset()
    self._rows = None

  def _run_scipy(self, itrs):
    events = []
    retried = False
    for _ in range(itrs):
      idx0, prev_w = self._native.active()
      nonempty = bool((prev_w > 0).any())
      prev_error = self.error()
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
