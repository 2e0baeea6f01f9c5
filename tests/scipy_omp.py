"""TEST HELPER: OrthoPursuit driven from the host with the reference's own `scipy.optimize.nnls` call for the reweight
(orthopursuit.py:37-42, snnls.py:41-78), on top of the device primitives bcg_solver_omp_select / _active_rows /
_set_weights.  Cross-check for the on-device Lawson-Hanson solver; never part of the product package."""
import numpy as np
from scipy.optimize import nnls


def run_scipy_omp(solver, itrs):
  """`solver`: a bayesiancoresets_b200.snnls.OrthoPursuit.  Returns the list of (code, f, error) events."""
  from bayesiancoresets_b200 import _native as nat
  native = solver._native
  events, retried, rows = [], False, None
  for _ in range(itrs):
    idx0, prev_w = native.active()
    nonempty = bool((prev_w > 0).any())
    prev_error = solver.error()
    f = native.omp_select()                               # orthopursuit.py:17-38 (w[f] = 1 on device)
    idx, w = native.active()
    have = 0 if rows is None else rows.shape[0]
    if idx.shape[0] > have:
      new = native.active_rows(have, idx.shape[0] - have)
      rows = new if have == 0 else np.vstack((rows, new))
    pos = np.flatnonzero(w > 0)
    pos = pos[np.argsort(idx[pos], kind='stable')]
    res = nnls(np.ascontiguousarray(rows[pos].T), solver.b, maxiter=100*solver.n_global)   # orthopursuit.py:40
    w_new = w.copy()
    w_new[pos] = res[0]
    native.set_weights(w_new)
    err = solver.error()
    ev = (nat.IT_OK, f, err)
    if nonempty and err > prev_error:                     # snnls.py:58-61
      revert = np.zeros(idx.shape[0])
      revert[:prev_w.shape[0]] = prev_w
      native.set_weights(revert)
      ev = (nat.IT_FAIL_MONOTONE, f, solver.error())
    elif nonempty:
      retried = False
    events.append(ev)
    if ev[0] != nat.IT_OK:
      if retried:
        break
      retried = True
  return events
