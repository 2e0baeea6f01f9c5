"""Sampling baselines (reference: snnls/sampling.py:6-37).  O(1) work per iteration -- there is no scan to accelerate;
they exist so that code written against `bc.snnls.ImportanceSampling / UniformSampling` runs unchanged.  The draws stay
on the host (same global-RNG call order as the reference); the weights are written to the device solver as a sparse
active set, so error() / weights() / optimize() are the same device paths as for the greedy solvers."""
import numpy as np
from .base import SparseNNLS
from .. import _native as nat


class ImportanceSampling(SparseNNLS):
  _alg = nat.ALG_FW          # the device state is only used for ||A w - b|| and the NNLS re-solve

  def __init__(self, A, b, comm=None):
    if comm is not None and comm.world > 1:
      raise NotImplementedError('the sampling baselines are single-process')
    super().__init__(A, b, check_error_monotone=False)
    n = self.n_global
    self.cts = np.zeros(n)
    self.ps = self._vecs.norms()                       # sqrt((A**2).sum(axis=0)), sampling.py:11
    if np.any(self.ps > 0):
      self.ps /= self.ps.sum()
    else:
      self.ps = np.ones(n)/float(n)

  def _check_zero_columns(self, total):
    pass                                               # sampling.py never rejects zero columns

  def reset(self):
    super().reset()
    self.cts = np.zeros(self.n_global)

  def _select(self):
    return np.random.choice(self.ps.shape[0], p=self.ps)   # sampling.py:28-29

  def _run(self, itrs):
    events = []
    for _ in range(itrs):
      f = self._select()
      self.cts[f] += 1                                 # sampling.py:32-33
      events.append(nat.IterEvent(nat.IT_OK, int((self.cts > 0).sum()), int(f), 0., 0., 0.))
    idx = np.flatnonzero(self.cts > 0)
    self._native.set_active(idx, (self.cts[idx]/self.cts.sum())/self.ps[idx])
    err = self.error()
    for e in events:
      e.error = err
    return events


class UniformSampling(ImportanceSampling):
  def __init__(self, A, b, comm=None):
    super().__init__(A, b, comm=comm)
    self.ps = np.ones(self.n_global)/float(self.n_global)   # sampling.py:37
