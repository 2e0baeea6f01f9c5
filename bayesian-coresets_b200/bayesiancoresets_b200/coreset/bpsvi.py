"""BatchPSVICoreset on the device (reference: coreset/bpsvi.py:6-63).

One gradient evaluation needs project(data).sum(axis=0) over all N data points -- computed by the
column-sum-only projection on the device, the N x S matrix is never written -- plus the
projection and (K, S, d) gradients of the K pseudo-points, which are K-sized host algebra."""
import numpy as np
from ..util import nn_opt
from ..comm import SerialComm, shard_layout, local_part
from .coreset import Coreset


class BatchPSVICoreset(Coreset):
  def __init__(self, data, ll_projector, opt_itrs, n_subsample_opt=None, step_sched=lambda i: 1./(1.+i), comm=None,
               **kw):
    # with a communicator `data` is this rank's shard of the rows: the data column sums are all-reduced
    # (one S-vector per gradient step); the K pseudo-points and their optimiser state are replicated
    self.comm = comm or SerialComm()
    self.row_offset, self.n_global, _ = shard_layout(self.comm, data.shape[0])
    self.data = data
    self.ll_projector = ll_projector
    self.opt_itrs = opt_itrs
    self.n_subsample_opt = None if n_subsample_opt is None else min(self.n_global, n_subsample_opt)
    self.step_sched = step_sched
    super().__init__(**kw)

  def _build(self, sz):
    init_idcs = np.random.choice(self.n_global, size=sz, replace=False)          # bpsvi.py:17 (same draw on every rank)
    if self.comm.world == 1:
      self.pts = np.array(self.data[init_idcs], dtype=np.float64)
    else:
      from ..comm import gather_rows
      self.pts = np.array(gather_rows(self.comm, self.data, self.row_offset, init_idcs), dtype=np.float64)
    self.wts = self.n_global/sz*np.ones(sz)
    self.idcs = -1*np.ones(sz)
    self._optimize()

  def _data_sum(self, w, p):
    """bpsvi.py:24-35: refresh the samples, column sums of the (sub)sampled tangent space"""
    prj = self.ll_projector
    prj.update(w, p)
    sub, scaling = None, 1.
    if self.n_subsample_opt is not None:
      # drawn over the GLOBAL index range, identically on every rank (SPMD); each rank projects its part
      sub = np.random.randint(self.n_global, size=self.n_subsample_opt)
      scaling = self.n_global/self.n_subsample_opt
      if self.comm.world > 1:
        sub = local_part(sub, self.row_offset, self.data.shape[0])[1]
    if hasattr(prj, 'project_sum'):
      local = prj.project_sum(self.data, cache=True, sub=sub)      # rows gathered on the device
    else:
      local = prj.project(self.data if sub is None else self.data[sub]).sum(axis=0)
    if self.comm.world > 1:
      local = self.comm.allreduce_sum(local)
    return scaling*local

  def gradient(self, x, sz, d):
    """bpsvi.py:46-55 -- one gradient evaluation (the "grad step" of BASELINE config 5)"""
    w = x[:sz]
    p = x[sz:].reshape((sz, d))
    total = self._data_sum(w, p)
    prj = self.ll_projector
    if hasattr(prj, 'grad_contract'):
      corevecs = prj.project(p)
      resid = total - w.dot(corevecs)
      ugrad = prj.grad_contract(p, w, resid)              # (K x S).(S x d) instead of the (K, S, d) array
    else:
      corevecs, pgrads = prj.project(p, grad=True)
      resid = total - w.dot(corevecs)
      ugrad = -(w[:, np.newaxis, np.newaxis]*pgrads*resid[np.newaxis, :, np.newaxis]).sum(axis=1)/corevecs.shape[1]
    wgrad = -corevecs.dot(resid)/corevecs.shape[1]
    return np.hstack((wgrad, ugrad.reshape(sz*d)))

  def _optimize(self):
    sz = self.wts.shape[0]
    d = self.pts.shape[1]
    x0 = np.hstack((self.wts, self.pts.reshape(sz*d)))
    xf = nn_opt(x0, lambda x: self.gradient(x, sz, d), nn_idcs=np.arange(sz), opt_itrs=self.opt_itrs,
                step_sched=self.step_sched)
    self.wts = xf[:sz]
    self.pts = xf[sz:].reshape((sz, d))

  def error(self):
    return 0.   # as the reference (bpsvi.py:62-63)
