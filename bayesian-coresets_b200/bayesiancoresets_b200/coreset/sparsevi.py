"""SparseVICoreset on the device (reference: coreset/sparsevi.py:7-79).

Every build iteration re-samples the projector at the current coreset and re-projects ALL data
(1 + opt_itrs) times.  Here the data are uploaded once; the selection pass materialises the
N x S matrix on the device and takes the correlation arg-max with the scan kernel
(sparsevi.py:51,56-57), and the opt_itrs gradient passes only need project(data).sum(axis=0)
(sparsevi.py:71-72), which the column-sum-only projection computes without ever writing the
N x S matrix.  The K core points are projected in float64 and read back (K x S, tiny).
Host-side pieces (sampler, nn_opt, the K-sized algebra) stay NumPy, in the reference's call
order, so seeded runs consume the global RNG identically."""
import numpy as np
from ..util import nn_opt
from ..comm import SerialComm, shard_layout, local_part
from .. import _native as nat
from .coreset import Coreset


class SparseVICoreset(Coreset):
  def __init__(self, data, ll_projector, n_subsample_select=None, n_subsample_opt=None, opt_itrs=100,
               step_sched=lambda i: 1./(1.+i), comm=None, **kw):
    # with a communicator `data` is this rank's contiguous shard of the rows; column sums are all-reduced,
    # the selection arg-max is resolved over the ranks (lowest global index wins ties), indices are global
    self.comm = comm or SerialComm()
    self.row_offset, self.n_global, _ = shard_layout(self.comm, data.shape[0])
    self.data = data
    self.ll_projector = ll_projector
    self.n_subsample_select = None if n_subsample_select is None else min(self.n_global, n_subsample_select)
    self.n_subsample_opt = None if n_subsample_opt is None else min(self.n_global, n_subsample_opt)
    self.step_sched = step_sched
    self.opt_itrs = opt_itrs
    super().__init__(**kw)

  def _build(self, itrs):
    for _ in range(itrs):
      self._select()
      self._optimize()

  # ---- projections -----------------------------------------------------------------------------
  def _draw(self, n_subsample, w, p):
    """sparsevi.py:25-35: refresh the samples, then choose the rows of the tangent space"""
    self.ll_projector.update(w, p)
    if n_subsample is None:
      return 1., None
    # the draw is over the GLOBAL index range; with N-sharding every rank draws the same indices (SPMD: identical
    # seeded RNG state) and projects the ones that fall into its shard
    sub = np.random.randint(self.n_global, size=n_subsample)
    return self.n_global/n_subsample, sub

  def _local(self, sub):
    """(positions within the draw, local rows) of this rank's part of a global draw; (None, None) without subsampling"""
    if sub is None:
      return None, None
    if self.comm.world == 1:
      return np.arange(sub.shape[0]), sub
    return local_part(sub, self.row_offset, self.data.shape[0])

  def _corevecs(self, S):
    if self.pts.size > 0:
      return self.ll_projector.project(self.pts)          # sparsevi.py:38
    return np.zeros((0, S))

  def _device_vecs(self, sub):
    prj = self.ll_projector
    if hasattr(prj, 'project_device'):
      return prj.project_device(self.data, cache=True, sub=sub)     # subsample rows are gathered on the device
    rows = self.data if sub is None else self.data[sub]
    return nat.DeviceVecs.from_host(prj.project(rows))              # user-callback projector: host evaluation

  def _sum(self, sub):
    prj = self.ll_projector
    if hasattr(prj, 'project_sum'):
      local = prj.project_sum(self.data, cache=True, sub=sub)
    else:
      local = prj.project(self.data if sub is None else self.data[sub]).sum(axis=0)
    return self.comm.allreduce_sum(local) if self.comm.world > 1 else local

  def _row(self, f):
    """data row with global index f (owned by exactly one rank)"""
    if self.comm.world == 1:
      return np.asarray(self.data[f], dtype=np.float64)
    mine = self.row_offset <= f < self.row_offset + self.data.shape[0]
    piece = np.asarray(self.data[f - self.row_offset], dtype=np.float64) if mine else None
    return [p for p in self.comm.allgather_object(piece) if p is not None][0]

  # ---- sparsevi.py:44-67 -------------------------------------------------------------------------
  def _select(self):
    scaling, sub = self._draw(self.n_subsample_select, self.wts, self.pts)
    pos, loc = self._local(sub)
    vecs = self._device_vecs(loc)
    S = vecs.shape[1]
    corevecs = self._corevecs(S)
    total = vecs.sum(axis=0)
    if self.comm.world > 1:
      total = self.comm.allreduce_sum(total)
    resid = scaling*total - self.wts.dot(corevecs)
    # corrs = vecs.dot(resid)/||vecs_n||/S : arg-max and maximum on the device
    # (np.argmax over the subsample: the lowest POSITION in the draw wins ties; without subsampling position = index)
    best, dot = vecs.argmax_dot(resid)
    if best >= 0:
      best = int(pos[best]) if pos is not None else self.row_offset + best
    if self.comm.world > 1:
      cands = [c for c in self.comm.allgather_object((dot, best)) if c[1] >= 0]
      dot, best = max(cands, key=lambda c: (c[0], -c[1]))
    corr_max = dot/S
    corecorrs = np.fabs(corevecs.dot(resid)/np.sqrt((corevecs**2).sum(axis=1)))/S
    if corecorrs.size == 0 or corr_max > corecorrs.max():
      f = sub[best] if sub is not None else best
      if f not in self.idcs:
        self.wts = np.append(self.wts, 0.)
        self.idcs = np.append(self.idcs, np.int64(f))
        row = self._row(f)[np.newaxis, :]
        self.pts = row if self.pts.size == 0 else np.vstack((self.pts, row))

  # ---- sparsevi.py:69-76 -------------------------------------------------------------------------
  def _optimize(self):
    def grd(w):
      scaling, sub = self._draw(self.n_subsample_opt, w, self.pts)
      colsum = self._sum(self._local(sub)[1])
      corevecs = self._corevecs(colsum.shape[0])
      resid = scaling*colsum - w.dot(corevecs)
      return -corevecs.dot(resid)/corevecs.shape[1]
    self.wts = nn_opt(self.wts, grd, opt_itrs=self.opt_itrs, step_sched=self.step_sched)

  def error(self):
    return 0.   # as the reference (sparsevi.py:78-79): no KL estimate
