"""HilbertCoreset on the device (reference: coreset/hilbert.py:7-48).

The projection is evaluated once -- on the device when the projector offers `project_device` --
and handed to the sparse-NNLS solver exactly as the reference does: `snnls(vecs.T, vecs.sum(0))`.
With a communicator the rows of `data` are this rank's shard of the N axis: b is all-reduced
once, the greedy loop exchanges candidates over NVLink inside the step kernel, and every rank
ends with the same (wts, idcs, pts)."""
import numpy as np
from ..snnls.giga import GIGA
from ..comm import SerialComm, gather_rows, shard_layout, local_part
from .. import _native as nat
from .coreset import Coreset


class HilbertCoreset(Coreset):
  def __init__(self, data, ll_projector, n_subsample=None, snnls=GIGA, comm=None, materialize=True, **kw):
    self.comm = comm or SerialComm()
    project = getattr(ll_projector, 'project_device', ll_projector.project)
    if not materialize:
      # never-materialising solver: the N x S matrix is not built; every selection pass re-evaluates
      # the rows from the raw data in float64 (8 N d bytes resident instead of 4 N S, ~10x the time per iteration)
      project = ll_projector.project_lazy
    if n_subsample is None:
      sub_idcs = None                                   # identity (the reference's np.arange(N) is an O(N) host array)
      vecs = project(data)
    else:
      # hilbert.py:13-22: sorted, de-duplicated subsample from the global RNG, zero rows removed.  With N-sharding
      # every rank draws the SAME subsample of the global index range (SPMD: identical seeded RNG state on the
      # ranks) and keeps the part that falls into its shard; the solver then shards the subsample.
      self._data_offset, n_total, _ = shard_layout(self.comm, data.shape[0])
      sub_global = np.unique(np.random.randint(n_total, size=n_subsample))
      _, mine = local_part(sub_global, self._data_offset, data.shape[0])
      # (the subsample is gathered on the host and projected once: uploading all of `data` for it would be wasted)
      sub_project = lambda idx: project(data[idx])
      vecs = sub_project(mine)
      nonzero = (vecs.norms() > 0.) if isinstance(vecs, nat.DeviceVecs) else (np.sqrt((vecs**2).sum(axis=1)) > 0.)
      if not nonzero.all():
        mine = mine[nonzero]
        vecs = sub_project(mine)
      if self.comm.world > 1:
        sub_idcs = np.concatenate(self.comm.allgather_object(mine + self._data_offset))
      else:
        sub_idcs = mine
    b = vecs.sum(axis=0)
    extra = {}
    if self.comm.world > 1:
      b = self.comm.allreduce_sum(b)
      extra['comm'] = self.comm
    self.snnls = snnls(vecs.T, b, **extra)
    self._sub_idcs = sub_idcs
    self.data = data
    super().__init__(**kw)

  @property
  def sub_idcs(self):
    """hilbert.py:11,16 -- materialised on demand in the identity case"""
    return np.arange(self.snnls.n_global) if self._sub_idcs is None else self._sub_idcs

  def reset(self):
    self.snnls.reset()
    super().reset()

  def _export(self):
    # hilbert.py:35-38: ascending index order, strictly positive weights
    idx, w = self.snnls.weights_sparse()
    self.wts = w
    if self.comm.world > 1:
      self.idcs = idx if self._sub_idcs is None else self._sub_idcs[idx]
      off = self.snnls.row_offset if self._sub_idcs is None else self._data_offset
      self.pts = gather_rows(self.comm, self.data, off, self.idcs)
    else:
      self.idcs = idx if self._sub_idcs is None else self._sub_idcs[idx]
      self.pts = self.data[self.idcs]

  def _build(self, itrs):
    self.snnls.build(itrs)
    self._export()

  def _optimize(self):
    self.snnls.optimize()
    self._export()

  def error(self):
    return self.snnls.error()
