"""Projectors: evaluate per-datapoint log-likelihood vectors into the N x S matrix.

Same plug-in contract as the reference (bayesiancoresets/projector.py:4-32): `project(pts,
grad=False)` -> (n, S) [and (n, S, d)], `update(wts, pts)`.  The model projectors below run the
N-dependent work on the device (csrc/project_kernels.cuh) and additionally offer
`project_device(pts)` -> DeviceVecs, which `HilbertCoreset` uses to keep the matrix in HBM.
"""
import numpy as np
from . import _native as nat


class Projector(object):
  def project(self, pts, grad=False):
    raise NotImplementedError

  def update(self, wts, pts):
    raise NotImplementedError


class BlackBoxProjector(Projector):
  """User-callback projector (projector.py:11-32): the callbacks are arbitrary host Python, so
  this one evaluates on the host and the result is uploaded by the solver.  Use the model
  projectors below to evaluate on the device."""
  def __init__(self, sampler, projection_dimension, loglikelihood, grad_loglikelihood=None):
    self.projection_dimension = projection_dimension
    self.sampler = sampler
    self.loglikelihood = loglikelihood
    self.grad_loglikelihood = grad_loglikelihood
    self.update(np.array([]), np.array([]))

  def project(self, pts, grad=False):
    lls = self.loglikelihood(pts, self.samples)
    lls -= lls.mean(axis=1)[:, np.newaxis]
    if not grad:
      return lls
    if self.grad_loglikelihood is None:
      raise ValueError('grad_loglikelihood was requested but not initialized in BlackBoxProjector.project')
    glls = self.grad_loglikelihood(pts, self.samples)
    glls -= glls.mean(axis=2)[:, :, np.newaxis]
    return lls, glls

  def update(self, wts, pts):
    self.samples = self.sampler(self.projection_dimension, wts, pts)


class _DeviceModelProjector(Projector):
  """sampler(n, wts, pts) -> (n, D) stays on the host (RNG parity, projector.py:31-32); the samples are
  uploaded with every projection (<= S*d float64).  The data are uploaded once per array object
  (SparseVI / BatchPSVI project the same `data` at every optimisation step) and projected on the device:

    project_device(pts) -> DeviceVecs          the resident unit-row matrix (Hilbert / SparseVI selection)
    project_sum(pts)    -> (S,) ndarray        column sums only; the N x S matrix is never written
    project(pts)        -> (n, S) ndarray      float64 rows read back (drop-in Projector.project)"""
  _model = None

  def __init__(self, sampler, projection_dimension, ctx=None):
    self.projection_dimension = projection_dimension
    self.sampler = sampler
    self.ctx = ctx
    self._cache = (None, None)
    self.update(np.array([]), np.array([]))

  def update(self, wts, pts):
    self.samples = np.asarray(self.sampler(self.projection_dimension, wts, pts), dtype=np.float64)

  def _siginv(self):
    return None

  def _dataset(self, pts, cache):
    if cache and self._cache[0] is pts:
      return self._cache[1]
    ds = nat.Dataset(pts, ctx=self.ctx)
    if cache:
      self._cache = (pts, ds)
    return ds

  def project_device(self, pts, cache=False, sub=None):
    """sub: optional row indices into pts -- only those rows are projected, gathered on the device"""
    if not cache and sub is None and self._cache[0] is not pts:
      # one-off projection of a host array (HilbertCoreset): upload pipelined with the projection kernels
      return nat.DeviceVecs.project_host(self._model, pts, self.samples, self._siginv(), ctx=self.ctx)
    return self._dataset(pts, cache or sub is not None).project(self._model, self.samples, self._siginv(), vecs=True,
                                                                sub=sub)[0]

  def project_lazy(self, pts):
    """the never-materialising projection (`HilbertCoreset(..., materialize=False)`): only norms and column sums are
    computed now; the solver re-evaluates the rows from the raw data at every selection pass"""
    return self._dataset(pts, False).project_lazy(self._model, self.samples, self._siginv())

  def project_sum(self, pts, cache=True, sub=None):
    return self._dataset(pts, cache or sub is not None).project(self._model, self.samples, self._siginv(), colsum=True,
                                                                sub=sub)[2]

  def grad_contract(self, pts, w, resid):
    """-(1/S) sum_s w_k resid_s (glls[k,s,:] - mean_d glls[k,s,:]) for the K pseudo-points (bpsvi.py:53 with the
    centring of projector.py:26), on the device, without materialising the (K, S, d) array (csrc/pseudo_grad_kernel.cuh)"""
    return nat.pseudo_grad(self._model, np.atleast_2d(pts), self.samples, self._siginv(), w=w, resid=resid, ctx=self.ctx)[1]

  def _grad_centred(self, pts):
    """(n, S, d) gradients, centred over the LAST axis as projector.py:26 does, evaluated on the device"""
    return nat.pseudo_grad(self._model, pts, self.samples, self._siginv(), full=True, ctx=self.ctx)[0]

  def project(self, pts, grad=False):
    pts2 = np.atleast_2d(pts)
    lls = self._dataset(pts2, False).project(self._model, self.samples, self._siginv(), rows=True)[1]
    if not grad:
      return lls
    # (n, S, d) gradients are only ever requested for the K pseudo-points of BatchPSVI (bpsvi.py:37)
    return lls, self._grad_centred(pts2)


class LogisticRegressionProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_lr.py:25-32 with z_n = y_n x_n (gradients: model_lr.py:50-57)."""
  _model = nat.MODEL_LR


class GaussianProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_gaussian.py:4-10, known covariance (gradients: model_gaussian.py:12-15)."""
  _model = nat.MODEL_GAUSSIAN

  def __init__(self, sampler, projection_dimension, Siginv, ctx=None):
    self.Siginv = np.asarray(Siginv, dtype=np.float64)
    super().__init__(sampler, projection_dimension, ctx=ctx)

  def _siginv(self):
    return self.Siginv


class PoissonProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_poiss.py:25-38, z_n = [x_n, y_n] (gradients: model_poiss.py:58-67 with the
  reference's broadcast defect repaired and a zero d/dy column, as documented in SURVEY.md section 8c)."""
  _model = nat.MODEL_POISSON
