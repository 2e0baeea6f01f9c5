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
  """sampler(n, wts, pts) -> (n, D) stays on the host (RNG parity, projector.py:31-32); the
  samples are uploaded with every projection (<= S*d float64)."""
  def __init__(self, sampler, projection_dimension, ctx=None):
    self.projection_dimension = projection_dimension
    self.sampler = sampler
    self.ctx = ctx
    self.update(np.array([]), np.array([]))

  def update(self, wts, pts):
    self.samples = np.asarray(self.sampler(self.projection_dimension, wts, pts), dtype=np.float64)

  def project_device(self, pts):
    raise NotImplementedError

  def _grad(self, pts):
    raise ValueError('grad_loglikelihood was requested but is not available for this projector')

  def project(self, pts, grad=False):
    lls = self.project_device(pts).to_numpy()
    if not grad:
      return lls
    # (n, S, d) gradients are only ever requested for the K pseudo-points of BatchPSVI
    # (bpsvi.py:37); K*S*d host arithmetic, centred over the LAST axis as projector.py:26 does
    glls = self._grad(np.atleast_2d(pts))
    glls -= glls.mean(axis=2)[:, :, np.newaxis]
    return lls, glls


class LogisticRegressionProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_lr.py:25-32 with z_n = y_n x_n."""
  def project_device(self, pts):
    return nat.DeviceVecs.project_lr(pts, self.samples, ctx=self.ctx)

  def _grad(self, z):
    m = -z.dot(self.samples.T)                           # model_lr.py:50-57
    sig = np.where(m < 100, np.exp(np.minimum(m, 100.))/(1. + np.exp(np.minimum(m, 100.))), 1.)
    return sig[:, :, np.newaxis]*self.samples[np.newaxis, :, :]


class GaussianProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_gaussian.py:4-10 (known covariance)."""
  def __init__(self, sampler, projection_dimension, Siginv, ctx=None):
    self.Siginv = np.asarray(Siginv, dtype=np.float64)
    super().__init__(sampler, projection_dimension, ctx=ctx)

  def project_device(self, pts):
    return nat.DeviceVecs.project_gaussian(pts, self.samples, self.Siginv, ctx=self.ctx)

  def _grad(self, x):
    return self.samples.dot(self.Siginv)[np.newaxis, :, :] - x.dot(self.Siginv)[:, np.newaxis, :]


class PoissonProjector(_DeviceModelProjector):
  """log-likelihood of examples/common/model_poiss.py:25-38, z_n = [x_n, y_n]."""
  def project_device(self, pts):
    return nat.DeviceVecs.project_poisson(pts, self.samples, ctx=self.ctx)
