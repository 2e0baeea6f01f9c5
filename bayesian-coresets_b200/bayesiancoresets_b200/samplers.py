"""Samplers that feed `Projector.update(wts, pts)` (projector.py:31-32) with the O(d^3) / O(S d^2) work on the device.

`GaussianPosteriorSampler` is the `sampler_w` of examples/gaussian/main.py:107-113: theta ~ N(mup, Sigp) for the weighted
Gaussian posterior of examples/common/model_gaussian.py:23-30.  The standard normals are drawn on the HOST with the
reference's own call (`np.random.randn(n, d)`), so a seeded run consumes the global RNG exactly as the reference does; the
Cholesky factorisation of the d x d posterior precision, its triangular inverse, the posterior mean and the S x d x d sample
transform run on the device (csrc/sampler_kernels.cuh).  SparseVI calls the sampler (1 + opt_itrs) times per build iteration."""
import numpy as np
from . import _native as nat


class GaussianPosteriorSampler(object):
  def __init__(self, mu0, Sig0inv, Siginv, ctx=None):
    self.mu0 = np.asarray(mu0, dtype=np.float64)
    self.Sig0inv = np.asarray(Sig0inv, dtype=np.float64)
    self.Siginv = np.asarray(Siginv, dtype=np.float64)
    self.ctx = ctx

  def weighted_post(self, pts, wts):
    """(mup, USigp) of model_gaussian.py:23-30 (Sigp = USigp USigp^T)"""
    _, mup, U = nat.gaussian_post_sample(self.mu0, self.Sig0inv, self.Siginv, pts, wts, np.zeros((0, self.mu0.shape[0])),
                                         want_post=True, ctx=self.ctx)
    return mup, U

  def __call__(self, n, wts, pts):
    if wts is None or pts is None or np.shape(pts)[0] == 0:            # gaussian/main.py:108-110
      wts, pts = np.zeros(1), np.zeros((1, self.mu0.shape[0]))
    E = np.random.randn(n, self.mu0.shape[0])                          # gaussian/main.py:112 (host RNG, reference order)
    return nat.gaussian_post_sample(self.mu0, self.Sig0inv, self.Siginv, pts, wts, E, ctx=self.ctx)
