"""Samplers that feed `Projector.update(wts, pts)` (projector.py:31-32) with the O(d^3) / O(S d^2) work on the device.

`GaussianPosteriorSampler` is the `sampler_w` of examples/gaussian/main.py:107-113: theta ~ N(mup, Sigp) for the weighted
Gaussian posterior of examples/common/model_gaussian.py:23-30.  The standard normals are drawn on the HOST with the
reference's own call (`np.random.randn(n, d)`), so a seeded run consumes the global RNG exactly as the reference does; the
Cholesky factorisation of the d x d posterior precision, its triangular inverse, the posterior mean and the S x d x d sample
transform run on the device (csrc/sampler_kernels.cuh).  SparseVI calls the sampler (1 + opt_itrs) times per build iteration."""
import numpy as np
from scipy.linalg import solve_triangular
from scipy.optimize import minimize
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


class LaplaceSampler(object):
  """The `sampler_w` of examples/logistic_poisson_regression/main.py:153-160: samples from the Laplace approximation of the
  posterior given the weighted coreset points, `get_laplace` (main.py:16-41).  The optimiser is SciPy's, as in the
  reference; every objective / gradient evaluation and the Hessian at the optimum -- the reductions over the K coreset
  points -- run on the device (csrc/sampler_kernels.cuh: glm_joint_kernel); the d x d Cholesky and the draw stay on the host
  (`np.random.randn(n, d)`, reference order).  model: 'lr' (examples/common/model_lr.py) or 'poisson' (model_poiss.py)."""
  def __init__(self, model, d, mu0=None, LSig0=None, ctx=None):
    self._model = {'lr': nat.MODEL_LR, 'poisson': nat.MODEL_POISSON}[model]
    self.d = int(d)
    self.mu0 = np.zeros(self.d) if mu0 is None else np.asarray(mu0, dtype=np.float64)
    self.LSig0 = np.eye(self.d) if LSig0 is None else np.asarray(LSig0, dtype=np.float64)
    self.ctx = ctx

  def get_laplace(self, wts, Z, mu_init):
    """(mu, LSig, LSigInv) of main.py:16-41 (diag = False)"""
    wts = np.asarray(wts, dtype=np.float64)
    Zw, ww = np.asarray(Z, dtype=np.float64)[wts > 0, :], wts[wts > 0]
    mu_init = np.asarray(mu_init, dtype=np.float64)
    trials = 10
    while True:
      try:
        res = minimize(lambda mu: -nat.glm_joint(self._model, Zw, ww, mu, ctx=self.ctx)[0], mu_init,
                       jac=lambda mu: -nat.glm_joint(self._model, Zw, ww, mu, ctx=self.ctx)[1])
        mu = res.x
      except Exception:
        mu_init = mu_init.copy()
        mu_init += np.sqrt((mu_init**2).sum())*0.1*np.random.randn(mu_init.shape[0])
        trials -= 1
        if trials <= 0:
          mu = mu_init
          break
        continue
      break
    H = nat.glm_joint(self._model, Zw, ww, mu, hess=True, ctx=self.ctx)[2]
    LSigInv = np.linalg.cholesky(-H)
    LSig = solve_triangular(LSigInv, np.eye(LSigInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
    return mu, LSig, LSigInv

  def __call__(self, n, wts, pts):
    if wts is None or pts is None or np.shape(pts)[0] == 0:            # main.py:154-156
      muw, LSigw = self.mu0, self.LSig0
    else:
      muw, LSigw, _ = self.get_laplace(wts, pts, np.zeros(self.d))
    return muw + np.random.randn(n, muw.shape[0]).dot(LSigw.T)        # main.py:159
