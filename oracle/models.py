"""Oracle (TEST INFRASTRUCTURE): float64 NumPy restatement of the model log-likelihoods and of
the black-box projection that builds the N x S matrix.

References (relative to the reference repository root):
  bayesiancoresets/projector.py:19-29            BlackBoxProjector.project (row-centring)
  examples/common/model_lr.py:25-32, 50-57       logistic regression
  examples/common/model_gaussian.py:4-15         Gaussian location model
  examples/common/model_poiss.py:25-38, 58-67    Poisson regression (softplus link)
"""
import numpy as np
from scipy.special import gammaln


# ---------------------------------------------------------------- logistic regression
def lr_loglik(z, th):
  """model_lr.py:25-32: ll[n,s] = -log(1+exp(-z_n.th_s)), linear branch when -z.th >= 100."""
  z = np.atleast_2d(z)
  th = np.atleast_2d(th)
  m = -z.dot(th.T)
  small = m < 100
  m[small] = -np.log1p(np.exp(m[small]))
  m[np.logical_not(small)] = -m[np.logical_not(small)]
  return m


def lr_grad_z_loglik(z, th):
  """model_lr.py:50-57: d ll[n,s] / d z_n = sigma(-z_n.th_s) * th_s  -> (n, S, d)."""
  z = np.atleast_2d(z)
  th = np.atleast_2d(th)
  m = -z.dot(th.T)
  small = m < 100
  m[small] = np.exp(m[small])/(1.+np.exp(m[small]))
  m[np.logical_not(small)] = 1.
  return m[:, :, np.newaxis]*th[np.newaxis, :, :]


# ---------------------------------------------------------------- Gaussian location model
def gaussian_loglik(x, th, Siginv, logdetSig):
  """model_gaussian.py:4-10"""
  x = np.atleast_2d(x)
  th = np.atleast_2d(th)
  xSx = (x*(x.dot(Siginv))).sum(axis=1)
  tSt = (th*(th.dot(Siginv))).sum(axis=1)
  xSt = x.dot(Siginv.dot(th.T))
  return -x.shape[1]/2*np.log(2*np.pi) - 1./2.*logdetSig - 1./2.*(xSx[:, np.newaxis] + tSt - 2*xSt)


def gaussian_grad_x_loglik(x, th, Siginv):
  """model_gaussian.py:12-15"""
  x = np.atleast_2d(x)
  th = np.atleast_2d(th)
  return th.dot(Siginv)[np.newaxis, :, :] - x.dot(Siginv)[:, np.newaxis, :]


def gaussian_weighted_post(th0, Sig0inv, Siginv, x, w):
  """model_gaussian.py:23-30: posterior mean and the upper factor U of the covariance (Sigp = U U^T)"""
  import scipy.linalg as sl
  LSigpInv = np.linalg.cholesky(Sig0inv + w.sum()*Siginv)
  USigp = sl.solve_triangular(LSigpInv, np.eye(LSigpInv.shape[0]), lower=True, overwrite_b=True, check_finite=False).T
  if w.shape[0] > 0:
    mup = np.dot(USigp.dot(USigp.T), np.dot(Sig0inv, th0) + np.dot(Siginv, (w[:, np.newaxis]*x).sum(axis=0)))
  else:
    mup = th0
  return mup, USigp, LSigpInv


def gaussian_sampler_w(th0, Sig0inv, Siginv):
  """examples/gaussian/main.py:107-113"""
  def sampler_w(n, wts, pts):
    if wts is None or pts is None or pts.shape[0] == 0:
      wts, pts = np.zeros(1), np.zeros((1, th0.shape[0]))
    muw, USigw, _ = gaussian_weighted_post(th0, Sig0inv, Siginv, pts, wts)
    return muw + np.random.randn(n, muw.shape[0]).dot(USigw.T)
  return sampler_w


def lr_log_joint(z, th, wts):
  """model_lr.py:34-39"""
  th = np.atleast_2d(th)
  return (wts[:, np.newaxis]*lr_loglik(z, th)).sum(axis=0) - 0.5*th.shape[1]*np.log(2.*np.pi) - 0.5*(th**2).sum(axis=1)


def lr_grad_th_log_joint(z, th, wts):
  """model_lr.py:41-48, 59-64"""
  z, th = np.atleast_2d(z), np.atleast_2d(th)
  m = -z.dot(th.T)
  idcs = m < 100
  m[idcs] = np.exp(m[idcs])/(1.+np.exp(m[idcs]))
  m[np.logical_not(idcs)] = 1.
  return -th + (wts[:, np.newaxis, np.newaxis]*(m[:, :, np.newaxis]*z[:, np.newaxis, :])).sum(axis=0)


def lr_hess_th_log_joint(z, th, wts):
  """model_lr.py:66-80"""
  z, th = np.atleast_2d(z), np.atleast_2d(th)
  m = -z.dot(th.T)
  idcs = m < 100
  m[idcs] = np.exp(m[idcs])/(1.+np.exp(m[idcs]))**2
  m[np.logical_not(idcs)] = 0.
  hl = -m[:, :, np.newaxis, np.newaxis]*z[:, np.newaxis, :, np.newaxis]*z[:, np.newaxis, np.newaxis, :]
  return np.tile(-np.eye(th.shape[1]), (th.shape[0], 1, 1)) + (wts[:, np.newaxis, np.newaxis, np.newaxis]*hl).sum(axis=0)


def get_laplace(wts, Z, mu0, log_joint, grad_log_joint, hess_log_joint):
  """examples/logistic_poisson_regression/main.py:16-41 (diag = False, no retries needed in the tests)"""
  from scipy.optimize import minimize
  from scipy.linalg import solve_triangular
  Zw, ww = Z[wts > 0, :], wts[wts > 0]
  res = minimize(lambda mu: -log_joint(Zw, mu, ww)[0], mu0, jac=lambda mu: -grad_log_joint(Zw, mu, ww)[0, :])
  mu = res.x
  LSigInv = np.linalg.cholesky(-hess_log_joint(Zw, mu, ww)[0, :, :])
  LSig = solve_triangular(LSigInv, np.eye(LSigInv.shape[0]), lower=True, overwrite_b=True, check_finite=False)
  return mu, LSig, LSigInv


# ---------------------------------------------------------------- Poisson regression
def poisson_log_rate(th, x):
  """model_poiss.py:25-30: s = log(softplus(x.th)), with s ~= x.th when x.th <= -100."""
  s = x.dot(th.T)
  big = s > -100
  s[big] = np.log(np.maximum(s[big], 0) + np.log1p(np.exp(-np.fabs(s[big]))))
  return s


def poisson_loglik(z, th):
  """model_poiss.py:32-38: z = [x, y]; ll = y*s - gammaln(y+1) - exp(s)."""
  th = np.atleast_2d(th)
  z = np.atleast_2d(z)
  x = z[:, :-1]
  y = np.tile(z[:, -1][:, np.newaxis], (1, th.shape[0]))
  s = poisson_log_rate(th, x)
  return y*s - gammaln(y+1) - np.exp(s)


def poisson_grad_z_loglik_fixed(z, th):
  """model_poiss.py:58-67 with the broadcast defect repaired.

  DEVIATION (documented in SURVEY.md section 8c): the reference multiplies by
  ``th[:, np.newaxis, :]`` (model_poiss.py:67), which raises a broadcast ValueError for n != S,
  and returns no derivative for the response column.  This oracle uses ``th[np.newaxis,:,:]``
  and appends a zero d/dy column so the result has the (n, S, d+1) shape BatchPSVI needs.
  """
  th = np.atleast_2d(th)
  z = np.atleast_2d(z)
  x = z[:, :-1]
  y = np.tile(z[:, -1][:, np.newaxis], (1, th.shape[0]))
  s = poisson_log_rate(th, x)
  g = y - np.exp(s)
  nz = np.exp(s) > 1e-15
  g[nz] = (y[nz]*np.exp(-s[nz]) - 1.)*(1. - np.exp(-np.exp(s[nz])))
  gx = g[:, :, np.newaxis]*th[np.newaxis, :, :]
  return np.concatenate((gx, np.zeros(gx.shape[:2] + (1,))), axis=2)


# ---------------------------------------------------------------- projection
def project(loglik, pts, samples, grad_loglik=None):
  """projector.py:19-29: evaluate and centre each row over the S samples.  The gradient branch
  centres over the LAST axis (d), as the reference does (projector.py:26)."""
  lls = loglik(pts, samples)
  lls -= lls.mean(axis=1)[:, np.newaxis]
  if grad_loglik is None:
    return lls
  glls = grad_loglik(pts, samples)
  glls -= glls.mean(axis=2)[:, :, np.newaxis]
  return lls, glls


class OracleProjector(object):
  """projector.py:11-32 (BlackBoxProjector) restated: sampler(n, wts, pts) -> (n, D)."""
  def __init__(self, sampler, projection_dimension, loglik, grad_loglik=None):
    self.projection_dimension = projection_dimension
    self.sampler = sampler
    self.loglik = loglik
    self.grad_loglik = grad_loglik
    self.update(np.array([]), np.array([]))

  def update(self, wts, pts):
    self.samples = self.sampler(self.projection_dimension, wts, pts)

  def project(self, pts, grad=False):
    if grad:
      if self.grad_loglik is None:
        raise ValueError('grad_loglikelihood was requested but not initialized')
      return project(self.loglik, pts, self.samples, self.grad_loglik)
    return project(self.loglik, pts, self.samples)


def poisson_log_joint(z, th, wts):
  """model_poiss.py:40-45"""
  th = np.atleast_2d(th)
  return (wts[:, np.newaxis]*poisson_loglik(z, th)).sum(axis=0) - 0.5*th.shape[1]*np.log(2.*np.pi) - 0.5*(th**2).sum(axis=1)


def poisson_grad_th_log_joint(z, th, wts):
  """model_poiss.py:47-56, 69-74"""
  th, z = np.atleast_2d(th), np.atleast_2d(z)
  x = z[:, :-1]
  y = np.tile(z[:, -1][:, np.newaxis], (1, th.shape[0]))
  s = poisson_log_rate(th, x)
  g = y - np.exp(s)
  idcs = np.exp(s) > 1e-15
  g[idcs] = (y[idcs]*np.exp(-s[idcs]) - 1.)*(1. - np.exp(-np.exp(s[idcs])))
  return -th + (wts[:, np.newaxis, np.newaxis]*(g[:, :, np.newaxis]*x[:, np.newaxis, :])).sum(axis=0)


def poisson_hess_th_log_joint(z, th, wts):
  """model_poiss.py:76-93"""
  th, z = np.atleast_2d(th), np.atleast_2d(z)
  x = z[:, :-1]
  y = np.tile(z[:, -1][:, np.newaxis], (1, th.shape[0]))
  s = poisson_log_rate(th, x)
  h = -(1.+y)*np.exp(s)
  idcs = np.exp(s) > 1e-15
  h[idcs] = (y[idcs]*np.exp(-s[idcs])*(1.-np.exp(-s[idcs]+np.exp(s[idcs]))+np.exp(-s[idcs])) - 1.)*(np.exp(-np.exp(s[idcs]))-np.exp(-2*np.exp(s[idcs])))
  hl = h[:, :, np.newaxis, np.newaxis]*x[:, np.newaxis, :, np.newaxis]*x[:, np.newaxis, np.newaxis, :]
  return np.tile(-np.eye(th.shape[1]), (th.shape[0], 1, 1)) + (wts[:, np.newaxis, np.newaxis, np.newaxis]*hl).sum(axis=0)
