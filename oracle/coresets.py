"""Oracle (TEST INFRASTRUCTURE): float64 NumPy restatement of the coreset drivers that call the
hot path, and of the projected moment optimiser they use.

References (relative to the reference repository root):
  bayesiancoresets/coreset/coreset.py:8-64     Coreset shell (build guard, get, size)
  bayesiancoresets/coreset/hilbert.py:7-48     HilbertCoreset
  bayesiancoresets/coreset/sparsevi.py:7-79    SparseVICoreset
  bayesiancoresets/coreset/bpsvi.py:6-63       BatchPSVICoreset
  bayesiancoresets/util/opt.py:4-28            nn_opt

All random draws use the global legacy ``np.random`` stream in the same call order as the
reference (hilbert.py:16, sparsevi.py:33, bpsvi.py:17,33), so seeded runs reproduce it.
"""
import numpy as np
from .greedy import GigaOracle


def nn_opt(x0, grd, nn_idcs=None, opt_itrs=1000, step_sched=lambda i: 1./(i+1), b1=0.9, b2=0.999, eps=1e-8):
  """util/opt.py:4-28 (verbose branch omitted): bias-corrected moment steps + projection on x>=0."""
  x = x0.copy()
  m1 = np.zeros(x.shape[0])
  m2 = np.zeros(x.shape[0])
  for i in range(opt_itrs):
    g = grd(x)
    m1 = b1*m1 + (1.-b1)*g
    m2 = b2*m2 + (1.-b2)*g**2
    upd = step_sched(i)*m1/(1.-b1**(i+1))/(eps + np.sqrt(m2/(1.-b2**(i+1))))
    x -= upd
    if nn_idcs is None:
      x = np.maximum(x, 0.)
    else:
      x[nn_idcs] = np.maximum(x[nn_idcs], 0.)
  return x


class _Shell(object):
  """coreset.py:8-44"""
  def __init__(self):
    self.reached_numeric_limit = False
    self.wts = np.array([])
    self.idcs = np.array([], dtype=np.int64)
    self.pts = np.array([])

  def size(self):
    return int((self.wts > 0).sum())

  def get(self):
    if self.wts.shape[0] == 0:
      return np.array([]), np.array([]), np.array([])
    keep = self.wts > 0
    return self.wts[keep], self.pts[keep, :], self.idcs[keep]

  def build(self, itrs):
    if self.reached_numeric_limit or itrs <= 0:
      return
    self._build(itrs)


class HilbertOracle(_Shell):
  """hilbert.py:7-48"""
  def __init__(self, data, projector, n_subsample=None, solver=GigaOracle):
    if n_subsample is None:
      sub = np.arange(data.shape[0])
      vecs = projector.project(data)
    else:
      sub = np.unique(np.random.randint(data.shape[0], size=n_subsample))   # hilbert.py:16
      vecs = projector.project(data[sub])
      nonzero = np.sqrt((vecs**2).sum(axis=1)) > 0.                         # hilbert.py:20-22
      sub = sub[nonzero]
      vecs = vecs[nonzero, :]
    self.solver = solver(vecs.T, vecs.sum(axis=0))                          # hilbert.py:24
    self.sub_idcs = sub
    self.data = data
    super().__init__()

  def _export(self):
    w = self.solver.weights()
    self.wts = w[w > 0]
    self.idcs = self.sub_idcs[w > 0]
    self.pts = self.data[self.idcs]

  def _build(self, itrs):
    self.solver.build(itrs)
    self._export()

  def optimize(self):
    self.solver.optimize()
    self._export()

  def error(self):
    return self.solver.error()


def _tangent_space(data, projector, n_subsample, w, p, pts_for_core, with_grad):
  """sparsevi.py:23-42 / bpsvi.py:24-40"""
  projector.update(w, p)
  if n_subsample is None:
    sub = None
    vecs = projector.project(data)
    scaling = 1.
  else:
    sub = np.random.randint(data.shape[0], size=n_subsample)
    vecs = projector.project(data[sub])
    scaling = data.shape[0]/n_subsample
  if pts_for_core.size > 0:
    core = projector.project(pts_for_core, grad=True) if with_grad else projector.project(pts_for_core)
  elif with_grad:
    core = (np.zeros((0, vecs.shape[1])), np.zeros((0, vecs.shape[1], pts_for_core.shape[1])))
  else:
    core = np.zeros((0, vecs.shape[1]))
  return vecs, scaling, sub, core


class SparseVIOracle(_Shell):
  """sparsevi.py:7-79"""
  def __init__(self, data, projector, n_subsample_select=None, n_subsample_opt=None, opt_itrs=100,
               step_sched=lambda i: 1./(1.+i)):
    self.data = data
    self.projector = projector
    self.n_subsample_select = None if n_subsample_select is None else min(data.shape[0], n_subsample_select)
    self.n_subsample_opt = None if n_subsample_opt is None else min(data.shape[0], n_subsample_opt)
    self.step_sched = step_sched
    self.opt_itrs = opt_itrs
    super().__init__()

  def _build(self, itrs):
    for _ in range(itrs):
      self._select()
      self._optimize()

  def _select(self):
    """sparsevi.py:44-67"""
    vecs, scaling, sub, corevecs = _tangent_space(self.data, self.projector, self.n_subsample_select,
                                                  self.wts, self.pts, self.pts, False)
    resid = scaling*vecs.sum(axis=0) - self.wts.dot(corevecs)
    corrs = vecs.dot(resid) / np.sqrt((vecs**2).sum(axis=1)) / vecs.shape[1]
    corecorrs = np.fabs(corevecs.dot(resid) / np.sqrt((corevecs**2).sum(axis=1))) / corevecs.shape[1]
    if corecorrs.size == 0 or corrs.max() > corecorrs.max():
      f = sub[np.argmax(corrs)] if sub is not None else np.argmax(corrs)
      if f not in self.idcs:
        self.wts = np.append(self.wts, 0.)
        self.idcs = np.append(self.idcs, f)
        self.pts = self.data[f][np.newaxis, :] if self.pts.size == 0 else np.vstack((self.pts, self.data[f]))

  def _optimize(self):
    """sparsevi.py:69-76"""
    def grd(w):
      vecs, scaling, sub, corevecs = _tangent_space(self.data, self.projector, self.n_subsample_opt,
                                                    w, self.pts, self.pts, False)
      resid = scaling*vecs.sum(axis=0) - w.dot(corevecs)
      return -corevecs.dot(resid) / corevecs.shape[1]
    self.wts = nn_opt(self.wts, grd, opt_itrs=self.opt_itrs, step_sched=self.step_sched)

  def error(self):
    return 0.


class BatchPSVIOracle(_Shell):
  """bpsvi.py:6-63"""
  def __init__(self, data, projector, opt_itrs, n_subsample_opt=None, step_sched=lambda i: 1./(1.+i)):
    self.data = data
    self.projector = projector
    self.opt_itrs = opt_itrs
    self.n_subsample_opt = None if n_subsample_opt is None else min(data.shape[0], n_subsample_opt)
    self.step_sched = step_sched
    super().__init__()

  def _build(self, sz):
    init = np.random.choice(self.data.shape[0], size=sz, replace=False)     # bpsvi.py:17
    self.pts = self.data[init]
    self.wts = self.data.shape[0]/sz*np.ones(sz)
    self.idcs = -1*np.ones(sz)
    self._optimize()

  def gradient(self, x, sz, d):
    """bpsvi.py:46-55 -- one gradient evaluation ("grad step" of BASELINE config 5)."""
    w = x[:sz]
    p = x[sz:].reshape((sz, d))
    vecs, scaling, sub, core = _tangent_space(self.data, self.projector, self.n_subsample_opt, w, p, p, True)
    corevecs, pgrads = core
    resid = scaling*vecs.sum(axis=0) - w.dot(corevecs)
    wgrad = -corevecs.dot(resid) / corevecs.shape[1]
    ugrad = -(w[:, np.newaxis, np.newaxis]*pgrads*resid[np.newaxis, :, np.newaxis]).sum(axis=1)/corevecs.shape[1]
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
    return 0.
