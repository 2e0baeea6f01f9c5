"""like racecheck_results_case.py, for the other kernels with block-wide logic: never-materialising solver, SparseVI on the
device sampler (left-looking Cholesky), BatchPSVI gradient, pseudo-point gradients, Laplace reductions, optimize()"""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bayesiancoresets_b200 as bc
from bayesiancoresets_b200 import _native as nat
from conftest import lr_problem


def dig(*arrs):
  h = hashlib.sha1()
  for a in arrs: h.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
  return h.hexdigest()[:12]


Z, theta = lr_problem(2, 1500, 6, 128)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 128)
for alg in ('GIGA', 'OrthoPursuit'):
  cs = bc.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg), materialize=False)
  cs.build(12)
  print('CASE lazy', alg, dig([e.f for e in cs.snnls.last_events], [e.error for e in cs.snnls.last_events]), flush=True)
cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.GIGA)
cs.build(20); cs.optimize()
print('CASE optimize', dig(cs.snnls.weights(), cs.error()), flush=True)
rng = np.random.RandomState(3)
x = rng.randn(1200, 24)
np.random.seed(5)
gp = bc.GaussianProjector(bc.GaussianPosteriorSampler(np.zeros(24), np.eye(24), np.eye(24)), 64, np.eye(24))
svi = bc.SparseVICoreset(x, gp, opt_itrs=5)
svi.build(4)
print('CASE sparsevi', dig(svi.idcs, svi.wts), flush=True)
np.random.seed(6)
y = np.random.RandomState(0).poisson(2., size=(1500, 1)).astype(float)
pp = bc.PoissonProjector(lambda n, w, p: theta[:64], 64)
bp = bc.BatchPSVICoreset(np.hstack((Z, y)), pp, opt_itrs=3)
bp.build(5)
print('CASE bpsvi', dig(bp.wts, bp.pts), flush=True)
g, u = nat.pseudo_grad(nat.MODEL_LR, Z[:9], theta, w=np.arange(1., 10.), resid=np.linspace(-1, 1, 128), full=True)
print('CASE pseudo_grad', dig(g, u), flush=True)
v, gr, H = nat.glm_joint(nat.MODEL_LR, Z[:300], np.linspace(0.5, 3., 300), theta[0], hess=True)
print('CASE glm_joint', dig(v, gr, H), flush=True)
