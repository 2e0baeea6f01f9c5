"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the development container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

It imports `bayesiancoresets` from /root/reference and the example models from
/root/reference/examples/common, runs them on seeded inputs (legacy global np.random stream,
the one the reference itself uses) and stores inputs' seeds/recipes plus the reference outputs.
The fixtures pin (1) oracle/*.py bit-for-bit and (2) the CUDA path within the tolerances
written in tests/.
"""
import os
import sys
import logging
import numpy as np

REF = os.environ.get('BC_REFERENCE', '/root/reference')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'examples', 'common'))

import bayesiancoresets as bc          # noqa: E402
import model_lr                        # noqa: E402
import model_gaussian                  # noqa: E402
import model_poiss                     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
ALGS = {'giga': bc.snnls.GIGA, 'fw': bc.snnls.FrankWolfe, 'omp': bc.snnls.OrthoPursuit}


class IDProjector(bc.Projector):
  def update(self, wts, pts):
    pass

  def project(self, pts, grad=False):
    return pts


def traced_build(solver, itrs):
  """run solver.build(itrs) recording the selected index and the error after every iteration"""
  sel, errs = [], []
  orig_select, orig_reweight = solver._select, solver._reweight

  def select():
    f = orig_select()
    sel.append(int(f))
    return f

  def reweight(f):
    orig_reweight(f)
    errs.append(float(solver.error()))
  solver._select, solver._reweight = select, reweight
  solver.build(itrs)
  solver._select, solver._reweight = orig_select, orig_reweight
  return np.array(sel, dtype=np.int64), np.array(errs)


def run_hilbert(X, projector, alg, itrs, **kw):
  cs = bc.HilbertCoreset(X, projector, snnls=ALGS[alg], **kw)
  sel, errs = traced_build(cs.snnls, itrs)
  # the coreset-level export normally happens in HilbertCoreset._build
  w = cs.snnls.weights()
  return dict(sel=sel, errs=errs, w=w, final_error=float(cs.snnls.error()), size=int(cs.snnls.size()),
              limit=bool(cs.snnls.reached_numeric_limit), sub_idcs=np.asarray(cs.sub_idcs))


def lr_problem(seed, N, d, S, spread=0.1):
  """SURVEY 8d C2 recipe (simple_lr/main.py:22-35 + Gaussian stand-in for the Laplace sampler)."""
  np.random.seed(seed)
  X = np.random.randn(N, d)
  th_true = np.random.randn(d)
  y = (np.random.rand(N) <= 1./(1.+np.exp(-X.dot(th_true)))).astype(np.float64)
  y[y == 0] = -1.
  Z = y[:, np.newaxis]*X
  theta = th_true + spread*np.random.randn(S, d)
  return Z, theta


def save(name, **arrs):
  path = os.path.join(OUT, name + '.npz')
  np.savez_compressed(path, **arrs)
  print('wrote', os.path.relpath(path), {k: np.asarray(v).shape for k, v in arrs.items()})


def main():
  os.makedirs(OUT, exist_ok=True)
  logging.getLogger().setLevel(logging.CRITICAL)

  # ---- C1: synthetic vectors, N=1000 S=50, build(100)  (examples/synthetic_vectors/main.py:82-99)
  for alg in ALGS:
    np.random.seed(1)
    X = np.random.randn(1000, 50)
    r = run_hilbert(X, IDProjector(), alg, 100)
    save('c1_normal_' + alg, seed=1, N=1000, S=50, itrs=100, **r)

  # ---- axis data X = I_12 (synthetic_vectors/main.py:64-65): exact ties, lowest index wins, then the
  #      numeric limit is reached
  for alg in ALGS:
    X = np.eye(12)
    r = run_hilbert(X, IDProjector(), alg, 20)
    save('axis12_' + alg, N=12, S=12, itrs=20, **r)

  # ---- LR projection + greedy, small (oracle-sized)
  Z, theta = lr_problem(0, 3000, 6, 96)
  prj = bc.BlackBoxProjector(lambda n, w, p: theta, 96, model_lr.log_likelihood, model_lr.grad_z_log_likelihood)
  vecs = prj.project(Z)
  save('lr_project_small', seed=0, N=3000, d=6, S=96, Z=Z, theta=theta, vecs=vecs)
  for alg in ALGS:
    r = run_hilbert(Z, prj, alg, 80)
    save('lr_small_' + alg, seed=0, N=3000, d=6, S=96, itrs=80, **r)
  lls, glls = prj.project(Z[:7], grad=True)
  save('lr_project_grad', lls=lls, glls=glls)

  # LR with saturated logits (exercise the m >= 100 branch, model_lr.py:29-31)
  np.random.seed(5)
  Zb = 60.*np.random.randn(200, 4)
  thb = 2.*np.random.randn(24, 4)
  prjb = bc.BlackBoxProjector(lambda n, w, p: thb, 24, model_lr.log_likelihood)
  save('lr_project_saturated', seed=5, Z=Zb, theta=thb, vecs=prjb.project(Zb))

  # ---- Hilbert with subsampling (hilbert.py:13-22)
  Z, theta = lr_problem(2, 2000, 5, 64)
  prj = bc.BlackBoxProjector(lambda n, w, p: theta, 64, model_lr.log_likelihood)
  np.random.seed(77)
  r = run_hilbert(Z, prj, 'giga', 40, n_subsample=700)
  save('lr_subsample_giga', seed=2, sub_seed=77, N=2000, d=5, S=64, n_subsample=700, itrs=40, **r)

  # ---- Gaussian projection (model_gaussian.py:4-10)
  np.random.seed(3)
  d = 12
  x = np.random.multivariate_normal(np.ones(d), np.eye(d), 400)
  L = np.random.randn(d, d)
  Sig = L.dot(L.T)/d + np.eye(d)
  Siginv = np.linalg.inv(Sig)
  logdet = np.linalg.slogdet(Sig)[1]
  thg = np.random.randn(40, d)
  prjg = bc.BlackBoxProjector(lambda n, w, p: thg, 40,
                              lambda x_, th_: model_gaussian.log_likelihood(x_, th_, Siginv, logdet))
  save('gaussian_project_small', seed=3, x=x, theta=thg, Siginv=Siginv, logdetSig=logdet, vecs=prjg.project(x))

  # ---- Poisson projection (model_poiss.py:25-38)
  np.random.seed(4)
  Xp = np.hstack((np.random.randn(300, 5), np.ones((300, 1))))
  thp_true = 0.5*np.random.randn(6)
  yp = np.random.poisson(np.log1p(np.exp(Xp.dot(thp_true)))).astype(np.float64)
  Zp = np.hstack((Xp, yp[:, np.newaxis]))
  thp = thp_true + 0.3*np.random.randn(48, 6)
  prjp = bc.BlackBoxProjector(lambda n, w, p: thp, 48, model_poiss.log_likelihood)
  save('poisson_project_small', seed=4, Z=Zp, theta=thp, vecs=prjp.project(Zp))
  # extreme linear predictors (exercise the s <= -100 branch, model_poiss.py:27-29)
  Zpx = Zp.copy()
  Zpx[:, :5] *= 80.
  save('poisson_project_extreme', Z=Zpx, theta=thp, vecs=prjp.project(Zpx))

  # ---- SparseVI (sparsevi.py) with the Gaussian model and the exact weighted posterior sampler
  #      (examples/gaussian/main.py:107-113, model_gaussian.py:23-30)
  np.random.seed(6)
  d = 5
  xs = np.random.multivariate_normal(np.ones(d), np.eye(d), 300)
  th0, Sig0inv, SigLinv = np.zeros(d), np.eye(d), np.eye(d)

  def sampler_w(n, wts, pts):
    if wts is None or pts is None or pts.shape[0] == 0:
      wts, pts = np.zeros(1), np.zeros((1, d))
    muw, USigw, _ = model_gaussian.weighted_post(th0, Sig0inv, SigLinv, pts, wts)
    return muw + np.random.randn(n, muw.shape[0]).dot(USigw.T)
  prjs = bc.BlackBoxProjector(sampler_w, 30, lambda x_, th_: model_gaussian.log_likelihood(x_, th_, SigLinv, 0.))
  svi = bc.SparseVICoreset(xs, prjs, opt_itrs=15, step_sched=lambda i: 1./(1.+i))
  svi.build(6)
  w, p, i = svi.get()
  save('sparsevi_gaussian', seed=6, N=300, d=d, S=30, itrs=6, opt_itrs=15, wts=w, pts=p, idcs=i,
       raw_wts=svi.wts, raw_idcs=svi.idcs)

  # ---- BatchPSVI gradient with the LR model (works unmodified)
  Z, theta = lr_problem(8, 500, 4, 32)
  np.random.seed(9)
  prjq = bc.BlackBoxProjector(lambda n, w, p: theta, 32, model_lr.log_likelihood, model_lr.grad_z_log_likelihood)
  bp = bc.BatchPSVICoreset(Z, prjq, opt_itrs=10)
  bp.build(8)
  save('bpsvi_lr', seed=8, build_seed=9, N=500, d=4, S=32, sz=8, opt_itrs=10, wts=bp.wts, pts=bp.pts)

  # ---- weighted Gaussian posterior + the sampler_w draw (model_gaussian.py:23-30, gaussian/main.py:107-113)
  rng = np.random.RandomState(12)
  d = 9
  B0, B1 = rng.randn(d, d), rng.randn(d, d)
  S0inv, S1inv = B0.dot(B0.T) + d*np.eye(d), B1.dot(B1.T) + d*np.eye(d)
  mu0 = rng.randn(d)
  ptsw, ww = rng.randn(6, d), rng.uniform(0.2, 3., size=6)
  mup, USigp, LSigpInv = model_gaussian.weighted_post(mu0, S0inv, S1inv, ptsw, ww)
  np.random.seed(21)
  draw = mup + np.random.randn(40, d).dot(USigp.T)
  mup0, USigp0, _ = model_gaussian.weighted_post(mu0, S0inv, S1inv, np.zeros((0, d)), np.zeros(0))
  save('gaussian_weighted_post', mu0=mu0, Sig0inv=S0inv, Siginv=S1inv, pts=ptsw, wts=ww, mup=mup, USigp=USigp, draw_seed=21,
       draw=draw, mup_empty=mup0, USigp_empty=USigp0)

  # ---- log-joint / gradient / Hessian of the GLM models over weighted points (the Laplace sampler's reductions:
  #      examples/logistic_poisson_regression/main.py:16-41 with model_lr.py:34-80, model_poiss.py:40-93)
  rng = np.random.RandomState(14)
  Zl, thl = lr_problem(15, 40, 5, 3)
  wl = rng.uniform(0., 4., size=40)
  wl[::7] = 0.
  Zq = np.hstack((rng.randn(40, 5), rng.poisson(3., size=(40, 1)).astype(np.float64)))
  thq = 0.4*rng.randn(3, 5)
  save('glm_joint', Z_lr=Zl, th_lr=thl, w=wl, lr_value=model_lr.log_joint(Zl, thl, wl), lr_grad=model_lr.grad_th_log_joint(Zl, thl, wl),
       lr_hess=model_lr.hess_th_log_joint(Zl, thl, wl), Z_poiss=Zq, th_poiss=thq, poiss_value=model_poiss.log_joint(Zq, thq, wl),
       poiss_grad=model_poiss.grad_th_log_joint(Zq, thq, wl), poiss_hess=model_poiss.hess_th_log_joint(Zq, thq, wl))


if __name__ == '__main__':
  main()
