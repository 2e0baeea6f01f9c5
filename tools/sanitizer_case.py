"""small end-to-end case for compute-sanitizer (memcheck / racecheck): specialised and general projection kernels,
persistent GIGA loop, OMP iterations with removals, optimize()"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bayesiancoresets_b200 as bc
from conftest import lr_problem
N, d = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 6
for S in (512, 256, 96):                  # 512 / 256: float16 pre-filter variants CH16 = 2 / 1 (512: the projection writes the copy)
  Z, theta = lr_problem(1, N, d, S)
  for mma, fast in (('2', '1'), ('0', '1'), ('0', '0')):            # DMMA / warp-per-row / general projection kernels
    os.environ['BCG_PROJ_MMA'], os.environ['BCG_PROJ_FAST'] = mma, fast
    prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
    cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.GIGA)
    cs.build(12)
    cs.snnls._native.set_force_exact(True)                          # exact float64 selection pass + relaunch
    cs.build(3)
  os.environ['BCG_PROJ_MMA'], os.environ['BCG_PROJ_FAST'] = '1', '1'
  y = np.random.RandomState(0).poisson(2., size=(N, 1)).astype(float)
  pp = bc.PoissonProjector(lambda n, w, p: theta, S)
  v = pp.project_device(np.hstack((Z, y)))
  cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
  cs.build(S//2 + 20)
  cs.optimize()
  print('S', S, 'omp size', cs.snnls.size(), 'error', cs.error(), flush=True)
# round-2 kernels: never-materialising solver, audit scorer, pseudo-point gradients, device sampler, streamed DMMA projection
from bayesiancoresets_b200 import _native as nat
Z, theta = lr_problem(2, N, d, 128)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 128)
for alg in (bc.snnls.GIGA, bc.snnls.OrthoPursuit):
  cs = bc.HilbertCoreset(Z, prj, snnls=alg, materialize=False)
  cs.build(8)
ds = nat.Dataset(Z)
sc, nr, csum = ds.audit(nat.MODEL_LR, theta, kind=nat.ALG_GIGA, dirs=np.random.RandomState(0).randn(2, 128), norms=True, colsum=True)
g, u = nat.pseudo_grad(nat.MODEL_POISSON, np.hstack((Z[:7], np.ones((7, 1)))), theta, w=np.ones(7), resid=np.ones(128), full=True)
rng = np.random.RandomState(3)
x = rng.randn(N, 40)
os.environ['BCG_PROJ_MMA'] = '2'
gp = bc.GaussianProjector(bc.GaussianPosteriorSampler(np.zeros(40), np.eye(40), np.eye(40)), 64, np.eye(40))
svi = bc.SparseVICoreset(x, gp, opt_itrs=3)
svi.build(2)
print('r2 kernels ok', float(sc.max()), float(u.sum()), svi.size(), flush=True)
print('SANITIZER CASE DONE')
