"""diagnostic: one model through the column-sum projection (for ncu captures)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
model, N, d, S = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.RandomState(0)
X = rng.randn(N, d)
th = rng.randn(S, d)/np.sqrt(d)
if model == 'lr':
  prj, data = bc.LogisticRegressionProjector(lambda n, w, p: th, S), X
elif model == 'poisson':
  prj, data = bc.PoissonProjector(lambda n, w, p: th, S), np.hstack((X, rng.poisson(1., (N, 1)).astype(float)))
else:
  prj, data = bc.GaussianProjector(lambda n, w, p: th, S, np.eye(d)), X
for _ in range(3):
  t0 = time.perf_counter(); prj.project_sum(data); print('%s %.2f ms' % (model, 1e3*(time.perf_counter() - t0)), flush=True)
