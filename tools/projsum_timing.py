"""diagnostic: time the column-sum-only projection kernel alone (Gaussian = pure GEMM, Poisson = GEMM + link)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
N, d, S = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.RandomState(0)
X = rng.randn(N, d)
th = rng.randn(S, d)/np.sqrt(d)
for name, prj, data in (('gaussian', bc.GaussianProjector(lambda n, w, p: th, S, np.eye(d)), X),
                        ('lr', bc.LogisticRegressionProjector(lambda n, w, p: th, S), X),
                        ('poisson', bc.PoissonProjector(lambda n, w, p: th, S), np.hstack((X, rng.poisson(1., (N, 1)).astype(float))))):
  prj.project_sum(data)
  t0 = time.perf_counter()
  for _ in range(3):
    prj.project_sum(data)
  dt = (time.perf_counter() - t0)/3
  print('%s N=%d d=%d S=%d: %.2f ms per pass, %.2f TFLOP/s (GEMM flops only)' % (name, N, d, S, dt*1e3, 2.*N*d*S/dt/1e12), flush=True)
