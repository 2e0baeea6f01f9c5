"""time of one sampler call, device (bc.GaussianPosteriorSampler) vs NumPy (the oracle restatement), d = 200, S = 512"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from oracle import models
d, S, K = 200, 512, 5
rng = np.random.RandomState(0)
pts, w = rng.randn(K, d), rng.uniform(1., 1e5, size=K)
dev = bc.GaussianPosteriorSampler(np.zeros(d), np.eye(d), np.eye(d))
host = models.gaussian_sampler_w(np.zeros(d), np.eye(d), np.eye(d))
for name, f in (('device', dev), ('numpy', host)):
  f(S, w, pts)
  t0 = time.perf_counter()
  for _ in range(50):
    f(S, w, pts)
  print(name, 'ms per call: %.3f' % ((time.perf_counter() - t0)/50*1e3), flush=True)
t0 = time.perf_counter()
for _ in range(50):
  np.random.randn(S, d)
print('np.random.randn(S, d) alone: %.3f ms' % ((time.perf_counter() - t0)/50*1e3))
