"""device time of the materialising projection kernels (BCG_PROJ_TRACE=1 prints the CUDA-event time of each launch):
   python tools/k3_timing.py            # LR N=1e7 d=10 S=512 (headline), LR N=1e6 d=10 S=256, Gaussian N=1e6 d=200 S=512"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
os.environ['BCG_PROJ_TRACE'] = '1'
import numpy as np
import bayesiancoresets_b200 as bc
from bayesiancoresets_b200 import _native as nat
from bench import lr_shard, lr_samples
cases = [('lr', 10_000_000, 10, 512), ('lr', 1_000_000, 10, 256), ('gauss', 1_000_000, 200, 512), ('poisson', 1_000_000, 16, 512)]
for name, N, d, S in cases:
  rng = np.random.RandomState(0)
  if name == 'lr':
    Z, th_true = lr_shard(0, 0, N, d); theta = lr_samples(0, th_true, S); model, si = nat.MODEL_LR, None
  elif name == 'gauss':
    Z = rng.randn(N, d) + 1.; theta = rng.randn(S, d); model, si = nat.MODEL_GAUSSIAN, np.eye(d)
  else:
    Z = np.hstack((rng.randn(N, d)/3., rng.poisson(2., size=(N, 1)).astype(float))); theta = rng.randn(S, d)/3.; model, si = nat.MODEL_POISSON, None
  ds = nat.Dataset(Z)
  for env in ({'BCG_PROJ_MMA': '2'}, {'BCG_PROJ_MMA': '0'}):
    os.environ.update(env)
    print('==', name, N, d, S, env, flush=True)
    for rep in range(3):
      v = ds.project(model, theta, si, vecs=True)[0]
      del v
  del ds
