"""one materialising projection per kernel variant, for ncu: python tools/k3_one.py [N] [d] [S] [lr|gauss]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
from bayesiancoresets_b200 import _native as nat
from bench import lr_shard, lr_samples
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
S = int(sys.argv[3]) if len(sys.argv) > 3 else 512
kind = sys.argv[4] if len(sys.argv) > 4 else 'lr'
if kind == 'lr':
  Z, th_true = lr_shard(0, 0, N, d); theta = lr_samples(0, th_true, S); model, si = nat.MODEL_LR, None
else:
  rng = np.random.RandomState(0); Z = rng.randn(N, d) + 1.; theta = rng.randn(S, d); model, si = nat.MODEL_GAUSSIAN, np.eye(d)
ds = nat.Dataset(Z)
for mma in ('1', '0'):
  os.environ['BCG_PROJ_MMA'] = mma
  for rep in range(2):
    v = ds.project(model, theta, si, vecs=True)[0]
    del v
