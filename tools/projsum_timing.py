"""device+call time of the column-sum-only projection passes (K3b), N=1e6 d=200 S=512, three models, and LR/Poisson at d=128"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
from bayesiancoresets_b200 import _native as nat
rng = np.random.RandomState(0)
for N, d, S in ((1_000_000, 200, 512), (2_000_000, 128, 512)):
  X = rng.randn(N, d)/np.sqrt(d)*3.
  Zp = np.hstack((X, rng.poisson(2., size=(N, 1)).astype(float)))
  th = rng.randn(S, d)
  for name, model, Z, si in (('gauss', nat.MODEL_GAUSSIAN, X, np.eye(d)), ('lr', nat.MODEL_LR, X, None), ('poisson', nat.MODEL_POISSON, Zp, None)):
    ds = nat.Dataset(Z)
    ds.project(model, th, si, colsum=True)
    ts = []
    for _ in range(7):
      t0 = time.perf_counter(); ds.project(model, th, si, colsum=True); ts.append(time.perf_counter() - t0)
    print('%-8s N=%d d=%d S=%d: %.2f ms per pass (median of 7), %.1f TFLOP/s' % (name, N, d, S, np.median(ts)*1e3, 2.*N*d*S/np.median(ts)/1e12), flush=True)
    del ds
