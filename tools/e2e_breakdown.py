"""diagnostic: where the end-to-end construction time goes (upload / projection / solver setup / build / read-back)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bayesiancoresets_b200 import _native as nat
from bench import lr_shard, lr_samples, WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else 'lr_giga_N1e7_S512'
N, d, S = WORKLOADS[wl]
Z, th = lr_shard(0, 0, N, d)
theta = lr_samples(0, th, S)
ctx = bc.Context.default()
def T(label, f):
  ctx.synchronize(); t0 = time.perf_counter(); r = f(); ctx.synchronize()
  print('%-28s %8.1f ms' % (label, 1e3*(time.perf_counter() - t0)), flush=True); return r
for rep in range(2):
  ds = T('dataset upload (%.0f MB)' % (Z.nbytes/1e6), lambda: nat.Dataset(Z))
  vecs = T('project LR -> unit rows', lambda: ds.project(nat.MODEL_LR, theta, vecs=True)[0])
  b = vecs.sum(axis=0)
  sol = T('solver create', lambda: bc.snnls.GIGA(vecs.T, b))
  T('build(200)', lambda: sol.build(200))
  T('weights_sparse + error', lambda: (sol.weights_sparse(), sol.error()))
  T('project_sum (K3b path)', lambda: ds.project(nat.MODEL_LR, theta, colsum=True))
  del sol, vecs, ds
