"""where does the host time of an OMP build go? (wall vs device, events, exact passes)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
N, S, K = int(float(sys.argv[1])), int(sys.argv[2]), 200
Z, th = lr_shard(0, 0, N, 10)
theta = lr_samples(0, th, S)
ctx = bc.Context.default()
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
vecs = prj.project_device(Z)
b = vecs.sum(axis=0)
for rep in range(2):
  sol = bc.snnls.OrthoPursuit(vecs.T, b)
  sol.build(5)
  ctx.synchronize()
  t0 = time.perf_counter()
  ev = sol._native.build(K, 1e-12)
  t1 = time.perf_counter()
  tm = sol._native.timing()
  print(json.dumps({'S': S, 'wall_ms': (t1 - t0)*1e3, 'device_ms': tm['build_ms'], 'events': len(ev), 'not_ok': sum(1 for e in ev if e.code != 0),
                    'exact': sol._native.exact_count(), 'launches': tm['step_launches'], 'scan_launches': tm['scan_launches']}), flush=True)
  del sol
