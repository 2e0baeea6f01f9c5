"""diagnostic: OrthoPursuit iterations/s (C2 shape by default) with the wide and the narrow K x S products"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
N, d, S, K = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000, 10, int(sys.argv[2]) if len(sys.argv) > 2 else 256, 200
modes = sys.argv[3].split(',') if len(sys.argv) > 3 else ['1', '0']
Z, th = lr_shard(0, 0, N, d)
theta = lr_samples(0, th, S)
ctx = bc.Context.default()
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
vecs = prj.project_device(Z)
b = vecs.sum(axis=0)
for wide in modes:
  os.environ['BCG_OMP_WIDE'] = wide
  sol = bc.snnls.OrthoPursuit(vecs.T, b)
  sol.build(5)
  ctx.synchronize()
  t0 = time.perf_counter()
  sol.build(K)
  ctx.synchronize()
  dt = time.perf_counter() - t0
  tm = sol._native.timing()
  print(json.dumps({'what': 'omp', 'N': N, 'S': S, 'wide': int(wide), 'iters': K, 'iters_per_s': K/dt, 'ms_per_iter': 1e3*dt/K,
                    'device_ms_per_iter': tm['build_ms']/K, 'size': int(sol.size()), 'error': sol.error(),
                    'scan_roofline_ms': 4.*N*S/6543.1e9*1e3}), flush=True)
  del sol
