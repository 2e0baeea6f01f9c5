"""time per iteration of the never-materialising solver (HilbertCoreset(..., materialize=False)) at the headline size"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
N, d, S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000, 10, 512
Z, th = lr_shard(0, 0, N, d)
theta = lr_samples(0, th, S)
ctx = bc.Context.default()
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
free0 = ctx.mem_info()[0]
t0 = time.perf_counter()
cs = bc.HilbertCoreset(Z, prj, materialize=False)
t_setup = time.perf_counter() - t0
used = free0 - ctx.mem_info()[0]
cs.build(2)
t0 = time.perf_counter()
cs.build(10)
ctx.synchronize()
dt = (time.perf_counter() - t0)/10
print(json.dumps({'what': 'never-materialising GIGA', 'N': N, 'S': S, 'setup_s (upload + norms + b)': t_setup, 's_per_iter': dt,
                  'device_bytes': used, 'resident_matrix_would_be': 4*N*S, 'error': cs.error()}))
