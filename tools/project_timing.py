"""diagnostic: time the materialising projection (K3) of a device-resident dataset, specialised vs general kernel"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bench import lr_shard, lr_samples
N, d, S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000, 10, int(sys.argv[2]) if len(sys.argv) > 2 else 512
Z, th = lr_shard(0, 0, N, d)
theta = lr_samples(0, th, S)
ctx = bc.Context.default()
ds = bc.Dataset(Z, ctx=ctx)
y = np.random.RandomState(0).poisson(2., size=(N, 1)).astype(np.float64)
dsp = bc.Dataset(np.hstack((Z, y)), ctx=ctx)
for model, name, dset, si in ((bc._native.MODEL_LR, 'lr', ds, None), (bc._native.MODEL_POISSON, 'poisson', dsp, None),
                              (bc._native.MODEL_GAUSSIAN, 'gaussian', ds, np.eye(d))):
  for fast in ('2', '1', '0'):
    os.environ['BCG_PROJ_FAST'] = fast
    v = dset.project(model, theta, si, vecs=True)[0]
    del v
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
      v = dset.project(model, theta, si, vecs=True)[0]
      del v
    ctx.synchronize()
    dt = (time.perf_counter() - t0)/3
    print(json.dumps({'what': 'K3 materialising projection (whole call incl. 4NS-byte cudaMalloc/cudaFree)', 'model': name, 'N': N, 'd': d, 'S': S,
                      'specialised': int(fast), 'ms': dt*1e3, 'write_GBs': 4.*N*S/dt/1e9, 'ps_per_element': dt/(N*S)*1e12}), flush=True)
