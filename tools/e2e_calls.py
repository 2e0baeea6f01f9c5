"""which C-ABI call (or which gap between calls) carries the sporadic 0.1 - 0.5 s stalls of bench.py's end-to-end job?
Same prelude as bench.py (torch dgemm peak, NVML clock sampler, one resident solver), then the job 10 times from a pinned
and from a pageable source, interleaved, with every call into libbcg_b200.so timed on the host.

  python tools/e2e_calls.py [notorch]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
from bayesiancoresets_b200 import _native as nat
import bench

LOG = []


class TimedLib(object):
  def __init__(self, L):
    self._L = L
    self._cache = {}

  def __getattr__(self, name):
    fn = self._cache.get(name)
    if fn is None:
      raw = getattr(self._L, name)

      def fn(*a, _raw=raw, _name=name):
        t0 = time.perf_counter()
        r = _raw(*a)
        LOG.append((_name, t0, time.perf_counter()))
        return r
      self._cache[name] = fn
    return fn


N, d, S, steps = 10_000_000, 10, 512, 20
ctx = bc.Context.default(0)
if 'notorch' not in sys.argv:
  print('dgemm peak', bench.measured_f64_peak(0), flush=True)
Z, th = bench.lr_shard(0, 0, N, d)
theta = bench.lr_samples(0, th, S)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.GIGA); cs.snnls.build(5)
smp = bench.ClockSampler(0); smp.start()
cs.snnls.build(steps); ctx.synchronize()
print('clocks', smp.stop(), flush=True)
del cs
nat.lib()
nat._lib = TimedLib(nat._lib)
srcs = (('pinned', bc.pinned_copy(Z)), ('pageable', Z))
for rep in range(10):                       # the two sources interleaved, so that box-level noise hits both alike
  for label, src in srcs:
    del LOG[:]
    t0 = time.perf_counter()
    cs2 = bc.HilbertCoreset(src, prj, snnls=bc.snnls.GIGA)
    t1 = time.perf_counter()
    cs2.build(steps)
    t2 = time.perf_counter()
    wts, pts, idcs = cs2.get()
    err = cs2.error()
    ctx.synchronize()
    t3 = time.perf_counter()
    del cs2
    calls = sorted(((b - a) * 1e3, n) for n, a, b in LOG)[::-1]
    print(json.dumps({'src': label, 'rep': rep, 'total_ms': round((t3 - t0) * 1e3, 1), 'ctor': round((t1 - t0) * 1e3, 1),
                      'build': round((t2 - t1) * 1e3, 1), 'export': round((t3 - t2) * 1e3, 1),
                      'top_calls': [(round(c, 1), n) for c, n in calls[:2]]}), flush=True)
