"""GIGA / Frank-Wolfe / OrthoPursuit on tiny problems: prints a digest of what each build did, so that a run under
compute-sanitizer (serialised warps, very different timing) can be compared with a normal run -- racecheck itself only sees
shared-memory hazards, a result that changes under it points at an ordering bug through global memory."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bayesiancoresets_b200 as bc
from conftest import lr_problem
for N, S in ((1500, 96), (4000, 256)):
  Z, theta = lr_problem(1, N, 6, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  for alg in ('GIGA', 'FrankWolfe', 'OrthoPursuit'):
    cs = bc.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg))
    cs.build(30); cs.build(25)
    ev = cs.snnls.last_events
    h = hashlib.sha1(repr([(e.code, e.f, e.error) for e in ev]).encode()).hexdigest()[:12]
    print('ENGINE', os.environ.get('BCG_ENGINE', '2'), 'N', N, 'S', S, alg, 'size', cs.snnls.size(), 'error %.12g' % cs.error(), 'digest', h,
          'filter16', cs.snnls._native.filter16_stats()[0], flush=True)
