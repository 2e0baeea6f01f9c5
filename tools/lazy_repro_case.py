"""never-materialising solver, tiny problem: separate digests of the selected indices and of the errors (is a run-to-run
difference in the last bits of error() -- float64 atomics in the lazy projection's column sums -- or in the selections?)"""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bayesiancoresets_b200 as bc
from conftest import lr_problem
Z, theta = lr_problem(2, 1500, 6, 128)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 128)
d = lambda a: hashlib.sha1(np.asarray(a, dtype=np.float64).tobytes()).hexdigest()[:10]
for alg in ('GIGA', 'OrthoPursuit'):
  cs = bc.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg), materialize=False)
  cs.build(12)
  ev = cs.snnls.last_events
  print('LAZY', alg, 'idx', d([e.f for e in ev]), 'err', d([e.error for e in ev]), 'err[-1] %.17g' % ev[-1].error, flush=True)
