"""OrthoPursuit on a tiny problem: prints what the build did (size, error, failure codes, exact passes).  Used to compare a
normal run with runs under compute-sanitizer, persistent kernel (BCG_OMP_LOOP=1) vs launch per iteration (0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bayesiancoresets_b200 as bc
from conftest import lr_problem
N, S = 1500, 96
Z, theta = lr_problem(1, N, 6, S)
prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
cs.build(S//2 + 20)
ev = cs.snnls.last_events
codes = [(i, e.code, e.f) for i, e in enumerate(ev) if e.code != 0]
print('OMP_LOOP', os.environ.get('BCG_OMP_LOOP', '1'), 'size', cs.snnls.size(), 'error', cs.error(), 'events', len(ev), 'failures', codes[:6],
      'exact', cs.snnls._native.exact_count(), 'first f', [e.f for e in ev[:12]], flush=True)
