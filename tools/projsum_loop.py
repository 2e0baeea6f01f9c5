"""diagnostic: repeated column-sum projections with SM clock / power sampled between calls"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
import numpy as np
import bayesiancoresets_b200 as bc
model, N, d, S, reps = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
rng = np.random.RandomState(0)
X = rng.randn(N, d)
th = rng.randn(S, d)/np.sqrt(d)
if model == 'lr':
  prj, data = bc.LogisticRegressionProjector(lambda n, w, p: th, S), X
elif model == 'poisson':
  prj, data = bc.PoissonProjector(lambda n, w, p: th, S), np.hstack((X, rng.poisson(1., (N, 1)).astype(float)))
else:
  prj, data = bc.GaussianProjector(lambda n, w, p: th, S, np.eye(d)), X
prj.project_sum(data)
ts = []
for i in range(reps):
  t0 = time.perf_counter(); prj.project_sum(data); ts.append(1e3*(time.perf_counter() - t0))
q = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown', '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
print(model, os.environ.get('BCG_PROJSUM_MMA', '1'), ' '.join('%.1f' % t for t in ts), '|', q, flush=True)
