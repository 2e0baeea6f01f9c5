"""build(25) of the headline workload under different ring geometries of the persistent kernel (env read at solver creation)
  python tools/filter_sweep.py            -> runs itself once per configuration in a subprocess
"""
import os, sys, json, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'one':
  sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200')); sys.path.insert(0, ROOT)
  os.environ['BCG_BENCH_NO_CLOCKS'] = '1'
  import numpy as np
  import bayesiancoresets_b200 as bc
  import bench
  N, d, S = int(float(os.environ.get('SWEEP_N', '1e7'))), 10, int(os.environ.get('SWEEP_S', '512'))
  alg = {'giga': bc.snnls.GIGA, 'fw': bc.snnls.FrankWolfe, 'omp': bc.snnls.OrthoPursuit}[os.environ.get('SWEEP_ALG', 'giga')]
  ctx = bc.Context.default(0)
  Z, th = bench.lr_shard(0, 0, N, d)
  theta = bench.lr_samples(0, th, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S, ctx=ctx)
  cs = bc.HilbertCoreset(Z, prj, snnls=alg)
  nat = cs.snnls._native
  cs.snnls.build(5)
  ms = []
  for _ in range(3):
    cs.snnls.build(20); ms.append(nat.timing()['build_ms'])
  on, rows = nat.filter16_stats()
  print(json.dumps({'cfg': os.environ.get('SWEEP_CFG'), 'ms_per_iter': round(min(ms)/20, 4), 'all': [round(m/20, 4) for m in ms],
                    'filter16': on, 'rescanned_rows_per_iter': rows/65., 'sel_hash': bench.sel_hash(cs.snnls.last_events)[:8]}), flush=True)
  sys.exit(0)
cfgs = [
  {'BCG_FILTER16': '0'},
  {},
  {'BCG_SCAN_WARPS': '10'},
  {'BCG_SCAN_WARPS': '11'},
  {'BCG_SCAN_WARPS': '7', 'BCG_SCAN_STAGES': '3'},
  {'BCG_SCAN_WARPS': '6', 'BCG_SCAN_STAGE_BYTES': '16384'},
  {'BCG_SCAN_WARPS': '5', 'BCG_SCAN_STAGE_BYTES': '16384'},
  {'BCG_SCAN_WARPS': '11', 'BCG_SCAN_EVICT_FIRST': '1'},
]
for c in cfgs:
  env = dict(os.environ); env.update(c); env['SWEEP_CFG'] = json.dumps(c)
  subprocess.call([sys.executable, os.path.abspath(__file__), 'one'], env=env)
