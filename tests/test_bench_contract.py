"""CPU checks of bench.py: the reference arm (the unmodified reference from baseline/_ref when it is installed, else the
oracle port, timed on the host cores) prints ONE JSON line with the
contract's keys, and the synthetic workload generator gives every shard the same rows as the unsharded problem."""
import json
import os
import subprocess
import sys
import numpy as np
from conftest import ROOT

sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
  env = dict(os.environ, OMP_NUM_THREADS='2')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'lr_giga_N2e5_S256',
                        '--steps', '3', '--warmup', '3'], capture_output=True, text=True, timeout=600, env=env)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['metric'] == 'greedy_iters_per_sec' and d['unit'] == 'iters/s'
  assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['dtype'] == 'f64' and d['data'] == 'synthetic'
  assert d['config']['workload'] == 'lr_giga_N2e5_S256' and d['steps'] == 3 and d['value'] > 0
  assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['cores'] >= 1
  assert d['e2e'] == {'value': d['value'], 'unit': 'iters/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert abs(d['ms_per_step']*d['value'] - 1e3) < 1e-6


def test_reference_arm_is_rank0_only_under_torchrun():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '2'],
                       capture_output=True, text=True, timeout=120, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ''


def test_synthetic_shards_tile_the_full_problem():
  import bench
  Z, th = bench.lr_shard(0, 0, 2_500_000, 4)
  for lo, hi in ((0, 7), (999_990, 1_000_020), (1_700_000, 2_500_000)):
    Zs, ths = bench.lr_shard(0, lo, hi, 4)
    assert np.array_equal(Zs, Z[lo:hi]) and np.array_equal(ths, th)
  assert bench.lr_samples(0, th, 8).shape == (8, 4)
