import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, 'bayesian-coresets_b200')
for p in (ROOT, PKG_DIR):
  if p not in sys.path:
    sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
  with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
    return {k: f[k] for k in f.files}


def lr_problem(seed, N, d, S, spread=0.1):
  """Same recipe as oracle/make_golden.py:lr_problem."""
  np.random.seed(seed)
  X = np.random.randn(N, d)
  th_true = np.random.randn(d)
  y = (np.random.rand(N) <= 1./(1.+np.exp(-X.dot(th_true)))).astype(np.float64)
  y[y == 0] = -1.
  Z = y[:, np.newaxis]*X
  theta = th_true + spread*np.random.randn(S, d)
  return Z, theta


@pytest.fixture(scope='session')
def golden():
  return load_golden
