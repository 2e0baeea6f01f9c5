"""The sparse replay oracle (oracle/replay.py) against the dense oracle (oracle/greedy.py, pinned bit-for-bit to the
reference): identical selection sequences and event codes, weights and errors to float64 round-off.  The replay is
what the GPU suite uses at N = 1e7, S = 512 together with the independent float64 audit kernel."""
import numpy as np
import pytest
from conftest import lr_problem
from oracle import greedy, models, replay


def run_pair(vecs, alg, itrs):
  b = vecs.sum(axis=0)
  o = greedy.ORACLES[alg](vecs.T, b)
  oev = o.build(itrs)
  kw = {'norm_sum': np.sqrt((vecs**2).sum(axis=1)).sum()} if alg == 'fw' else {}
  r = replay.REPLAYS[alg](vecs.shape[0], b, replay.dense_score_fn(vecs), lambda idx: vecs[idx], **kw)
  rev = r.build(itrs)
  return o, oev, r, rev


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_replay_equals_dense_oracle_c1(alg):
  np.random.seed(1)
  X = np.random.randn(1000, 50)
  o, oev, r, rev = run_pair(X, alg, 100)
  assert [(e[0], e[1]) for e in rev] == [(e[0], e[1]) for e in oev]
  np.testing.assert_allclose(r.w, o.w, rtol=1e-9, atol=1e-12*np.abs(o.w).max())
  np.testing.assert_allclose([e[2] for e in rev], [e[2] for e in oev], rtol=1e-8, atol=1e-11)
  assert len(r.diag) >= 1 and all(g[1] >= 0 and g[2] >= 1 for g in r.diag)


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_replay_equals_dense_oracle_lr(alg):
  Z, theta = lr_problem(4, 5000, 6, 64)
  vecs = models.project(models.lr_loglik, Z, theta)
  o, oev, r, rev = run_pair(vecs, alg, 50)
  assert [(e[0], e[1]) for e in rev] == [(e[0], e[1]) for e in oev]
  np.testing.assert_allclose(r.w, o.w, rtol=1e-8, atol=1e-11*np.abs(o.w).max())


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_replay_failure_path_axis(alg):
  """X = I: every score ties (lowest index wins), then the solvers hit the numeric limit exactly as the reference"""
  X = np.eye(12)
  o, oev, r, rev = run_pair(X, alg, 20)
  assert [(e[0], e[1]) for e in rev] == [(e[0], e[1]) for e in oev]
  assert r.reached_numeric_limit == o.reached_numeric_limit
  np.testing.assert_allclose(r.w, o.w, rtol=1e-12, atol=1e-14)
