"""CPU logic check of the single-block step logic (csrc/step_logic.h compiled for the host by
tests/hostcheck/hostcheck.cpp) against the oracle.  TEST-ONLY: validates the iteration state
machine (winner resolution + float64 re-scoring, reweight, monotone check, retry / latch) on a
box without a GPU; the product library never runs this code on the CPU."""
import ctypes
import os
import subprocess
import tempfile
import numpy as np
import pytest
from conftest import ROOT, load_golden
from oracle import greedy

ALG = {'giga': 0, 'fw': 1, 'omp': 2}
# tolerance of the float32-storage engine against the float64 reference (BASELINE north_star):
# 1e-5 relative to the scale of the weight vector, 1e-5 relative on error()
W_RTOL = 1e-5


def assert_errors_close(err, err_ref, vecs, w_ref):
  """error() = ||A w - b||: 1e-5 relative, plus the a-priori bound of storing A as float32 unit rows
  (every element of A carries a relative rounding of 2^-24, so ||dA w|| <= 2^-24 sum_k w_k ||a_k||)."""
  norms = np.sqrt((vecs**2).sum(axis=1))
  atol = 2.**-24*float(np.abs(w_ref).dot(norms))
  np.testing.assert_allclose(err, err_ref, rtol=W_RTOL, atol=atol)


def assert_weights_close(w, w_ref):
  np.testing.assert_allclose(w, w_ref, rtol=W_RTOL, atol=W_RTOL*np.abs(w_ref).max())


class Event(ctypes.Structure):
  _fields_ = [('code', ctypes.c_int32), ('nact', ctypes.c_int32), ('f', ctypes.c_int64), ('error', ctypes.c_double),
              ('aux0', ctypes.c_double), ('aux1', ctypes.c_double)]


@pytest.fixture(scope='module')
def lib():
  out = os.path.join(tempfile.gettempdir(), 'bcg_hostcheck_%d.so' % os.getuid())
  src = os.path.join(ROOT, 'tests', 'hostcheck', 'hostcheck.cpp')
  subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', out, src])
  return ctypes.CDLL(out)


def device_layout(vecs):
  """what bcg_vecs_from_host_f64 produces: unit float32 rows (ld multiple of 4) + float64 norms"""
  N, S = vecs.shape
  ld = (S + 3)//4*4
  norms = np.sqrt((vecs**2).sum(axis=1))
  An = np.zeros((N, ld), dtype=np.float32)
  An[:, :S] = (vecs/norms[:, None]).astype(np.float32)
  return An, norms, ld


def run_host(lib, vecs, alg, itrs, builds=1, tol=1e-12):
  N, S = vecs.shape
  An, norms, ld = device_layout(vecs)
  b = vecs.sum(axis=0)
  ev = (Event*(itrs*builds))()
  nev, k, halted = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
  err = ctypes.c_double(0)
  idx = np.zeros(itrs*builds + 8, dtype=np.int64)
  w = np.zeros(itrs*builds + 8)
  P = ctypes.c_void_p
  lib.hostcheck_run(ctypes.c_int(ALG[alg]), P(An.ctypes.data), P(norms.ctypes.data), P(b.ctypes.data), ctypes.c_int(S),
                    ctypes.c_int(ld), ctypes.c_int64(N), ctypes.c_double(float(norms.sum())), ctypes.c_int(itrs),
                    ctypes.c_int(builds), ctypes.c_double(tol), ev, ctypes.byref(nev), P(idx.ctypes.data),
                    P(w.ctypes.data), ctypes.byref(k), ctypes.byref(err), ctypes.byref(halted))
  events = [(e.code, e.f, e.error) for e in ev[:nev.value]]
  wd = np.zeros(N)
  wd[idx[:k.value]] = w[:k.value]
  return events, wd, err.value, bool(halted.value)


@pytest.mark.parametrize('alg', ['giga', 'fw'])
def test_step_logic_matches_oracle_c1(lib, alg):
  np.random.seed(1)
  X = np.random.randn(1000, 50)
  o = greedy.ORACLES[alg](X.T, X.sum(axis=0))
  oev = o.build(100)
  ev, w, err, halted = run_host(lib, X, alg, 100)
  assert [e[1] for e in ev] == [e[1] for e in oev]
  assert [e[0] for e in ev] == [e[0] for e in oev]
  assert_errors_close([e[2] for e in ev], [e[2] for e in oev], X, o.w)
  assert_weights_close(w, o.w)
  assert_errors_close(err, o.error(), X, o.w)
  assert not halted


@pytest.mark.parametrize('alg', ['giga', 'fw'])
def test_step_logic_lr_small(lib, alg):
  g = load_golden('lr_small_' + alg)
  vecs = load_golden('lr_project_small')['vecs']
  ev, w, err, halted = run_host(lib, vecs, alg, int(g['itrs']))
  assert [e[1] for e in ev if e[0] == 0] == list(g['sel'])
  assert_weights_close(w, g['w'])
  assert_errors_close(err, float(g['final_error']), vecs, g['w'])


@pytest.mark.parametrize('alg', ['giga', 'fw'])
def test_step_logic_axis_ties_and_latch(lib, alg):
  """X = I_12: exact ties -> lowest index; GIGA then fails twice (cdirnrm < TOL) and latches."""
  X = np.eye(12)
  o = greedy.ORACLES[alg](X.T, X.sum(axis=0))
  oev = o.build(20)
  ev, w, err, halted = run_host(lib, X, alg, 20)
  assert [e[1] for e in ev][:12] == list(range(12))
  assert [(e[0], e[1]) for e in ev] == [(e[0], e[1]) for e in oev]
  assert halted == o.reached_numeric_limit
  np.testing.assert_allclose(w, o.w, rtol=1e-9, atol=1e-12)


def test_step_logic_retry_flag_is_per_build_call(lib):
  """snnls.py:40: `retried_already` is local to build(); build(1) repeated never latches."""
  X = np.eye(6)
  o = greedy.GigaOracle(X.T, X.sum(axis=0))
  for _ in range(10):
    o.build(1)
  ev, w, err, halted = run_host(lib, X, 'giga', 1, builds=10)
  assert not o.reached_numeric_limit and not halted
  assert [(e[0], e[1]) for e in ev] == [(e[0], e[1]) for e in o.events]


def test_omp_select_matches_oracle(lib):
  np.random.seed(3)
  X = np.random.randn(400, 24)
  An, norms, ld = device_layout(X)
  b = X.sum(axis=0)
  o = greedy.OrthoPursuitOracle(X.T, b)
  P = ctypes.c_void_p
  for it in range(20):     # K reaches S = 24 afterwards: residual ~1e-13, scores are rounding noise
    act = np.flatnonzero(o.w > 0).astype(np.int64)
    wa = o.w[act].copy()
    f = ctypes.c_int64(-2)
    lib.hostcheck_omp_select(P(An.ctypes.data), P(norms.ctypes.data), P(b.ctypes.data), ctypes.c_int(24),
                             ctypes.c_int(ld), ctypes.c_int64(400), P(act.ctypes.data), P(wa.ctypes.data),
                             ctypes.c_int(len(act)), ctypes.byref(f))
    fo = int(o.select())
    assert f.value == fo, it
    o.reweight(fo)


# ---------------------------------------------------------------- device NNLS logic vs scipy.optimize.nnls
def run_nnls_sequence(lib, vecs, b, cols, counts, from_scratch=0, downdate=1):
  An, norms, ld = device_layout(vecs)
  N, S = vecs.shape
  cols = np.asarray(cols, dtype=np.int64)
  counts = np.asarray(counts, dtype=np.int32)
  out = np.zeros((len(counts), len(cols)))
  reb = ctypes.c_int(0)
  P = ctypes.c_void_p
  lib.hostcheck_nnls_sequence(P(An.ctypes.data), P(norms.ctypes.data), P(b.ctypes.data), ctypes.c_int(S), ctypes.c_int(ld),
                              ctypes.c_int64(N), P(cols.ctypes.data), ctypes.c_int(len(cols)), P(counts.ctypes.data),
                              ctypes.c_int(len(counts)), ctypes.c_int(from_scratch), ctypes.c_int(downdate),
                              P(out.ctypes.data), ctypes.byref(reb))
  A32 = (An[:, :S].astype(np.float64)*norms[:, None])           # the columns exactly as the device holds them
  return out, A32, reb.value


@pytest.mark.parametrize('seed,shift', [(0, 0.0), (1, 0.5), (2, 2.0)])
def test_nnls_logic_matches_scipy(lib, seed, shift):
  """growing column sets, warm-started (OMP pattern) and from scratch (optimize()); shift > 0 makes the
  columns strongly correlated so that weights hit zero and the step-back / removal path runs"""
  from scipy.optimize import nnls
  rng = np.random.RandomState(seed)
  S, N, K = 40, 300, 28
  vecs = rng.randn(N, S) + shift
  b = vecs[rng.choice(N, 60)].sum(axis=0) + 0.3*rng.randn(S)
  cols = rng.choice(N, K, replace=False)
  counts = list(range(1, K + 1))
  for scratch, downdate in ((0, 1), (1, 1), (0, 0)):
    out, A32, reb = run_nnls_sequence(lib, vecs, b, cols, counts, from_scratch=scratch, downdate=downdate)
    active = np.zeros(K, dtype=bool)
    for t, cnt in enumerate(counts):
      # the problem at step t: columns that were positive after step t-1 plus the new one (orthopursuit.py:38-41)
      active[cnt - 1] = True
      ref = np.zeros(K)
      ref[active] = nnls(A32[cols[active]].T, b, maxiter=100000)[0]
      np.testing.assert_allclose(out[t], ref, rtol=1e-7, atol=1e-9*max(1., np.abs(ref).max()))
      active = ref > 0
    if shift >= 2.0:
      assert (out[-1] == 0).any()          # some weights were driven to zero along the way
      assert reb > 0                       # ... by the removal path (Givens downdate / rebuild)


# ---------------------------------------------------------------- full OMP iteration (omp_iteration) vs the oracle
def run_host_omp(lib, vecs, itrs, builds=1, downdate=1):
  N, S = vecs.shape
  An, norms, ld = device_layout(vecs)
  b = vecs.sum(axis=0)
  ev = (Event*(itrs*builds))()
  nev, k, halted = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
  err = ctypes.c_double(0)
  idx = np.zeros(itrs*builds + 8, dtype=np.int64)
  w = np.zeros(itrs*builds + 8)
  P = ctypes.c_void_p
  lib.hostcheck_run_omp(P(An.ctypes.data), P(norms.ctypes.data), P(b.ctypes.data), ctypes.c_int(S), ctypes.c_int(ld),
                        ctypes.c_int64(N), ctypes.c_int(itrs), ctypes.c_int(builds), ctypes.c_int(downdate), ev, ctypes.byref(nev),
                        P(idx.ctypes.data), P(w.ctypes.data), ctypes.byref(k), ctypes.byref(err), ctypes.byref(halted))
  events = [(e.code, e.f, e.error) for e in ev[:nev.value]]
  wd = np.zeros(N)
  wd[idx[:k.value]] = w[:k.value]
  return events, wd, err.value, bool(halted.value)


def test_omp_iteration_matches_oracle_c1(lib):
  """SURVEY 8c known answers for OrthoPursuit on C1 (first selections, err@10) through the warm-started device
  NNLS logic; stops before K = S, where the residual is rounding noise and selections are not comparable"""
  L = lib
  np.random.seed(1)
  X = np.random.randn(1000, 50)
  o = greedy.OrthoPursuitOracle(X.T, X.sum(axis=0))
  oev = o.build(45)
  ev, w, err, halted = run_host_omp(L, X, 45)
  assert [e[1] for e in ev][:10] == [582, 668, 335, 58, 733, 196, 200, 629, 823, 949]
  assert [(e[0], e[1]) for e in ev] == [(e[0], e[1]) for e in oev]
  assert_errors_close([e[2] for e in ev], [e[2] for e in oev], X, o.w)
  assert_weights_close(w, o.w)
  assert_errors_close(err, o.error(), X, o.w)
  assert not halted


def test_omp_iteration_split_builds_and_small_lr(lib):
  """build(k) repeated == build(n*k) (the direction is prepared per call), on the LR golden projection"""
  L = lib
  g = load_golden('lr_small_omp')
  vecs = load_golden('lr_project_small')['vecs']
  itrs = int(g['itrs'])
  ev, w, err, halted = run_host_omp(L, vecs, itrs)
  assert [e[1] for e in ev if e[0] == 0] == list(g['sel'])
  assert_weights_close(w, g['w'])
  assert_errors_close(err, float(g['final_error']), vecs, g['w'])
  if itrs % 2 == 0:
    ev2, w2, err2, _ = run_host_omp(L, vecs, itrs//2, builds=2)
    assert [e[1] for e in ev2] == [e[1] for e in ev]
    np.testing.assert_allclose(w2, w, rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------- table-driven link functions (softplus_table.h)
def _link(lib, model, lin, y=None):
  lin = np.ascontiguousarray(lin, dtype=np.float64)
  y = np.zeros_like(lin) if y is None else np.ascontiguousarray(y, dtype=np.float64)
  out = np.zeros_like(lin)
  P = ctypes.c_void_p
  lib.hostcheck_link(ctypes.c_int(model), P(lin.ctypes.data), P(y.ctypes.data), ctypes.c_int64(lin.size), P(out.ctypes.data))
  return out


def test_softplus_table_relative_accuracy(lib):
  """g(t) = log1p(exp(t)), t <= 0: the table keeps RELATIVE accuracy over the whole range, tail included"""
  rng = np.random.RandomState(0)
  t = np.concatenate([-rng.rand(200000)*45., -np.arange(0, 38*8 + 1)/8., -np.arange(0, 38*8)/8. - 1e-13,
                      [-0.0, 0.0, -36.999999, -37.0, -37.000001, -700., -745.2]])
  ref = np.log1p(np.exp(t.astype(np.longdouble))).astype(np.float64)
  got = _link(lib, 0, t)
  ok = ref > 0
  assert np.max(np.abs(got[ok]/ref[ok] - 1.)) < 1e-15
  assert np.all(got[~ok] == 0.)
  assert np.isnan(_link(lib, 0, np.array([np.nan]))[0])


def test_fast_links_match_reference_formulas(lib):
  """LR and Poisson links against the oracle's NumPy restatement of model_lr.py:25-32 / model_poiss.py:25-38"""
  from oracle import models
  rng = np.random.RandomState(1)
  lin = np.concatenate([rng.randn(100000)*5., rng.randn(20000)*60., [0., -0., 99.9, 100., 100.1, -99.9, -100., -100.1,
                                                                   -36.9, -37.1, 37., 700., -700., -120.]])
  # LR: loglik of a 1-d datapoint z = 1 against samples theta = lin  (m = -lin)
  ref = models.lr_loglik(np.ones((1, 1)), lin[:, None])[0]
  got = _link(lib, 1, lin)
  np.testing.assert_allclose(got, ref, rtol=2e-15, atol=0.)
  # Poisson: y s - exp(s) (the row-constant gammaln(y + 1) is added back for the comparison)
  from scipy.special import gammaln
  y = rng.poisson(3., size=lin.size).astype(np.float64)
  Z = np.array([[1., 0.]])
  for yy in (0., 1., 7.):
    Z[0, 1] = yy
    ref = models.poisson_loglik(Z, lin[:, None])[0] + gammaln(yy + 1.)
    got = _link(lib, 2, lin, np.full(lin.size, yy))
    # y s and exp(s) cancel near the mode: absolute rounding of O(10) terms
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=5e-14)


def test_branch_free_links_match_reference_formulas(lib):
  """the *_nb forms used inside the projection kernels (no tail branch; the two highest Taylor orders in float32): exact to
  2e-15 wherever link_needs_tail() is false (|lin| <= 37)"""
  from oracle import models
  from scipy.special import gammaln
  rng = np.random.RandomState(2)
  lin = np.concatenate([rng.randn(200000)*6., rng.uniform(-37., 37., size=100000), [0., -0., 36.99, -36.99, 37., -37., 1e-300, -1e-9]])
  lin = lin[np.abs(lin) <= 37.]
  ref = models.lr_loglik(np.ones((1, 1)), lin[:, None])[0]
  np.testing.assert_allclose(_link(lib, 3, lin), ref, rtol=2e-15, atol=0.)
  Z = np.array([[1., 0.]])
  for yy in (0., 2., 9.):
    Z[0, 1] = yy
    ref = models.poisson_loglik(Z, lin[:, None])[0] + gammaln(yy + 1.)
    np.testing.assert_allclose(_link(lib, 4, lin, np.full(lin.size, yy)), ref, rtol=1e-13, atol=5e-14)


def test_nnls_givens_removal_stress(lib):
  """many removals (more candidate columns than dimensions, strongly correlated columns): the Givens column
  removal keeps the warm-started factorisation consistent with scipy.optimize.nnls over a long sequence"""
  from scipy.optimize import nnls
  rng = np.random.RandomState(7)
  S, N, K = 48, 400, 110
  vecs = rng.randn(N, S) + 1.5
  b = vecs[rng.choice(N, 80)].sum(axis=0) + 0.5*rng.randn(S)
  cols = rng.choice(N, K, replace=False)
  counts = list(range(1, K + 1))
  out, A32, reb = run_nnls_sequence(lib, vecs, b, cols, counts, from_scratch=0, downdate=1)
  assert reb >= 20
  active = np.zeros(K, dtype=bool)
  for t, cnt in enumerate(counts):
    active[cnt - 1] = True
    ref = np.zeros(K)
    ref[active] = nnls(A32[cols[active]].T, b, maxiter=100000)[0]
    res_ref = np.linalg.norm(A32[cols].T.dot(ref) - b)
    res_got = np.linalg.norm(A32[cols].T.dot(out[t]) - b)
    assert np.all(out[t] >= 0) and np.all(out[t][~active] == 0)
    assert res_got <= res_ref*(1. + 1e-9) + 1e-9
    np.testing.assert_allclose(out[t], ref, rtol=1e-6, atol=1e-7*max(1., np.abs(ref).max()))
    active = ref > 0


def test_nnls_dependent_columns_keep_the_optimal_residual(lib):
  """duplicated columns (two data rows with the same vector): the dependent copy is refused by the Gram-Schmidt append
  and stays at zero; the minimiser is then not unique, so the comparison with scipy.optimize.nnls is on the residual and
  on feasibility, for the Givens-removal and the rebuild variant alike"""
  from scipy.optimize import nnls
  rng = np.random.RandomState(3)
  S, N, K = 24, 120, 40
  vecs = rng.randn(N, S) + 0.8
  vecs[1::3] = vecs[0::3][:vecs[1::3].shape[0]]          # every third row duplicates its predecessor's source row
  b = vecs[rng.choice(N, 30)].sum(axis=0) + 0.2*rng.randn(S)
  cols = rng.permutation(N)[:K]
  counts = list(range(1, K + 1))
  for downdate in (1, 0):
    out, A32, reb = run_nnls_sequence(lib, vecs, b, cols, counts, from_scratch=0, downdate=downdate)
    active = np.zeros(K, dtype=bool)
    for t, cnt in enumerate(counts):
      active[cnt - 1] = True
      ref = np.zeros(K)
      ref[active] = nnls(A32[cols[active]].T, b, maxiter=100000)[0]
      res_ref = np.linalg.norm(A32[cols].T.dot(ref) - b)
      res_got = np.linalg.norm(A32[cols].T.dot(out[t]) - b)
      assert np.all(out[t] >= 0) and np.all(out[t][~active] == 0)
      assert res_got <= res_ref*(1. + 1e-8) + 1e-9, (downdate, t, res_got, res_ref)
      active = out[t] > 0                                  # continue from OUR active set (what the device loop does)


# ---- float16 pre-filter of the persistent scan: the proof obligation of csrc/filter_bounds.h ---------------------------
def _filter_case(lib, An, d0, d1, giga, order):
  n, ld = An.shape
  S = d0.shape[0]
  sc, lb, ub = (np.empty(n, np.float32) for _ in range(3))
  f32p = ctypes.POINTER(ctypes.c_float)
  lib.hostcheck_filter_bounds(An.ctypes.data_as(f32p), ctypes.c_int64(n), ld, S, d0.ctypes.data_as(f32p),
                              d1.ctypes.data_as(f32p), giga, order, sc.ctypes.data_as(f32p), lb.ctypes.data_as(f32p),
                              ub.ctypes.data_as(f32p))
  return sc, lb, ub


@pytest.mark.parametrize('S', [130, 256, 500, 512])
@pytest.mark.parametrize('giga', [0, 1])
def test_filter16_bounds_contain_the_float32_score(lib, S, giga):
  """every float32 score lies inside [lb, ub] computed from the float16-rounded row, for random rows, rows (anti)parallel
  to the directions, sparse rows, rows of tiny (float16-subnormal) entries and zero rows; both summation orders"""
  rng = np.random.RandomState(S + giga)
  ld = (S + 3)//4*4
  n = 4000
  X = rng.randn(n, S)
  X[100:200] *= rng.rand(100, S) < 0.02                     # sparse rows
  X[200:300] = rng.randn(100, 1)*np.ones(S) + 1e-3*rng.randn(100, S)   # nearly constant rows
  X[300:400] *= 10.**rng.uniform(-8, 0, size=(100, S))      # wide dynamic range: many float16 subnormals after scaling
  d0 = rng.randn(S); d0 /= np.linalg.norm(d0)
  d1 = rng.randn(S); d1 -= d1.dot(d0)*d0; d1 /= np.linalg.norm(d1)
  X[400:450] = d1 + 10.**rng.uniform(-7, -1, size=(50, 1))*rng.randn(50, S)    # nearly parallel to the iterate (den -> 0)
  X[450:500] = -d1 + 10.**rng.uniform(-7, -1, size=(50, 1))*rng.randn(50, S)
  X[500:550] = d0 + 10.**rng.uniform(-7, -1, size=(50, 1))*rng.randn(50, S)
  X[550] = d1; X[551] = -d1; X[552] = d0
  nrm = np.linalg.norm(X, axis=1)
  nrm[nrm == 0] = 1.
  An = np.zeros((n, ld), np.float32)
  An[:, :S] = (X/nrm[:, None]).astype(np.float32)
  An[600:610] = 0.                                          # zero rows
  d0f, d1f = d0.astype(np.float32), d1.astype(np.float32)
  for order in (0, 1):
    sc, lb, ub = _filter_case(lib, An, d0f, d1f, giga, order)
    assert np.all(lb <= sc) and np.all(sc <= ub), (np.flatnonzero(~((lb <= sc) & (sc <= ub)))[:10])
    # the bound is tight where it is finite: a few 1e-4 around the score for well-conditioned rows
    fin = np.isfinite(lb) & np.isfinite(ub)
    assert fin[:100].all()
    assert np.max((ub - lb)[:100]) < 4e-3
    # the filter's decision, as the kernel takes it: L = the largest lower bound; every row that can matter to the selection
    # (the float32 arg-max and everything inside its near-tie window) has an upper bound that reaches filter_threshold(L)
    fin_lb = lb[np.isfinite(lb)]
    L = fin_lb.max() if fin_lb.size else -np.inf
    thr = lib_threshold(lib, L)
    top = sc.max()
    window = np.float32(2e-5) + np.float32(1e-5)*abs(top)
    must = sc >= top - window
    assert np.all(ub[must] >= thr)
    assert L <= top
  # the threshold keeps every row inside the near-tie window of ANY maximum >= L
  for L in (-0.5, -1e-3, 0., 1e-3, 0.2, 1., 50.):
    thr = lib_threshold(lib, L)
    for top in (L, L + 1e-6, L + 0.5*abs(L) + 1e-3, L + 10.):
      assert thr <= np.float32(top) - (np.float32(2e-5) + np.float32(1e-5)*abs(np.float32(top)))
  assert lib_threshold(lib, np.inf) >= 3e38 and lib_threshold(lib, -np.inf) == -np.inf


def lib_threshold(lib, L):
  lib.hostcheck_filter_threshold.restype = ctypes.c_float
  return lib.hostcheck_filter_threshold(ctypes.c_float(L))


@pytest.mark.parametrize('S', [256, 512])
def test_filter16_bounds_hold_in_the_aligned_worst_case(lib, S):
  """the adversarial row for the Cauchy-Schwarz step of the bound: |a_i| proportional to |d_i|, every element on a float16
  rounding MIDPOINT with an even lower neighbour (round-to-nearest-even moves all of them the same way, by the full half
  ulp), signs aligned with the direction -- the quantisation errors add up coherently to ~2^-11 |a||d|, 90 % of the bound"""
  rng = np.random.RandomState(S)
  ld = S
  d0 = rng.randn(S); d0 /= np.linalg.norm(d0)
  d1 = rng.randn(S); d1 -= d1.dot(d0)*d0; d1 /= np.linalg.norm(d1)
  rows = []
  for sgn in (1., -1.):
    for dirv in (d0, d1):
      m = 0.999*np.abs(dirv)
      e = np.floor(np.log2(np.maximum(m, 2.**-14)))          # float16 exponent (normal range)
      ulp = 2.**(e - 10)
      k = np.floor(m/ulp)
      k -= (k % 2)                                           # even lower neighbour: the tie rounds DOWN
      a = (k + 0.5)*ulp                                      # exactly representable in float32
      a[m < 2.**-14] = 0.
      rows.append(sgn*np.sign(dirv)*a)
  An = np.ascontiguousarray(np.array(rows), dtype=np.float32)
  assert np.array_equal(An.astype(np.float64), np.array(rows))                 # no float32 rounding of the construction
  assert np.all(np.linalg.norm(An.astype(np.float64), axis=1) <= 1.0001)
  d0f, d1f = d0.astype(np.float32), d1.astype(np.float32)
  worst, half = 0., 0.
  for giga in (0, 1):
    for order in (0, 1):
      sc, lb, ub = _filter_case(lib, An, d0f, d1f, giga, order)
      assert np.all(lb <= sc) and np.all(sc <= ub), (giga, order, sc, lb, ub)
      if not giga:
        mid = 0.5*(lb.astype(np.float64) + ub)                # = the float16 inner product t
        worst = max(worst, float(np.max(np.abs(mid - sc))))
        half = 0.5*float((ub - lb)[0])                        # = E, the bound on |t - s|
  # the construction really stresses the bound: a half ulp is 2^-11 of the value only at the bottom of a binade, ~0.7 of
  # that on average, so the coherent sum reaches ~0.6 E
  assert 0.5*half < worst <= half, (worst, half)


def test_filter16_slot_policy_simulation(lib):
  """the bookkeeping of the pre-filter, simulated: stages arrive at the warps in arbitrary order, the shared lower bound L is
  seen with arbitrary delay, every warp has a handful of slots which it compacts against its current L.  Invariant: a warp
  either reports overflow or ends with every stage that holds a row inside the near-tie window of the true maximum still in
  its slots -- a dropped entry was dropped against a threshold that only rises (loop_kernel.cuh: scan_cta_body, pass A / B)."""
  rng = np.random.RandomState(11)
  S, n, rows_per_stage, cap = 256, 6000, 8, 4
  d0 = rng.randn(S); d0 /= np.linalg.norm(d0)
  X = rng.randn(n, S)
  X[::97] = d0 + 0.3*rng.randn(len(X[::97]), S)              # a population of high scorers, some of them close together
  X[1000:1003] = X[1000]                                     # exact duplicates of one of them
  An = (X/np.linalg.norm(X, axis=1)[:, None]).astype(np.float32)
  d0f = d0.astype(np.float32)
  sc, lb, ub = _filter_case(lib, An, d0f, d0f, 0, 1)
  top = sc.max()
  must_rows = sc >= top - (np.float32(2e-5) + np.float32(1e-5)*abs(top))
  nst = n//rows_per_stage
  st_ub = ub.reshape(nst, rows_per_stage).max(1)
  st_lb = lb.reshape(nst, rows_per_stage).max(1)
  st_must = must_rows.reshape(nst, rows_per_stage).any(1)
  for trial in range(20):
    W = int(rng.choice([1, 3, 16, 64]))
    owner = rng.randint(0, W, size=nst)
    order = rng.permutation(nst)                             # global arrival order of the stages
    Lglob = -np.inf                                          # the atomicMax word
    Lw = np.full(W, -np.inf); Lseen = np.full(W, -np.inf)
    slots = [[] for _ in range(W)]
    overflow = np.zeros(W, bool)
    for s_ in order:
      w = owner[s_]
      Lw[w] = max(Lw[w], st_lb[s_])
      if rng.rand() < 0.3:                                   # a delayed look at the shared word
        Lseen[w] = Lglob
      Lw[w] = max(Lw[w], Lseen[w])
      Lglob = max(Lglob, Lw[w])                              # publish
      thr = lib_threshold(lib, Lw[w])
      if st_ub[s_] >= thr:
        if len(slots[w]) == cap:
          slots[w] = [e for e in slots[w] if st_ub[e] >= thr]
        if len(slots[w]) < cap:
          slots[w].append(s_)
        else:
          overflow[w] = True
    for w in range(W):
      thr = lib_threshold(lib, max(Lw[w], Lglob if rng.rand() < 0.5 else -np.inf))
      kept = {e for e in slots[w] if st_ub[e] >= thr}
      mine_must = {int(e) for e in np.flatnonzero(st_must & (owner == w))}
      assert overflow[w] or mine_must <= kept, (trial, w)
    assert Lglob <= top
