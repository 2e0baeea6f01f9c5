"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C-ABI
(ctypes -> libbcg_b200.so); the checker is the oracle / the reference-generated golden fixtures.

Tolerances (BASELINE north_star): selected indices bit-exact (ties absent), weights within 1e-5
relative to the scale of the weight vector, error() within 1e-5 relative plus the a-priori bound
of float32 storage of the unit rows (2^-24 sum_k w_k ||a_k||)."""
import numpy as np
import pytest
from conftest import load_golden, lr_problem
from oracle import greedy, models

pytestmark = pytest.mark.gpu
W_RTOL = 1e-5


@pytest.fixture(scope='module')
def bc():
  import bayesiancoresets_b200 as bc
  bc.Context.default()        # raises when there is no GPU: no silent fallback
  return bc


class IDProjector(object):
  def update(self, wts, pts):
    pass

  def project(self, pts, grad=False):
    return pts


def assert_weights_close(w, w_ref):
  np.testing.assert_allclose(w, w_ref, rtol=W_RTOL, atol=W_RTOL*np.abs(w_ref).max())


def assert_errors_close(err, err_ref, vecs, w_ref):
  norms = np.sqrt((vecs**2).sum(axis=1))
  atol = 2.**-24*float(np.abs(w_ref).dot(norms)) + 1e-300
  np.testing.assert_allclose(err, err_ref, rtol=W_RTOL, atol=atol)


def algs(bc):
  return {'giga': bc.snnls.GIGA, 'fw': bc.snnls.FrankWolfe, 'omp': bc.snnls.OrthoPursuit}


# ---------------------------------------------------------------- matrix construction
@pytest.mark.parametrize('shape', [(1000, 50), (257, 3), (64, 512), (33, 1000), (1, 7)])
def test_from_host_roundtrip(bc, shape):
  rng = np.random.RandomState(0)
  X = rng.randn(*shape)*rng.uniform(0.1, 30., size=(shape[0], 1))
  v = bc.DeviceVecs.from_host(X)
  assert v.shape == shape
  np.testing.assert_allclose(v.norms(), np.sqrt((X**2).sum(axis=1)), rtol=1e-14)
  np.testing.assert_allclose(v.sum(axis=0), X.sum(axis=0), rtol=1e-12, atol=1e-12*np.abs(X).sum(axis=0).max())
  np.testing.assert_allclose(v.to_numpy(), X, rtol=0, atol=2.**-23*np.abs(X).max(axis=1, keepdims=True).max())
  assert v.norm_sum() == pytest.approx(np.sqrt((X**2).sum(axis=1)).sum(), rel=1e-13)
  assert v.zero_rows() == 0


def test_from_host_strided_and_zero_rows(bc):
  rng = np.random.RandomState(1)
  big = rng.randn(100, 40)
  X = big[:, :24]                      # row stride 40, S = 24
  X[5] = 0.
  v = bc.DeviceVecs.from_host(X)
  assert v.zero_rows() == 1
  np.testing.assert_allclose(v.sum(axis=0), X.sum(axis=0), rtol=1e-12, atol=1e-13)
  with pytest.raises(ValueError):
    bc.snnls.GIGA(X.T, X.sum(axis=0))
  Y = rng.randn(10, 4)
  with pytest.raises(bc.util.NumericalPrecisionError):
    bc.snnls.GIGA(Y.T, np.zeros(4))


def check_projection(v, ref):
  got = v.to_numpy()
  scale = np.sqrt((ref**2).sum(axis=1, keepdims=True))
  # unit rows are float32: absolute error <= 2^-24 * row norm (+ float64 evaluation noise)
  assert np.max(np.abs(got - ref)/scale) < 2.**-23
  np.testing.assert_allclose(v.norms(), scale[:, 0], rtol=1e-9)
  np.testing.assert_allclose(v.sum(axis=0), ref.sum(axis=0), rtol=1e-9, atol=1e-9*np.abs(ref).sum(axis=0).max())


def test_project_lr_golden(bc):
  g = load_golden('lr_project_small')
  check_projection(bc.DeviceVecs.project_lr(g['Z'], g['theta']), g['vecs'])
  g = load_golden('lr_project_saturated')
  check_projection(bc.DeviceVecs.project_lr(g['Z'], g['theta']), g['vecs'])


def test_project_gaussian_poisson_golden(bc):
  g = load_golden('gaussian_project_small')
  check_projection(bc.DeviceVecs.project_gaussian(g['x'], g['theta'], g['Siginv']), g['vecs'])
  for name in ('poisson_project_small', 'poisson_project_extreme'):
    g = load_golden(name)
    check_projection(bc.DeviceVecs.project_poisson(g['Z'], g['theta']), g['vecs'])


@pytest.mark.parametrize('d,S', [(40, 200), (130, 96), (7, 1000)])
def test_project_shapes_vs_oracle(bc, d, S):
  rng = np.random.RandomState(d)
  Z = rng.randn(300, d)
  th = rng.randn(S, d)/np.sqrt(d)
  check_projection(bc.DeviceVecs.project_lr(Z, th), models.project(models.lr_loglik, Z, th))
  Siginv = np.eye(d) + 0.1*np.ones((d, d))
  f = lambda x, t: models.gaussian_loglik(x, t, Siginv, 0.)
  check_projection(bc.DeviceVecs.project_gaussian(Z, th, Siginv), models.project(f, Z, th))


# ---------------------------------------------------------------- greedy loop vs golden / oracle
def run_gpu(bc, vecs, alg, itrs):
  cs = bc.HilbertCoreset(vecs, IDProjector(), snnls=algs(bc)[alg])
  cs.build(itrs)
  ev = cs.snnls.last_events
  return cs, ev


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_c1_normal_golden(bc, alg):
  g = load_golden('c1_normal_' + alg)
  np.random.seed(int(g['seed']))
  X = np.random.randn(int(g['N']), int(g['S']))
  cs, ev = run_gpu(bc, X, alg, int(g['itrs']))
  itok = 100 if alg != 'omp' else 45       # OMP: after K ~ S the residual is rounding noise
  assert [e.f for e in ev][:itok] == list(g['sel'][:itok])
  assert all(e.code == 0 for e in ev)
  if alg != 'omp':
    assert_errors_close([e.error for e in ev], g['errs'], X, g['w'])
    assert_weights_close(cs.snnls.weights(), g['w'])
    assert cs.snnls.size() == int(g['size'])
    wts, pts, idcs = cs.get()
    assert np.array_equal(idcs, np.flatnonzero(g['w'] > 0))
    assert np.array_equal(pts, X[idcs])
  else:
    assert cs.error() < 1e-6*np.sqrt((X.sum(axis=0)**2).sum())
    assert cs.snnls.size() <= 50
    # weights and errors where they are meaningful: the oracle (= the reference) stopped after the same 45 iterations
    o = greedy.OrthoPursuitOracle(X.T, X.sum(axis=0))
    oev = o.build(itok)
    c2, ev2 = run_gpu(bc, X, alg, itok)
    assert [e.f for e in ev2] == [e[1] for e in oev]
    assert_weights_close(c2.snnls.weights(), o.w)
    assert_errors_close([e.error for e in ev2], [e[2] for e in oev], X, o.w)


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_axis_ties_lowest_index_and_latch(bc, alg):
  g = load_golden('axis12_' + alg)
  X = np.eye(12)
  cs, ev = run_gpu(bc, X, alg, int(g['itrs']))
  assert [e.f for e in ev][:12] == list(range(12))
  assert_weights_close(cs.snnls.weights(), g['w'])
  if alg == 'giga':
    assert [e.code for e in ev[12:]] == [1, 1]          # cdirnrm < TOL twice -> latch
    assert cs.snnls.reached_numeric_limit and cs.reached_numeric_limit is False
    n_before = len(cs.snnls.last_events)
    cs.snnls.build(5)                                    # latched: returns immediately
    assert len(cs.snnls.last_events) == n_before


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_lr_small_golden(bc, alg):
  g = load_golden('lr_small_' + alg)
  p = load_golden('lr_project_small')
  prj = bc.LogisticRegressionProjector(lambda n, w, pts: p['theta'], int(p['S']))
  cs = bc.HilbertCoreset(p['Z'], prj, snnls=algs(bc)[alg])
  cs.build(int(g['itrs']))
  ev = cs.snnls.last_events
  nsel = len(g['sel']) if alg != 'omp' else 60
  assert [e.f for e in ev if e.code == 0][:nsel] == list(g['sel'][:nsel])
  if alg != 'omp':
    assert_weights_close(cs.snnls.weights(), g['w'])
    assert_errors_close(cs.error(), float(g['final_error']), p['vecs'], g['w'])
  else:
    # OMP: past K ~ S the residual is rounding noise and so are the selections; the WEIGHTS and the per-iteration errors are
    # compared where they are meaningful -- against the oracle (bit-identical to the reference) stopped after 60 iterations
    o = greedy.OrthoPursuitOracle(p['vecs'].T, p['vecs'].sum(axis=0))
    oev = o.build(nsel)
    c2 = bc.HilbertCoreset(p['Z'], prj, snnls=algs(bc)[alg])
    c2.build(nsel)
    assert [e.f for e in c2.snnls.last_events] == [e[1] for e in oev] == list(g['sel'][:nsel])
    assert_weights_close(c2.snnls.weights(), o.w)
    assert_errors_close([e.error for e in c2.snnls.last_events], [e[2] for e in oev], p['vecs'], o.w)


def test_incremental_build_and_retry_flag(bc):
  """build(k) is incremental; `retried_already` is local to each build() call (snnls.py:40)."""
  np.random.seed(1)
  X = np.random.randn(1000, 50)
  a, _ = run_gpu(bc, X, 'giga', 60)
  b = bc.HilbertCoreset(X, IDProjector())
  for k in (1, 9, 20, 30):
    b.build(k)
  assert_weights_close(b.snnls.weights(), a.snnls.weights())
  e = bc.HilbertCoreset(np.eye(6), IDProjector())
  for _ in range(10):
    e.build(1)
  assert not e.snnls.reached_numeric_limit
  e.reset()
  assert e.snnls.size() == 0 and e.error() == pytest.approx(np.sqrt(6.))


def test_optimize_never_worse(bc):
  np.random.seed(2)
  X = np.random.randn(2000, 40)
  cs = bc.HilbertCoreset(X, IDProjector(), snnls=bc.snnls.FrankWolfe)
  cs.build(30)
  e0 = cs.error()
  o = greedy.FrankWolfeOracle(X.T, X.sum(axis=0))
  o.build(30)
  o.optimize()
  cs.optimize()
  assert cs.error() <= e0*(1 + 1e-12)
  assert cs.error() == pytest.approx(o.error(), rel=1e-5)


@pytest.mark.parametrize('alg,N,S,itrs', [('giga', 200000, 256, 40), ('fw', 100000, 512, 40), ('giga', 50001, 100, 60)])
def test_medium_vs_oracle(bc, alg, N, S, itrs):
  Z, theta = lr_problem(11, N, 8, S)
  vecs = models.project(models.lr_loglik, Z, theta)
  o = greedy.ORACLES[alg](vecs.T, vecs.sum(axis=0))
  oev = o.build(itrs)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  cs = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  cs.build(itrs)
  ev = cs.snnls.last_events
  assert [e.f for e in ev] == [e[1] for e in oev]
  assert_weights_close(cs.snnls.weights(), o.w)
  assert_errors_close([e.error for e in ev], [e[2] for e in oev], vecs, o.w)


def test_full_size_properties(bc):
  """BASELINE configs[1] size (N=1e6, S=256): size-independent properties instead of the oracle."""
  N, S, itrs = 1000000, 256, 60
  Z, theta = lr_problem(0, N, 10, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  cs = bc.HilbertCoreset(Z, prj)
  bnorm = cs.error()
  cs.build(itrs)
  ev = cs.snnls.last_events
  errs = np.array([e.error for e in ev])
  assert len(ev) == itrs and all(e.code == 0 for e in ev)
  assert errs[0] < bnorm and np.all(np.diff(errs) <= 0)            # monotone (snnls.py:56-61)
  wts, pts, idcs = cs.get()
  assert wts.shape[0] <= itrs and np.all(wts > 0)
  assert np.all(np.diff(idcs) > 0) and idcs.min() >= 0 and idcs.max() < N
  assert np.array_equal(pts, Z[idcs])
  # error() must equal ||A w - b|| recomputed on the host in float64 from the selected points only
  sub = models.project(models.lr_loglik, Z[idcs], theta)
  b = cs.snnls.b
  assert cs.error() == pytest.approx(np.sqrt(((wts.dot(sub) - b)**2).sum()), rel=1e-5)
  # the first selection maximises cos(a_n, b): check against a float64 host scan over a slab
  f0 = ev[0].f
  slab = models.project(models.lr_loglik, Z[:50000], theta)
  cosb = slab.dot(b)/np.sqrt((slab**2).sum(axis=1))/np.sqrt((b**2).sum())
  row_f0 = models.project(models.lr_loglik, Z[f0:f0+1], theta)[0]
  cos_f0 = row_f0.dot(b)/np.sqrt((row_f0**2).sum())/np.sqrt((b**2).sum())
  assert cos_f0 >= cosb.max() - 1e-12


def test_empty_and_tiny_inputs(bc):
  s = bc.snnls.GIGA(np.zeros((5, 0)), np.ones(5))
  s.build(3)
  assert s.size() == 0 and s.weights().shape == (0,)
  X = np.array([[3., 4.]])
  cs = bc.HilbertCoreset(X, IDProjector())
  cs.build(3)
  wts, pts, idcs = cs.get()
  assert list(idcs) == [0] and wts[0] == pytest.approx(1., rel=1e-6) and cs.error() < 1e-5


# ---------------------------------------------------------------- projector API, SparseVI, BatchPSVI
def test_projector_project_rows_sum_and_argmax(bc):
  g = load_golden('lr_project_small')
  prj = bc.LogisticRegressionProjector(lambda n, w, p: g['theta'], int(g['S']))
  rows = prj.project(g['Z'][:100])                      # float64 read-back: what Projector.project returns
  np.testing.assert_allclose(rows, g['vecs'][:100], rtol=1e-11, atol=1e-12)
  np.testing.assert_allclose(prj.project_sum(g['Z']), g['vecs'].sum(axis=0), rtol=1e-11, atol=1e-9)
  lls, glls = prj.project(g['Z'][:7], grad=True)
  g2 = load_golden('lr_project_grad')
  np.testing.assert_allclose(lls, g2['lls'], rtol=1e-11, atol=1e-12)
  np.testing.assert_allclose(glls, g2['glls'], rtol=1e-12, atol=1e-14)
  vecs = prj.project_device(g['Z'])
  rng = np.random.RandomState(3)
  for _ in range(3):
    r = rng.randn(int(g['S']))
    corr = g['vecs'].dot(r)/np.sqrt((g['vecs']**2).sum(axis=1))
    f, val = vecs.argmax_dot(r)
    assert f == int(corr.argmax())
    assert val == pytest.approx(corr.max(), rel=1e-6)


def test_sparsevi_golden(bc):
  import scipy.linalg as sl
  g = load_golden('sparsevi_gaussian')
  d = int(g['d'])
  np.random.seed(int(g['seed']))
  xs = np.random.multivariate_normal(np.ones(d), np.eye(d), int(g['N']))

  def sampler_w(n, wts, pts):
    if wts is None or pts is None or pts.shape[0] == 0:
      wts, pts = np.zeros(1), np.zeros((1, d))
    L = np.linalg.cholesky(np.eye(d) + wts.sum()*np.eye(d))
    U = sl.solve_triangular(L, np.eye(d), lower=True, overwrite_b=True, check_finite=False).T
    mu = np.dot(U.dot(U.T), np.dot(np.eye(d), np.zeros(d)) + np.dot(np.eye(d), (wts[:, np.newaxis]*pts).sum(axis=0)))
    return mu + np.random.randn(n, d).dot(U.T)
  prj = bc.GaussianProjector(sampler_w, int(g['S']), np.eye(d))
  svi = bc.SparseVICoreset(xs, prj, opt_itrs=int(g['opt_itrs']))
  svi.build(int(g['itrs']))
  assert np.array_equal(svi.idcs, g['raw_idcs'])
  np.testing.assert_allclose(svi.wts, g['raw_wts'], rtol=1e-6, atol=1e-9)
  wts, pts, idcs = svi.get()
  assert np.array_equal(pts, g['pts']) and svi.error() == 0.


def test_sparsevi_subsampled_vs_oracle(bc):
  from oracle import coresets
  rng = np.random.RandomState(5)
  x = rng.randn(2000, 4) + 1.
  th = rng.randn(24, 4)

  def make(prj_cls, **kw):
    return prj_cls(lambda n, w, p: th + 0.01*np.random.randn(*th.shape), 24, **kw)
  np.random.seed(3)
  a = bc.SparseVICoreset(x, make(bc.GaussianProjector, Siginv=np.eye(4)), n_subsample_select=500, n_subsample_opt=300, opt_itrs=8)
  a.build(5)
  np.random.seed(3)
  f = lambda xx, tt: models.gaussian_loglik(xx, tt, np.eye(4), 0.)
  o = coresets.SparseVIOracle(x, models.OracleProjector(lambda n, w, p: th + 0.01*np.random.randn(*th.shape), 24, f),
                              n_subsample_select=500, n_subsample_opt=300, opt_itrs=8)
  o.build(5)
  assert np.array_equal(a.idcs, o.idcs)
  np.testing.assert_allclose(a.wts, o.wts, rtol=1e-6, atol=1e-9)


def test_bpsvi_golden(bc):
  g = load_golden('bpsvi_lr')
  Z, theta = lr_problem(int(g['seed']), int(g['N']), int(g['d']), int(g['S']))
  np.random.seed(int(g['build_seed']))
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, int(g['S']))
  bp = bc.BatchPSVICoreset(Z, prj, opt_itrs=int(g['opt_itrs']))
  bp.build(int(g['sz']))
  np.testing.assert_allclose(bp.wts, g['wts'], rtol=1e-9, atol=1e-12)
  np.testing.assert_allclose(bp.pts, g['pts'], rtol=1e-9, atol=1e-12)


def test_bpsvi_poisson_gradient_vs_oracle(bc):
  """config 5 shape in miniature: Poisson BatchPSVI gradient (with the documented repaired grad_z)"""
  from oracle import coresets
  g = load_golden('poisson_project_small')
  Z, th = g['Z'], g['theta']
  prj = bc.PoissonProjector(lambda n, w, p: th, th.shape[0])
  bp = bc.BatchPSVICoreset(Z, prj, opt_itrs=1)
  o = coresets.BatchPSVIOracle(Z, models.OracleProjector(lambda n, w, p: th, th.shape[0], models.poisson_loglik,
                                                         models.poisson_grad_z_loglik_fixed), opt_itrs=1)
  sz, d = 6, Z.shape[1]
  x = np.hstack((np.full(sz, Z.shape[0]/sz), Z[:sz].reshape(-1)))
  np.testing.assert_allclose(bp.gradient(x.copy(), sz, d), o.gradient(x.copy(), sz, d), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('n,d,S', [(5000, 40, 200), (4500, 130, 512), (4200, 25, 600)])
def test_project_sum_tiled_kernel_vs_oracle(bc, n, d, S):
  """K3b register-tiled float64 kernel (n >= 4096, d >= 24): column sums without the N x S matrix"""
  rng = np.random.RandomState(n + d)
  X = rng.randn(n, d)
  th = rng.randn(S, d)/np.sqrt(d)
  ref = models.project(models.lr_loglik, X, th).sum(axis=0)
  got = bc.LogisticRegressionProjector(lambda k, w, p: th, S).project_sum(X)
  np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9*np.abs(ref).max())
  Siginv = np.eye(d) + 0.05*np.ones((d, d))
  f = lambda x, t: models.gaussian_loglik(x, t, Siginv, 0.)
  ref = models.project(f, X, th).sum(axis=0)
  got = bc.GaussianProjector(lambda k, w, p: th, S, Siginv).project_sum(X)
  np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9*np.abs(ref).max())
  y = rng.poisson(np.log1p(np.exp(X.dot(th[0])))).astype(np.float64)
  Zp = np.hstack((X, y[:, None]))
  ref = models.project(models.poisson_loglik, Zp, th).sum(axis=0)
  got = bc.PoissonProjector(lambda k, w, p: th, S).project_sum(Zp)
  np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9*np.abs(ref).max())


# ---------------------------------------------------------------- remaining API paths
def test_hilbert_subsample_golden(bc):
  """hilbert.py:13-22: sorted de-duplicated subsample from the global RNG"""
  g = load_golden('lr_subsample_giga')
  Z, theta = lr_problem(int(g['seed']), int(g['N']), int(g['d']), int(g['S']))
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, int(g['S']))
  np.random.seed(int(g['sub_seed']))
  cs = bc.HilbertCoreset(Z, prj, n_subsample=int(g['n_subsample']))
  cs.build(int(g['itrs']))
  assert np.array_equal(cs.sub_idcs, g['sub_idcs'])
  assert [e.f for e in cs.snnls.last_events] == list(g['sel'])
  wts, pts, idcs = cs.get()
  assert np.array_equal(idcs, g['sub_idcs'][g['w'] > 0]) and np.array_equal(pts, Z[idcs])
  assert_weights_close(cs.snnls.weights(), g['w'])


@pytest.mark.parametrize('alg', ['giga', 'fw'])
def test_wide_rows_use_launch_per_iteration_engine(bc, alg):
  """S > 512 is outside the persistent kernel's control warp: the scan + step kernels take over"""
  rng = np.random.RandomState(12)
  X = rng.randn(3000, 640)*rng.uniform(0.5, 2., size=(3000, 1))
  o = greedy.ORACLES[alg](X.T, X.sum(axis=0))
  oev = o.build(25)
  cs, ev = run_gpu(bc, X, alg, 25)
  assert [e.f for e in ev] == [e[1] for e in oev]
  assert_weights_close(cs.snnls.weights(), o.w)


def test_coreset_optimize_and_blackbox_projector(bc):
  """Coreset.optimize (coreset.py:47-64) through HilbertCoreset._optimize; BlackBoxProjector host callbacks"""
  g = load_golden('lr_project_small')
  prj = bc.BlackBoxProjector(lambda n, w, p: g['theta'], int(g['S']), models.lr_loglik)
  cs = bc.HilbertCoreset(g['Z'], prj, snnls=bc.snnls.GIGA)
  cs.build(30)
  e0 = cs.error()
  cs.optimize()
  assert cs.error() <= e0*(1 + 1e-12) and not cs.reached_numeric_limit
  o = greedy.GigaOracle(g['vecs'].T, g['vecs'].sum(axis=0))
  o.build(30)
  o.optimize()
  assert cs.error() == pytest.approx(o.error(), rel=1e-5)
  wts, pts, idcs = cs.get()
  assert np.array_equal(idcs, np.flatnonzero(o.w > 0))


def test_sparsevi_with_blackbox_projector(bc):
  """a user-callback projector still works with SparseVI: host evaluation, device arg-max"""
  from oracle import coresets
  rng = np.random.RandomState(7)
  x = rng.randn(800, 3) + 1.
  th = rng.randn(16, 3)
  f = lambda xx, tt: models.gaussian_loglik(xx, tt, np.eye(3), 0.)
  np.random.seed(1)
  a = bc.SparseVICoreset(x, bc.BlackBoxProjector(lambda n, w, p: th, 16, f), opt_itrs=5)
  a.build(3)
  np.random.seed(1)
  o = coresets.SparseVIOracle(x, models.OracleProjector(lambda n, w, p: th, 16, f), opt_itrs=5)
  o.build(3)
  assert np.array_equal(a.idcs, o.idcs)
  np.testing.assert_allclose(a.wts, o.wts, rtol=1e-6, atol=1e-9)


def test_device_nnls_equals_scipy_path(bc, monkeypatch):
  """OMP with the on-device NNLS (default) and with the reference's SciPy call give the same coreset"""
  g = load_golden('lr_project_small')
  prj = bc.LogisticRegressionProjector(lambda n, w, p: g['theta'], int(g['S']))
  a = bc.HilbertCoreset(g['Z'], prj, snnls=bc.snnls.OrthoPursuit)
  a.build(50)
  from scipy_omp import run_scipy_omp
  b = bc.HilbertCoreset(g['Z'], prj, snnls=bc.snnls.OrthoPursuit)
  bev = run_scipy_omp(b.snnls, 50)
  assert [e.f for e in a.snnls.last_events] == [e[1] for e in bev]
  np.testing.assert_allclose(a.snnls.weights(), b.snnls.weights(), rtol=1e-6, atol=1e-9*b.snnls.weights().max())
  assert a.error() == pytest.approx(b.error(), rel=1e-6)
  # incremental builds continue from the kept factorisation
  c = bc.HilbertCoreset(g['Z'], prj, snnls=bc.snnls.OrthoPursuit)
  for k in (1, 4, 20, 25):
    c.build(k)
  np.testing.assert_allclose(c.snnls.weights(), a.snnls.weights(), rtol=1e-6, atol=1e-9*a.snnls.weights().max())


def test_handles_release_device_memory(bc):
  """every handle frees what it allocated (matrices, solver state, NNLS work space, datasets, probes)"""
  import gc
  ctx = bc.Context.default()
  g = load_golden('lr_project_small')
  prj = bc.LogisticRegressionProjector(lambda n, w, p: g['theta'], int(g['S']))

  def cycle():
    for cls in (bc.snnls.GIGA, bc.snnls.FrankWolfe, bc.snnls.OrthoPursuit):
      cs = bc.HilbertCoreset(g['Z'], prj, snnls=cls)
      cs.build(12)
      cs.optimize()
    v = prj.project_device(g['Z'])
    v.argmax_dot(np.ones(int(g['S'])))
    prj.project_sum(g['Z'], cache=False)
    svi = bc.SparseVICoreset(g['Z'][:500], prj, opt_itrs=2)
    svi.build(2)
  cycle()
  gc.collect()
  ctx.synchronize()
  free0, _ = ctx.mem_info()
  for _ in range(5):
    cycle()
  gc.collect()
  ctx.synchronize()
  free1, _ = ctx.mem_info()
  assert free0 - free1 < (8 << 20), (free0, free1)


def test_duplicate_rows_resolve_to_lowest_index(bc):
  """exact duplicates score identically in float32 and float64: ndarray.argmax picks the first one"""
  rng = np.random.RandomState(21)
  base = rng.randn(6000, 96)
  X = np.vstack((base, base[::-1].copy()))           # row i and row 11999 - i are identical
  o = greedy.GigaOracle(X.T, X.sum(axis=0))
  oev = o.build(40)
  cs, ev = run_gpu(bc, X, 'giga', 40)
  assert [e.f for e in ev] == [e[1] for e in oev]
  assert all(e.f < 6000 for e in ev)


@pytest.mark.parametrize('alg', ['giga', 'fw'])
def test_tolerance_and_failure_latch_follow_the_reference(bc, alg):
  """util.TOL is honoured (giga.py:28): with a huge tolerance GIGA's selection fails at once, is retried once
  and latches the numeric limit; FW does not use TOL and keeps going"""
  np.random.seed(4)
  X = np.random.randn(500, 30)
  bc.util.set_tolerance(10.)
  try:
    o = greedy.ORACLES[alg](X.T, X.sum(axis=0), tol=10.)
    oev = o.build(6)
    cs, ev = run_gpu(bc, X, alg, 6)
  finally:
    bc.util.set_tolerance(1e-12)
  assert [(e.code, e.f) for e in ev] == [(e[0], e[1]) for e in oev]
  assert cs.snnls.reached_numeric_limit == o.reached_numeric_limit
  assert cs.snnls.size() == o.size()


def test_device_row_gather_and_grad_contract(bc):
  """subsampled projections gather rows on the device; the pseudo-point gradient contraction equals the
  reference's (K, S, d) formulation (bpsvi.py:53) for all three models"""
  g = load_golden('lr_project_small')
  Z, th = g['Z'], g['theta']
  prj = bc.LogisticRegressionProjector(lambda n, w, p: th, th.shape[0])
  sub = np.array([5, 17, 5, 2999, 0, 1024, 77], dtype=np.int64)
  np.testing.assert_allclose(prj.project_sum(Z, sub=sub), g['vecs'][sub].sum(axis=0), rtol=1e-11, atol=1e-11)
  np.testing.assert_allclose(prj.project_device(Z, sub=sub).norms(), np.sqrt((g['vecs'][sub]**2).sum(axis=1)), rtol=1e-9)
  big = np.random.RandomState(0).randint(3000, size=5000)
  Zw = np.hstack((Z, np.random.RandomState(1).randn(3000, 30)))       # d = 36 -> tiled column-sum kernel
  thw = np.random.RandomState(2).randn(64, 36)/6.
  prjw = bc.LogisticRegressionProjector(lambda n, w, p: thw, 64)
  np.testing.assert_allclose(prjw.project_sum(Zw, sub=big), models.project(models.lr_loglik, Zw[big], thw).sum(axis=0),
                             rtol=1e-9, atol=1e-9)
  rng = np.random.RandomState(3)
  w, resid = rng.rand(7) + 0.1, rng.randn(th.shape[0])
  pts = Z[:7]

  def reference(glls):
    glls = glls - glls.mean(axis=2)[:, :, None]
    return -(w[:, None, None]*glls*resid[None, :, None]).sum(axis=1)/th.shape[0]
  np.testing.assert_allclose(prj.grad_contract(pts, w, resid), reference(models.lr_grad_z_loglik(pts, th)), rtol=1e-10, atol=1e-12)
  Siginv = np.eye(6) + 0.1
  gp = bc.GaussianProjector(lambda n, w_, p: th, th.shape[0], Siginv)
  np.testing.assert_allclose(gp.grad_contract(pts, w, resid), reference(models.gaussian_grad_x_loglik(pts, th, Siginv)),
                             rtol=1e-10, atol=1e-12)
  zp = np.hstack((pts, rng.poisson(2., (7, 1)).astype(float)))
  pp = bc.PoissonProjector(lambda n, w_, p: th, th.shape[0])
  np.testing.assert_allclose(pp.grad_contract(zp, w, resid), reference(models.poisson_grad_z_loglik_fixed(zp, th)),
                             rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------- round-1 late additions: wide OMP products, table links, pinned sources
def test_omp_wide_products_equal_narrow(bc, monkeypatch):
  """the warp-split K x S products of the OMP / NNLS iteration (default) against the one-thread-per-output
  forms (BCG_OMP_WIDE=0), S above and below one 256-column round, K past S/2"""
  for N, d, S, itrs in ((20000, 6, 300, 120), (5000, 5, 96, 60)):
    Z, theta = lr_problem(3, N, d, S)
    prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
    a = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
    a.build(itrs)
    monkeypatch.setenv('BCG_OMP_WIDE', '0')
    b = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
    b.build(itrs)
    monkeypatch.delenv('BCG_OMP_WIDE')
    assert [(e.code, e.f) for e in a.snnls.last_events] == [(e.code, e.f) for e in b.snnls.last_events]
    wa, wb = a.snnls.weights(), b.snnls.weights()
    np.testing.assert_allclose(wa, wb, rtol=1e-7, atol=1e-10*wb.max())
    assert a.error() == pytest.approx(b.error(), rel=1e-7)
    a.optimize()
    b.optimize()
    np.testing.assert_allclose(a.snnls.weights(), b.snnls.weights(), rtol=1e-7, atol=1e-10*wb.max())


def test_omp_medium_vs_oracle(bc):
  """OrthoPursuit at a size where the oracle (SciPy NNLS per iteration) still finishes in seconds"""
  N, d, S, itrs = 60000, 8, 256, 60
  Z, theta = lr_problem(5, N, d, S)
  vecs = models.project(models.lr_loglik, Z, theta)
  o = greedy.OrthoPursuitOracle(vecs.T, vecs.sum(axis=0))
  oev = o.build(itrs)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  cs = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
  cs.build(itrs)
  assert [e.f for e in cs.snnls.last_events] == [e[1] for e in oev]
  assert_weights_close(cs.snnls.weights(), o.w)
  assert_errors_close(cs.error(), o.error(), vecs, o.w)


def test_table_links_equal_libdevice_links(bc, monkeypatch):
  """softplus-table links (default) against the libdevice exp / log1p links (BCG_FAST_LINK=0, read when a
  context is created): materialised rows, float64 read-backs and K3b column sums, LR and Poisson"""
  fast = bc.Context.default()
  monkeypatch.setenv('BCG_FAST_LINK', '0')
  slow = bc.Context(0)
  monkeypatch.delenv('BCG_FAST_LINK')
  rng = np.random.RandomState(11)
  for d, S, n in ((10, 512, 3000), (40, 200, 5000)):
    X = rng.randn(n, d)*rng.choice([0.3, 1., 8., 60.], size=(n, 1))      # includes saturated rows (|m| >> 37)
    th = rng.randn(S, d)/np.sqrt(d)
    y = rng.poisson(2., size=n).astype(np.float64)
    Zp = np.hstack((X, y[:, None]))
    for model, Z in ((bc._native.MODEL_LR, X), (bc._native.MODEL_POISSON, Zp)):
      outs = []
      for ctx in (fast, slow):
        ds = bc.Dataset(Z, ctx=ctx)
        v, rows, cs = ds.project(model, th, vecs=True, rows=True, colsum=True)
        cs2 = ds.project(model, th, colsum=True)[2]                         # K3b kernels (d >= 24) or K3 without stores
        outs.append((v.to_numpy(), v.norms(), rows, cs, cs2))
      a, b = outs
      scale = np.abs(b[2]).max(axis=1, keepdims=True) + 1e-300
      assert np.max(np.abs(a[2] - b[2])/scale) < 1e-13                      # float64 centred rows
      np.testing.assert_allclose(a[1], b[1], rtol=1e-11)
      assert np.max(np.abs(a[0] - b[0])/(b[1][:, None] + 1e-300)) < 2.**-23
      for k in (3, 4):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-11, atol=1e-11*np.abs(b[k]).max())
      np.testing.assert_allclose(a[4], b[3], rtol=1e-9, atol=1e-9*np.abs(b[3]).max())


def test_pinned_source_equals_pageable_source(bc):
  """page-locked input arrays (bc.pinned_empty / pinned_copy) are uploaded by DMA without staging: same matrix"""
  Z, theta = lr_problem(9, 700000, 10, 64)                                   # 56 MB: several pipeline chunks
  Zp = bc.pinned_copy(Z)
  assert Zp.shape == Z.shape and np.array_equal(Zp, Z)
  a = bc.DeviceVecs.project_lr(Z, theta)
  b = bc.DeviceVecs.project_lr(Zp, theta)
  assert np.array_equal(a.norms(), b.norms())
  assert np.array_equal(a.sum(axis=0), b.sum(axis=0))
  assert np.array_equal(a.to_numpy(1000, 50), b.to_numpy(1000, 50))
  da, db = bc.Dataset(Z), bc.Dataset(Zp)
  ca = da.project(bc._native.MODEL_LR, theta, colsum=True)[2]
  cb = db.project(bc._native.MODEL_LR, theta, colsum=True)[2]
  assert np.array_equal(ca, cb)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 64)
  cs = bc.HilbertCoreset(Zp, prj)
  cs.build(10)
  wts, pts, idcs = cs.get()
  assert np.array_equal(pts, Z[idcs])
  del Zp, b, db, cs


@pytest.mark.parametrize('S,d', [(64, 3), (128, 32), (256, 10), (512, 10), (512, 24), (256, 200), (512, 70), (128, 16)])
def test_specialised_projection_kernels_vs_general_and_oracle(bc, monkeypatch, S, d):
  """the three materialising projection kernels -- project_mma_kernel (DMMA; BCG_PROJ_MMA=2 forces it for every S in {64,128,256,512}: sample tile
  resident for d <= 16, streamed k tiles above), project_fast_kernel (BCG_PROJ_MMA=0; d <= 32) and the general kernel
  (BCG_PROJ_MMA=0 BCG_PROJ_FAST=0) -- against each other and the oracle: three models, a row count that is no multiple of
  any row block, a device row gather, column-sum-only passes"""
  rng = np.random.RandomState(S + d)
  n = 3001
  X = rng.randn(n, d)*rng.choice([0.3, 1., 5.], size=(n, 1))/max(1., np.sqrt(d/10.))
  th = rng.randn(S, d)/np.sqrt(d)
  y = rng.poisson(2., size=n).astype(np.float64)
  Zp = np.hstack((X, y[:, None]))
  Siginv = np.eye(d) + 0.1*np.ones((d, d))/d
  sub = rng.randint(n, size=777)
  cases = [(bc._native.MODEL_LR, X, None, models.project(models.lr_loglik, X, th)),
           (bc._native.MODEL_POISSON, Zp, None, models.project(models.poisson_loglik, Zp, th)),
           (bc._native.MODEL_GAUSSIAN, X, Siginv, models.project(lambda x, t: models.gaussian_loglik(x, t, Siginv, 0.), X, th))]
  variants = [{'BCG_PROJ_MMA': '2'}, {'BCG_PROJ_MMA': '0', 'BCG_PROJ_FAST': '1'}, {'BCG_PROJ_MMA': '0', 'BCG_PROJ_FAST': '0'}]
  for model, Z, si, ref in cases:
    ds = bc.Dataset(Z)
    res = []
    for env in variants:
      for k, val in env.items():
        monkeypatch.setenv(k, val)
      v = ds.project(model, th, si, vecs=True)[0]
      vs = ds.project(model, th, si, vecs=True, sub=sub)[0]
      cs = ds.project(model, th, si, colsum=True)[2]
      vh = bc.DeviceVecs.project_host(model, Z, th, si)            # pipelined host path (chunked launches)
      res.append((v.to_numpy(), v.norms(), v.sum(axis=0), vs.to_numpy(), vs.norms(), cs, v.norm_sum(), v.zero_rows()))
      check_projection(v, ref)
      check_projection(vs, ref[sub])
      check_projection(vh, ref)
      for k in env:
        monkeypatch.delenv(k)
    for a, b in ((res[0], res[2]), (res[1], res[2])):
      compare_projection_outputs(a, b, ref)


def compare_projection_outputs(a, b, ref):
    assert np.max(np.abs(a[0] - b[0])) <= 2.**-22 and np.max(np.abs(a[3] - b[3])) <= 2.**-22   # unit rows, float32 rounding
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-12)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-10, atol=1e-10*np.abs(b[2]).max())
    np.testing.assert_allclose(a[5], b[5], rtol=1e-10, atol=1e-10*np.abs(b[5]).max())
    np.testing.assert_allclose(a[5], ref.sum(axis=0), rtol=1e-9, atol=1e-9*np.abs(ref).sum(axis=0).max())
    assert a[6] == pytest.approx(b[6], rel=1e-12) and a[7] == b[7] == 0


# ---------------------------------------------------------------- independent float64 audit (bcg_dataset_audit)
def audit_score_fn(nat, ds, model, theta, Siginv=None):
  def fn(kind, dirs):
    return ds.audit(model, theta, Siginv, kind=nat.ALG_GIGA if kind == 'giga' else nat.ALG_FW, dirs=dirs)[0]
  return fn


def test_audit_kernel_vs_oracle(bc):
  """the audit scorer itself against the dense oracle: scores of one GIGA and one FW iteration, norms, column sums"""
  import bayesiancoresets_b200._native as nat
  Z, theta = lr_problem(5, 20000, 7, 200)
  vecs = models.project(models.lr_loglik, Z, theta)
  ds = nat.Dataset(Z)
  for alg in ('giga', 'fw'):
    o = greedy.ORACLES[alg](vecs.T, vecs.sum(axis=0))
    o.build(7)
    ref = o.scores()
    if alg == 'giga':
      xw, nw = o._unit_iterate()
      xw = xw/nw
      cdir = o.bn - o.bn.dot(xw)*xw
      dirs = np.vstack((cdir/np.sqrt((cdir**2).sum()), xw))
    else:
      dirs = (o.b - o.A.dot(o.w))[np.newaxis, :]
    sc, nr, cs = ds.audit(nat.MODEL_LR, theta, kind=nat.ALG_GIGA if alg == 'giga' else nat.ALG_FW, dirs=dirs, norms=True,
                          colsum=True)
    np.testing.assert_allclose(sc, ref, rtol=1e-9, atol=1e-11*np.abs(ref).max())
    assert sc.argmax() == ref.argmax()
    np.testing.assert_allclose(nr, np.sqrt((vecs**2).sum(axis=1)), rtol=1e-11)
    np.testing.assert_allclose(cs, vecs.sum(axis=0), rtol=1e-9, atol=1e-10*np.abs(vecs).sum(axis=0).max())
  # Gaussian and Poisson models: residual scores against dense NumPy
  rng = np.random.RandomState(3)
  x = rng.randn(3000, 9) + 1.
  th = rng.randn(70, 9)
  Si = np.eye(9) + 0.2*np.ones((9, 9))
  vg = models.project(lambda a, t: models.gaussian_loglik(a, t, Si, 0.), x, th)
  r = rng.randn(70)
  sc = nat.Dataset(x).audit(nat.MODEL_GAUSSIAN, th, Si, dirs=r[np.newaxis, :])[0]
  ref = vg.dot(r)/np.sqrt((vg**2).sum(axis=1))
  np.testing.assert_allclose(sc, ref, rtol=1e-8, atol=1e-10*np.abs(ref).max())
  g = load_golden('poisson_project_small')
  vp, thp = g['vecs'], g['theta']
  r = rng.randn(thp.shape[0])
  sc = nat.Dataset(g['Z']).audit(nat.MODEL_POISSON, thp, dirs=r[np.newaxis, :])[0]
  ref = vp.dot(r)/np.sqrt((vp**2).sum(axis=1))
  np.testing.assert_allclose(sc, ref, rtol=1e-8, atol=1e-10*np.abs(ref).max())


FULL_SIZE_REPORT = []


@pytest.mark.parametrize('alg,N,S,itrs', [('giga', 1000000, 256, 60), ('fw', 1000000, 256, 60), ('omp', 1000000, 256, 60),
                                         ('giga', 10000000, 512, 40), ('fw', 10000000, 512, 30), ('omp', 10000000, 512, 30)])
def test_full_size_selection_equals_float64_audit(bc, alg, N, S, itrs):
  """BASELINE configs[1] and the north-star / configs[3] size: the engine's whole build against the reference's
  algorithm replayed in float64 (oracle/replay.py), where every O(N S) scoring pass is the independent audit kernel
  (float64 from the raw data; no float32 storage, no shared code).  Indices exact, weights / error to 1e-5."""
  import bayesiancoresets_b200._native as nat
  from oracle import replay
  Z, theta = lr_problem(0, N, 10, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  cs = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  cs.build(itrs)
  ev = cs.snnls.last_events
  ds = nat.Dataset(Z)
  _, norms, b = ds.audit(nat.MODEL_LR, theta, norms=True, colsum=True)
  np.testing.assert_allclose(cs.snnls.b, b, rtol=1e-9, atol=1e-9*np.abs(b).max())
  kw = {'norm_sum': norms.sum()} if alg == 'fw' else {}
  r = replay.REPLAYS[alg](N, b, audit_score_fn(nat, ds, nat.MODEL_LR, theta),
                          lambda idx: models.project(models.lr_loglik, Z[idx], theta), **kw)
  rev = r.build(itrs)
  gaps = np.array([g[1] for g in r.diag])
  inwin = np.array([g[2] for g in r.diag])
  FULL_SIZE_REPORT.append('audit %s N=%d S=%d: %d iterations, min top-2 gap %.3e (median %.3e), max rows inside the '
                          'float32 near-tie window %d' % (alg, N, S, len(rev), gaps.min(), np.median(gaps), inwin.max()))
  print(FULL_SIZE_REPORT[-1])
  assert [(e.code, e.f) for e in ev] == [(e[0], e[1]) for e in rev]
  w = cs.snnls.weights()
  assert_weights_close(w, r.w)
  norms_act = norms[r.idx]
  atol = 2.**-24*float(np.abs(r.wa).dot(norms_act)) + 1e-300
  np.testing.assert_allclose([e.error for e in ev], [e[2] for e in rev], rtol=W_RTOL, atol=atol)


def test_full_size_vs_dense_host_oracle(bc):
  """the affordable host test: the dense float64 oracle (= the reference, bit for bit) at N = 1e6, S = 256, 60 GIGA
  iterations -- indices exact, weights and per-iteration error() to 1e-5 (snnls.py:31-79)"""
  N, S, itrs = 1000000, 256, 60
  Z, theta = lr_problem(0, N, 10, S)
  vecs = models.project(models.lr_loglik, Z, theta)
  o = greedy.GigaOracle(vecs.T, vecs.sum(axis=0))
  oev = o.build(itrs)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  cs = bc.HilbertCoreset(Z, prj)
  cs.build(itrs)
  ev = cs.snnls.last_events
  assert [(e.code, e.f) for e in ev] == [(e[0], e[1]) for e in oev]
  assert_weights_close(cs.snnls.weights(), o.w)
  assert_errors_close([e.error for e in ev], [e[2] for e in oev], vecs, o.w)


# ---------------------------------------------------------------- exactness of the selection, API holes
@pytest.mark.parametrize('alg,engine', [('giga', '2'), ('fw', '2'), ('omp', '2'), ('giga', '1')])
def test_forced_exact_selection_equals_normal_path(bc, monkeypatch, alg, engine):
  """every selection through the exact float64 pass (exact_scan_kernel; for GIGA / FW the persistent kernel stops and
  is relaunched at every iteration) gives the same events as the float32 candidate path"""
  monkeypatch.setenv('BCG_ENGINE', engine)
  Z, theta = lr_problem(13, 30000, 6, 128)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 128)
  a = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  a.build(25)
  b = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  b.snnls._native.set_force_exact(True)
  b.build(12)
  b.build(13)                                   # incremental: the retry flag / state survive the relaunches
  eva = [(e.code, e.f) for e in a.snnls.last_events]
  assert b.snnls._native.exact_count() == 25
  assert a.snnls._native.exact_count() == 0
  np.testing.assert_allclose(b.snnls.weights(), a.snnls.weights(), rtol=1e-12, atol=0)
  assert b.error() == pytest.approx(a.error(), rel=1e-12)
  assert eva[12:] == [(e.code, e.f) for e in b.snnls.last_events]


def test_near_ties_beyond_the_candidate_set_take_the_exact_pass(bc):
  """13 rows whose scores differ by ~1e-6 (well inside the float32 near-tie window of 2e-5 + 1e-5 |top|, well above the
  ~4e-9 noise of float32 storage), 12 of them adjacent: one warp scans them and publishes ONE candidate.  The engine
  must notice the unpublished in-window scores and return the float64 arg-max (= the oracle's selection)."""
  rng = np.random.RandomState(4)
  N, S = 40000, 96
  X = rng.randn(N, S)
  f0 = greedy.GigaOracle(X.T, X.sum(axis=0)).select()
  r = X[f0].copy()
  X[20000:20012] = r
  b = X.sum(axis=0)
  bh, rh = b/np.linalg.norm(b), r/np.linalg.norm(r)
  p = bh - bh.dot(rh)*rh
  p /= np.linalg.norm(p)
  for k in range(12):                             # cos(row, b) grows by ~1e-6 per step; the best one comes last
    X[20000 + k] = np.linalg.norm(r)*(rh + 1e-6*(k + 1)*p)
  for alg in ('giga', 'fw'):
    o = greedy.ORACLES[alg](X.T, X.sum(axis=0))
    oev = o.build(6)
    assert oev[0][1] == 20011
    s = algs(bc)[alg](X.T, X.sum(axis=0))
    s.build(6)
    assert [e.f for e in s.last_events] == [e[1] for e in oev]
    assert s._native.exact_count() >= 1


@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_check_error_monotone_false_follows_the_reference(bc, alg):
  """snnls.py:9,46,56: without the monotone test the retry flag is never cleared by a successful step"""
  X = np.eye(12)
  o = greedy.ORACLES[alg](X.T, X.sum(axis=0), check_error_monotone=False)
  oev = o.build(20)
  s = algs(bc)[alg](X.T, X.sum(axis=0), check_error_monotone=False)
  s.build(20)
  # (FW: once all 12 axes are in, the residual is round-off noise and so is the next selection)
  nsel = 12 if alg == 'fw' else len(oev)
  assert [(e.code, e.f) for e in s.last_events][:nsel] == [(e[0], e[1]) for e in oev][:nsel]
  if alg != 'fw':
    assert s.reached_numeric_limit == o.reached_numeric_limit
  np.random.seed(3)
  Y = np.random.randn(3000, 40)
  o = greedy.ORACLES[alg](Y.T, Y.sum(axis=0), check_error_monotone=False)
  oev = o.build(60)
  s = algs(bc)[alg](Y.T, Y.sum(axis=0), check_error_monotone=False)
  s.build(60)
  nsel = len(oev) if alg != 'omp' else 24
  assert [(e.code, e.f) for e in s.last_events][:nsel] == [(e[0], e[1]) for e in oev][:nsel]


def test_sampling_baselines_follow_the_reference(bc):
  """snnls/sampling.py and coreset/sampling.py: same draws (global RNG), same weights, error() from the device"""
  ref, _ = _reference_package()
  rng = np.random.RandomState(0)
  X = rng.randn(500, 30)*rng.uniform(0.2, 4., size=(500, 1))
  for name in ('ImportanceSampling', 'UniformSampling'):
    np.random.seed(7)
    r = getattr(ref.snnls, name)(X.T, X.sum(axis=0))
    r.build(40); r.build(25)
    np.random.seed(7)
    s = getattr(bc.snnls, name)(X.T, X.sum(axis=0))
    s.build(40); s.build(25)
    np.testing.assert_allclose(s.weights(), r.weights(), rtol=1e-12)
    assert s.error() == pytest.approx(r.error(), rel=1e-6) and s.size() == r.size()
    r.optimize(); s.optimize()
    assert s.error() == pytest.approx(r.error(), rel=1e-5)
  np.random.seed(3)
  a = ref.UniformSamplingCoreset(X); a.build(60)
  np.random.seed(3)
  b = bc.UniformSamplingCoreset(X); b.build(60)
  for u, v in zip(a.get(), b.get()):
    assert np.array_equal(u, v)
  # assigning the dense weight vector is a sparse active-set write
  g = bc.snnls.GIGA(X.T, X.sum(axis=0))
  w = np.zeros(500); w[[3, 77, 400]] = [1.5, 0.25, 2.]
  g.w = w
  assert np.array_equal(g.weights(), w)
  assert g.error() == pytest.approx(np.linalg.norm(w.dot(X) - X.sum(axis=0)), rel=1e-6)


@pytest.mark.parametrize('alg,engine', [('omp', '2'), ('giga', '1'), ('giga', '2'), ('fw', '1')])
def test_non_finite_rows_fail_cleanly(bc, monkeypatch, alg, engine):
  """NaN in the data: b and the direction become NaN, the scan yields no candidate.  Every engine must report an
  error (BCG_ERR_STATE) instead of reading out of bounds; the context stays usable afterwards."""
  monkeypatch.setenv('BCG_ENGINE', engine)
  rng = np.random.RandomState(0)
  X = rng.randn(5000, 32)
  X[17, 3] = np.nan
  s = algs(bc)[alg](X.T, X.sum(axis=0))
  with pytest.raises(bc.BcgError):
    s.build(5)
  Y = rng.randn(500, 32)
  t = bc.snnls.GIGA(Y.T, Y.sum(axis=0))
  t.build(5)
  assert t.size() > 0


# ---------------------------------------------------------------- the drop-in boundary, exercised from the REFERENCE side
def _reference_package(tmp_path=None):
  """the unmodified reference installed under baseline/_ref (it travels to the GPU box); a private copy when the test
  adds the INTEGRATION.md stub to its tree"""
  import importlib
  import shutil
  import sys
  from conftest import ROOT
  import os
  src = os.path.join(ROOT, 'baseline', '_ref')
  if not os.path.isdir(os.path.join(src, 'bayesiancoresets')):
    src = '/root/reference'
  if not os.path.isdir(os.path.join(src, 'bayesiancoresets')):
    pytest.skip('the reference package is not available on this box')
  if tmp_path is not None:
    shutil.copytree(os.path.join(src, 'bayesiancoresets'), os.path.join(str(tmp_path), 'bayesiancoresets'))
    src = str(tmp_path)
  for m in [m for m in sys.modules if m == 'bayesiancoresets' or m.startswith('bayesiancoresets.')]:
    del sys.modules[m]
  sys.path.insert(0, src)
  try:
    return importlib.import_module('bayesiancoresets'), src
  finally:
    sys.path.remove(src)


@pytest.mark.parametrize('alg', ['GIGA', 'FrankWolfe', 'OrthoPursuit'])
def test_reference_hilbert_coreset_runs_on_this_repos_snnls(bc, alg):
  """the literal drop-in of hilbert.py:7,24: the REFERENCE's own HilbertCoreset (and BlackBoxProjector) with this repo's
  solver class passed as `snnls=` -- build / get / error / optimize / reset -- against the pure reference run"""
  ref, _ = _reference_package()
  Z, theta = lr_problem(2, 4000, 5, 64)
  prj = ref.BlackBoxProjector(lambda n, w, p: theta, 64, models.lr_loglik)
  pure = ref.HilbertCoreset(Z, prj, snnls=getattr(ref.snnls, alg))
  ours = ref.HilbertCoreset(Z, prj, snnls=getattr(bc.snnls, alg))
  for cs in (pure, ours):
    cs.build(20)
    cs.build(15)
  (w0, p0, i0), (w1, p1, i1) = pure.get(), ours.get()
  assert np.array_equal(i0, i1) and np.array_equal(p0, p1)
  np.testing.assert_allclose(w1, w0, rtol=1e-5, atol=1e-5*np.abs(w0).max())
  assert ours.error() == pytest.approx(pure.error(), rel=1e-5, abs=1e-7*np.sqrt((models.project(models.lr_loglik, Z, theta)**2).sum()))
  pure.optimize(); ours.optimize()
  assert ours.error() == pytest.approx(pure.error(), rel=1e-4, abs=1e-7*np.sqrt((models.project(models.lr_loglik, Z, theta)**2).sum()))
  ours.reset()
  assert ours.size() == 0


def test_integration_md_stub_runs_inside_the_reference_tree(bc, tmp_path):
  """INTEGRATION.md section 2 verbatim: the ctypes stub a maintainer would add as bayesiancoresets/snnls/_b200.py,
  dropped into a copy of the reference tree and used as `snnls=` of the reference's HilbertCoreset"""
  import importlib
  import os
  import re
  import sys
  from conftest import ROOT
  from bayesiancoresets_b200 import _native as nat
  text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  sec = text[text.index('## 2.'):text.index('## 3.')]
  code = re.search(r'```python\n(.*?)```', sec, re.S).group(1)
  assert "ctypes.CDLL('libbcg_b200.so')" in code
  code = code.replace("ctypes.CDLL('libbcg_b200.so')", 'ctypes.CDLL(%r)' % nat.LIB_PATH)
  ref, src = _reference_package(tmp_path)
  with open(os.path.join(src, 'bayesiancoresets', 'snnls', '_b200.py'), 'w') as f:
    f.write(code)
  sys.path.insert(0, src)
  try:
    stub = importlib.import_module('bayesiancoresets.snnls._b200')
  finally:
    sys.path.remove(src)
  np.random.seed(1)
  X = np.random.randn(1000, 50)

  class IDP(ref.Projector):
    def project(self, pts, grad=False):
      return pts

    def update(self, wts, pts):
      pass
  a = ref.HilbertCoreset(X, IDP(), snnls=stub.GIGA)
  a.build(100)
  wts, pts, idcs = a.get()
  g = load_golden('c1_normal_giga')
  o = greedy.GigaOracle(X.T, X.sum(axis=0))
  o.build(100)
  assert np.array_equal(idcs, np.flatnonzero(o.w > 0))
  assert_weights_close(wts, o.w[o.w > 0])
  assert_errors_close(a.error(), o.error(), X, o.w)


# ---------------------------------------------------------------- never-materialising solver
@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
def test_never_materialising_solver_equals_materialised_and_oracle(bc, alg):
  """HilbertCoreset(..., materialize=False): no N x S matrix; every selection pass re-evaluates the rows from the raw data
  in float64 (lazy_select_kernel).  Same events / weights / error as the oracle and as the resident-matrix engine."""
  Z, theta = lr_problem(6, 30000, 7, 160)
  vecs = models.project(models.lr_loglik, Z, theta)
  o = greedy.ORACLES[alg](vecs.T, vecs.sum(axis=0))
  oev = o.build(30)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 160)
  a = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg], materialize=False)
  free0 = bc.Context.default().mem_info()[0]
  assert a.snnls._vecs.shape == (30000, 160)
  np.testing.assert_allclose(a.snnls.b, vecs.sum(axis=0), rtol=1e-9, atol=1e-9*np.abs(vecs).sum(axis=0).max())
  a.build(18)
  a.build(12)
  ev = [(e.code, e.f) for e in a.snnls.last_events]
  assert ev == [(e[0], e[1]) for e in oev][18:]
  assert_weights_close(a.snnls.weights(), o.w)
  assert_errors_close(a.error(), o.error(), vecs, o.w)
  wts, pts, idcs = a.get()
  assert np.array_equal(idcs, np.flatnonzero(o.w > 0)) and np.array_equal(pts, Z[idcs])
  a.optimize()
  o.optimize()
  assert a.error() == pytest.approx(o.error(), rel=1e-5)
  with pytest.raises(bc.BcgError):
    a.snnls._vecs.to_numpy(0, 1)                    # there are no stored rows


def test_never_materialising_solver_other_models(bc):
  g = load_golden('poisson_project_small')
  prj = bc.PoissonProjector(lambda n, w, p: g['theta'], g['theta'].shape[0])
  o = greedy.GigaOracle(g['vecs'].T, g['vecs'].sum(axis=0))
  oev = o.build(15)
  a = bc.HilbertCoreset(g['Z'], prj, materialize=False)
  a.build(15)
  assert [e.f for e in a.snnls.last_events] == [e[1] for e in oev]
  assert_weights_close(a.snnls.weights(), o.w)
  g = load_golden('gaussian_project_small')
  prj = bc.GaussianProjector(lambda n, w, p: g['theta'], g['theta'].shape[0], g['Siginv'])
  o = greedy.FrankWolfeOracle(g['vecs'].T, g['vecs'].sum(axis=0))
  oev = o.build(15)
  a = bc.HilbertCoreset(g['x'], prj, snnls=bc.snnls.FrankWolfe, materialize=False)
  a.build(15)
  assert [e.f for e in a.snnls.last_events] == [e[1] for e in oev]
  assert_weights_close(a.snnls.weights(), o.w)


def test_omp_capacity_growth_keeps_the_warm_start_and_stays_small(bc):
  """the NNLS factorisation is sized by min(capacity, S + 1) and survives the growth of the active-set capacity: several
  build() calls that cross the growth points give the one-shot result, and build(20000) does not ask for gigabytes"""
  Z, theta = lr_problem(9, 20000, 6, 96)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, 96)
  one = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
  one.build(130)
  inc = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
  for k in (30, 30, 30, 40):
    inc.build(k)
  ev1 = [e.f for e in one.snnls.last_events]
  assert [e.f for e in inc.snnls.last_events] == ev1[90:]
  np.testing.assert_allclose(inc.snnls.weights(), one.snnls.weights(), rtol=1e-9, atol=1e-12*one.snnls.weights().max())
  free0 = bc.Context.default().mem_info()[0]
  big = bc.HilbertCoreset(Z, prj, snnls=bc.snnls.OrthoPursuit)
  big.build(20000)                                   # stops at the numeric limit long before; the work space is what matters
  assert free0 - bc.Context.default().mem_info()[0] < (1 << 30)


# ---------------------------------------------------------------- device sampler
def test_gaussian_posterior_sampler_on_the_device(bc):
  """bc.GaussianPosteriorSampler = the sampler_w of examples/gaussian/main.py:107-113 with weighted_post
  (model_gaussian.py:23-30) on the device: posterior mean / factor against the reference-generated fixture, the seeded
  draw (host RNG, reference order) to 1e-10, and the SparseVI golden run reproduced with it"""
  g = load_golden('gaussian_weighted_post')
  smp = bc.GaussianPosteriorSampler(g['mu0'], g['Sig0inv'], g['Siginv'])
  mup, U = smp.weighted_post(g['pts'], g['wts'])
  np.testing.assert_allclose(mup, g['mup'], rtol=1e-11, atol=1e-13)
  np.testing.assert_allclose(U, g['USigp'], rtol=1e-11, atol=1e-14)
  mup0, U0 = smp.weighted_post(np.zeros((0, 9)), np.zeros(0))
  np.testing.assert_allclose(mup0, g['mup_empty'], rtol=1e-12, atol=0)
  np.testing.assert_allclose(U0, g['USigp_empty'], rtol=1e-11, atol=1e-14)
  np.random.seed(int(g['draw_seed']))
  draw = smp(40, g['wts'], g['pts'])
  np.testing.assert_allclose(draw, g['draw'], rtol=1e-10, atol=1e-12)
  # larger dimension against the oracle restatement (d = 200 is the Gaussian example's size)
  rng = np.random.RandomState(1)
  d = 200
  B = rng.randn(d, d)/np.sqrt(d)
  Si = B.dot(B.T) + np.eye(d)
  pts, w = rng.randn(30, d), rng.uniform(0.1, 50., size=30)
  big = bc.GaussianPosteriorSampler(rng.randn(d), np.eye(d), Si)
  rm, rU, _ = models.gaussian_weighted_post(big.mu0, big.Sig0inv, big.Siginv, pts, w)
  mup, U = big.weighted_post(pts, w)
  np.testing.assert_allclose(mup, rm, rtol=1e-9, atol=1e-11)
  np.testing.assert_allclose(U, rU, rtol=1e-9, atol=1e-12)
  # the SparseVI fixture (generated by the reference with its own sampler_w) reproduced with the device sampler
  g = load_golden('sparsevi_gaussian')
  d = int(g['d'])
  np.random.seed(int(g['seed']))
  xs = np.random.multivariate_normal(np.ones(d), np.eye(d), int(g['N']))
  prj = bc.GaussianProjector(bc.GaussianPosteriorSampler(np.zeros(d), np.eye(d), np.eye(d)), int(g['S']), np.eye(d))
  svi = bc.SparseVICoreset(xs, prj, opt_itrs=int(g['opt_itrs']))
  svi.build(int(g['itrs']))
  assert np.array_equal(svi.idcs, g['raw_idcs'])
  np.testing.assert_allclose(svi.wts, g['raw_wts'], rtol=1e-6, atol=1e-9)
  with pytest.raises(bc.BcgError):
    bc.GaussianPosteriorSampler(np.zeros(3), -np.eye(3), np.eye(3))(4, np.zeros(1), np.zeros((1, 3)))


def test_laplace_sampler_reductions_on_the_device(bc):
  """bcg_glm_joint against the reference-generated fixture (log_joint / grad / hess of model_lr.py and model_poiss.py), and
  bc.LaplaceSampler (= sampler_w of examples/logistic_poisson_regression/main.py:153-160) against the oracle's get_laplace"""
  import bayesiancoresets_b200._native as nat
  g = load_golden('glm_joint')
  for name, model, Z, th in (('lr', nat.MODEL_LR, g['Z_lr'], g['th_lr']), ('poiss', nat.MODEL_POISSON, g['Z_poiss'], g['th_poiss'])):
    for s in range(th.shape[0]):
      v, gr, H = nat.glm_joint(model, Z, g['w'], th[s], hess=True)
      assert v == pytest.approx(g[name + '_value'][s], rel=1e-12)
      np.testing.assert_allclose(gr, g[name + '_grad'][s], rtol=1e-11, atol=1e-12)
      np.testing.assert_allclose(H, g[name + '_hess'][s], rtol=1e-11, atol=1e-12)
  Z, theta = lr_problem(3, 400, 6, 4)
  rng = np.random.RandomState(5)
  idx = rng.choice(400, size=60, replace=False)
  w = rng.uniform(1., 12., size=60)
  mu, LSig, _ = models.get_laplace(w, Z[idx], np.zeros(6), models.lr_log_joint, models.lr_grad_th_log_joint, models.lr_hess_th_log_joint)
  smp = bc.LaplaceSampler('lr', 6)
  mu_d, LSig_d, _ = smp.get_laplace(w, Z[idx], np.zeros(6))
  np.testing.assert_allclose(mu_d, mu, rtol=1e-4, atol=1e-6)             # two runs of SciPy's BFGS to its gradient tolerance
  np.testing.assert_allclose(LSig_d, LSig, rtol=1e-4, atol=1e-7)
  np.random.seed(2)
  a = smp(50, w, Z[idx])
  np.random.seed(2)
  b = mu + np.random.randn(50, 6).dot(LSig.T)
  np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-5)
  np.random.seed(4)
  c = smp(5, None, None)                                                   # no coreset yet: prior (main.py:154-156)
  np.random.seed(4)
  assert np.array_equal(c, np.random.randn(5, 6))
  # and as the sampler of a device projector inside SparseVI (runs end to end)
  prj = bc.LogisticRegressionProjector(bc.LaplaceSampler('lr', 6), 32)
  np.random.seed(1)
  svi = bc.SparseVICoreset(Z, prj, opt_itrs=4)
  svi.build(3)
  assert 1 <= svi.size() <= 3 and np.all(svi.wts >= 0)


@pytest.mark.parametrize('model,alg', [('gaussian', 'fw'), ('poisson', 'giga')])
def test_full_size_audit_other_models(bc, model, alg):
  """the Gaussian and Poisson projections at N = 1e6 against the replay oracle + independent float64 audit (the LR model
  is covered at N = 1e6 and N = 1e7 by test_full_size_selection_equals_float64_audit)"""
  import bayesiancoresets_b200._native as nat
  from oracle import replay
  N, S, itrs = 1000000, 256, 25
  rng = np.random.RandomState(7)
  if model == 'gaussian':
    d = 20
    Z = rng.randn(N, d) + 1.
    theta = 1. + 0.3*rng.randn(S, d)
    Si = np.eye(d) + 0.05*np.ones((d, d))
    prj = bc.GaussianProjector(lambda n, w, p: theta, S, Si)
    mid, f = nat.MODEL_GAUSSIAN, (lambda x, t: models.gaussian_loglik(x, t, Si, 0.))
  else:
    d = 8
    th_true = rng.randn(d)/np.sqrt(d)
    X = np.hstack((rng.randn(N, d - 1), np.ones((N, 1))))
    Z = np.hstack((X, rng.poisson(np.log1p(np.exp(X.dot(th_true)))).astype(np.float64)[:, None]))
    theta = th_true + 0.1*rng.randn(S, d)
    Si = None
    prj = bc.PoissonProjector(lambda n, w, p: theta, S)
    mid, f = nat.MODEL_POISSON, models.poisson_loglik
  cs = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  cs.build(itrs)
  ev = cs.snnls.last_events
  ds = nat.Dataset(Z)
  _, norms, b = ds.audit(mid, theta, Si, norms=True, colsum=True)
  np.testing.assert_allclose(cs.snnls.b, b, rtol=1e-9, atol=1e-9*np.abs(b).max())
  kw = {'norm_sum': norms.sum()} if alg == 'fw' else {}
  r = replay.REPLAYS[alg](N, b, audit_score_fn(nat, ds, mid, theta, Si), lambda idx: models.project(f, Z[idx], theta), **kw)
  rev = r.build(itrs)
  assert [(e.code, e.f) for e in ev] == [(e[0], e[1]) for e in rev]
  assert_weights_close(cs.snnls.weights(), r.w)
  atol = 2.**-24*float(np.abs(r.wa).dot(norms[r.idx])) + 1e-300
  np.testing.assert_allclose([e.error for e in ev], [e[2] for e in rev], rtol=W_RTOL, atol=atol)


# ---------------------------------------------------------------- float16 pre-filter of the persistent kernels
def _events(s):
  return [(e.code, e.nact, e.f, e.error, e.aux0, e.aux1) for e in s.last_events]


@pytest.mark.gpu
@pytest.mark.parametrize('alg', ['giga', 'fw', 'omp'])
@pytest.mark.parametrize('N,d,S', [(60000, 6, 130), (200000, 8, 256), (50001, 5, 300), (150000, 10, 512)])
def test_filter16_changes_nothing_but_the_bytes(bc, alg, N, d, S):
  """the persistent kernel that streams the float16 copy and re-scans by bounds (csrc/filter_bounds.h) produces the SAME
  event log, bit for bit, as the one that streams the float32 rows -- selections, errors, weights, float64 scores"""
  Z, theta = lr_problem(5 + S, N, d, S)
  prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
  a = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  b = bc.HilbertCoreset(Z, prj, snnls=algs(bc)[alg])
  on, _ = a.snnls._native.filter16_stats()
  assert on, 'the float16 pre-filter should be available for S = %d' % S
  b.snnls._native.set_filter16(False)
  assert b.snnls._native.filter16_stats()[0] is False
  for itrs in (1, 30, 9):                                    # incremental builds: the state survives relaunches
    a.build(itrs)
    b.build(itrs)
    assert _events(a.snnls) == _events(b.snnls)
  np.testing.assert_array_equal(a.snnls.weights(), b.snnls.weights())
  assert a.error() == b.error()
  on, rows = a.snnls._native.filter16_stats()
  # every scan warp re-scans at least the ring stage that holds its own maximum; beyond that almost nothing
  assert on and 0 < rows < 40*(148*11*2*32 + 0.01*N), rows
  assert b.snnls._native.filter16_stats()[1] == 0


@pytest.mark.gpu
def test_filter16_near_ties_duplicates_and_slot_overflow(bc):
  """(a) near ties inside one warp's rows still reach the exact pass; (b) exact duplicates resolve to the lowest index;
  (c) a matrix made of 8 distinct rows repeated 150000 times puts a copy of the maximum into EVERY ring stage (16 rows)
  and gives every scan warp more stages than it has re-scan slots (32): the slots overflow, the iteration goes to the exact pass, the solver falls back to the float32 stream -- same result as
  with the filter off"""
  rng = np.random.RandomState(4)
  N, S = 40000, 256
  X = rng.randn(N, S)
  f0 = greedy.GigaOracle(X.T, X.sum(axis=0)).select()
  r = X[f0].copy()
  b = X.sum(axis=0)
  bh, rh = b/np.linalg.norm(b), r/np.linalg.norm(r)
  p = bh - bh.dot(rh)*rh
  p /= np.linalg.norm(p)
  for k in range(12):
    X[20000 + k] = np.linalg.norm(r)*(rh + 1e-6*(k + 1)*p)
  for alg in ('giga', 'fw'):
    oev = greedy.ORACLES[alg](X.T, X.sum(axis=0)).build(6)
    s = algs(bc)[alg](X.T, X.sum(axis=0))
    assert s._native.filter16_stats()[0]
    s.build(6)
    assert [e.f for e in s.last_events] == [e[1] for e in oev]
    assert s._native.exact_count() >= 1
  base = rng.randn(6000, 200)
  D = np.vstack((base, base[::-1].copy()))
  oev = greedy.GigaOracle(D.T, D.sum(axis=0)).build(30)
  s = algs(bc)['giga'](D.T, D.sum(axis=0))
  s.build(30)
  assert [e.f for e in s.last_events] == [e[1] for e in oev]
  R = np.tile(rng.randn(8, 160), (150000, 1))        # 75000 stages of 16 rows over 148 x 11 warps: 46 per warp
  res = []
  for filt in (True, False):
    s = algs(bc)['fw'](R.T, R.sum(axis=0))
    s._native.set_filter16(filt)
    s.build(8)
    res.append((_events(s), s._native.filter16_stats()[0], s._native.exact_count()))
  assert res[0][0] == res[1][0]
  assert all(e[2] < 8 for e in res[0][0])                    # lowest index among the 150000 copies
  assert res[0][1] is False and res[0][2] >= 1               # overflowed -> exact pass -> float32 stream from then on
