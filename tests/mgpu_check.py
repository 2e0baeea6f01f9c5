"""Multi-GPU parity check, one process per GPU.  Any launcher that exports RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR /
MASTER_PORT works (tests/test_gpu_multi.py spawns the ranks itself; torchrun also does):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank holds a contiguous shard of the rows; the result (selection sequence, weights, error, points)
must equal the single-process oracle on the full data, on every rank.  The workers never import torch: the process
group is the library's own (bcg_comm_*)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'bayesian-coresets_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
  import bayesiancoresets_b200 as bc
  from oracle import greedy, models
  from conftest import lr_problem
  comm = bc.comm.default_comm()
  rank, world = comm.rank, comm.world
  algs = {'giga': bc.snnls.GIGA, 'fw': bc.snnls.FrankWolfe, 'omp': bc.snnls.OrthoPursuit}

  def check(N, d, S, itrs, alg, shard=None):
    Z, theta = lr_problem(21, N, d, S)
    vecs = models.project(models.lr_loglik, Z, theta)
    o = greedy.ORACLES[alg](vecs.T, vecs.sum(axis=0))
    oev = o.build(itrs)
    lo, hi = shard(rank, world) if shard else bc.comm.even_shard(N, rank, world)
    prj = bc.LogisticRegressionProjector(lambda n, w, p: theta, S)
    cs = bc.HilbertCoreset(Z[lo:hi], prj, snnls=algs[alg], comm=comm)
    cs.build(itrs)
    ev = cs.snnls.last_events
    nsel = len(oev) if alg != 'omp' else min(len(oev), int(0.6*S))
    got, ref = [e.f for e in ev][:nsel], [e[1] for e in oev][:nsel]
    assert got == ref, (alg, rank, got[:10], ref[:10])
    if alg != 'omp':
      w = cs.snnls.weights()
      assert np.max(np.abs(w - o.w)) <= 1e-5*np.abs(o.w).max(), (alg, rank)
      norms = np.sqrt((vecs**2).sum(axis=1))
      assert abs(cs.error() - o.error()) <= 1e-5*o.error() + 2.**-24*np.abs(o.w).dot(norms)
      wts, pts, idcs = cs.get()
      assert np.array_equal(idcs, np.flatnonzero(o.w > 0)) and np.array_equal(pts, Z[idcs])
    comm.barrier()
    if rank == 0:
      print('mgpu %s N=%d S=%d world=%d: %d iterations identical to the oracle' % (alg, N, S, world, nsel), flush=True)

  check(20000, 6, 128, 40, 'giga')
  check(20000, 6, 128, 40, 'fw')
  check(5000, 5, 64, 25, 'omp')
  check(30011, 8, 512, 30, 'giga')
  check(4099, 4, 50, 30, 'giga')
  # very uneven shards: rank 0 owns 3 rows, the last rank the rest
  def lopsided(r, w):
    N = 9000
    cuts = [0] + [3 + i for i in range(w - 1)] + [N]
    return cuts[r], cuts[r + 1]
  check(9000, 6, 256, 30, 'giga', shard=lopsided)
  # SparseVI / BatchPSVI N-sharded: same result as the single-process oracle
  from oracle import coresets
  rng = np.random.RandomState(5)
  x = rng.randn(3000, 4) + 1.
  th = rng.randn(24, 4)
  lo, hi = bc.comm.even_shard(3000, rank, world)
  np.random.seed(3)
  a = bc.SparseVICoreset(x[lo:hi], bc.GaussianProjector(lambda n, w, p: th + 0.01*np.random.randn(*th.shape), 24, np.eye(4)),
                         opt_itrs=6, comm=comm)
  a.build(4)
  np.random.seed(3)
  fg = lambda xx, tt: models.gaussian_loglik(xx, tt, np.eye(4), 0.)
  o = coresets.SparseVIOracle(x, models.OracleProjector(lambda n, w, p: th + 0.01*np.random.randn(*th.shape), 24, fg), opt_itrs=6)
  o.build(4)
  assert np.array_equal(a.idcs, o.idcs), (a.idcs, o.idcs)
  assert np.allclose(a.wts, o.wts, rtol=1e-6, atol=1e-9) and np.array_equal(a.pts, o.pts)
  Zb, thb = lr_problem(8, 2000, 4, 32)
  np.random.seed(9)
  bp = bc.BatchPSVICoreset(Zb[slice(*bc.comm.even_shard(2000, rank, world))], bc.LogisticRegressionProjector(lambda n, w, p: thb, 32),
                           opt_itrs=5, comm=comm)
  bp.build(6)
  np.random.seed(9)
  ob = coresets.BatchPSVIOracle(Zb, models.OracleProjector(lambda n, w, p: thb, 32, models.lr_loglik, models.lr_grad_z_loglik), opt_itrs=5)
  ob.build(6)
  assert np.allclose(bp.wts, ob.wts, rtol=1e-8) and np.allclose(bp.pts, ob.pts, rtol=1e-8, atol=1e-10)
  if rank == 0:
    print('mgpu sparsevi / bpsvi world=%d: identical to the oracle' % world, flush=True)
  os.environ['BCG_ENGINE'] = '1'          # launch-per-iteration engine with the block-wide exchange
  check(20000, 6, 128, 30, 'giga')
  # subsampling under N-sharding (hilbert.py:13-22): every rank draws the same global subsample
  os.environ['BCG_ENGINE'] = '2'
  Zs, ths = lr_problem(31, 12000, 5, 64)
  np.random.seed(17)
  ref_idx = np.unique(np.random.randint(12000, size=3000))
  vs = models.project(models.lr_loglik, Zs[ref_idx], ths)
  o = greedy.GigaOracle(vs.T, vs.sum(axis=0))
  o.build(20)
  np.random.seed(17)
  lo, hi = bc.comm.even_shard(12000, rank, world)
  cs = bc.HilbertCoreset(Zs[lo:hi], bc.LogisticRegressionProjector(lambda n, w, p: ths, 64), n_subsample=3000, comm=comm)
  cs.build(20)
  wts, pts, idcs = cs.get()
  assert np.array_equal(idcs, ref_idx[o.w > 0]), (rank, idcs[:5], ref_idx[o.w > 0][:5])
  assert np.allclose(wts, o.w[o.w > 0], rtol=1e-5) and np.array_equal(pts, Zs[idcs])
  if rank == 0:
    print('mgpu subsampled Hilbert world=%d: identical to the oracle' % world, flush=True)
  assert 'torch' not in sys.modules, 'the workers must not need torch'
  comm.barrier()
  if rank == 0:
    print('MGPU OK', flush=True)
  comm.close()


if __name__ == '__main__':
  main()
