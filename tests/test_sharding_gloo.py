"""world_size-2 gloo tests (CPU) of the host-side N-sharding logic: shard layout, the one-off
all-reduce of b, mailbox-handle exchange, global index mapping, gathering the selected points
from the row-sharded dataset.  The native layer is replaced by tests/fake_native.py (a float64
NumPy test double of the device protocol); the result must equal the single-process oracle."""
import os
import sys
import socket
import numpy as np
import pytest
from conftest import ROOT


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, N, S, itrs, q, kind='gloo'):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  sys.path.insert(0, os.path.join(ROOT, 'tests'))
  import bayesiancoresets_b200 as bc
  from bayesiancoresets_b200 import _native as nat
  from bayesiancoresets_b200.snnls import base
  import fake_native
  fake_native.RANK[0] = rank
  if kind == 'gloo':
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    comm = bc.comm.TorchComm()
  else:
    comm = bc.comm.default_comm()         # the library's own torch-free process group (bcg_comm_*, TCP loopback)
    assert isinstance(comm, bc.comm.NativeComm) and 'torch' not in sys.modules
  nat.DeviceVecs.from_host = fake_native.FakeVecs.from_host
  base.nat.NativeSolver = fake_native.make_fake_solver(comm)
  np.random.seed(1)
  X = np.random.randn(N, S)
  lo, hi = bc.comm.even_shard(N, rank, world)

  class ShardProjector(object):
    def project(self, pts, grad=False):
      return pts
  cs = bc.HilbertCoreset(X[lo:hi], ShardProjector(), comm=comm)
  cs.build(itrs)
  wts, pts, idcs = cs.get()
  q.put((rank, cs.snnls.row_offset, cs.snnls.n_global, wts, pts, idcs, cs.error(), cs.snnls.weights()))
  comm.barrier()
  if kind == 'gloo':
    dist.destroy_process_group()
  else:
    comm.close()


@pytest.mark.parametrize('N,kind', [(1000, 'gloo'), (37, 'gloo'), (1000, 'native'), (37, 'native')])
def test_two_rank_sharded_hilbert_matches_oracle(N, kind):
  import multiprocessing as mp
  from oracle import greedy
  S, itrs, world = 20, 15, 2
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, N, S, itrs, q, kind)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  np.random.seed(1)
  X = np.random.randn(N, S)
  o = greedy.GigaOracle(X.T, X.sum(axis=0))
  o.build(itrs)
  assert [r[1] for r in res] == [0, (N + 1)//2] and all(r[2] == N for r in res)
  for rank, off, ng, wts, pts, idcs, err, w in res:
    assert np.array_equal(idcs, np.flatnonzero(o.w > 0))
    np.testing.assert_allclose(wts, o.w[o.w > 0], rtol=1e-9)
    assert np.array_equal(pts, X[idcs])           # gathered from both shards
    assert err == pytest.approx(o.error(), rel=1e-9)
    np.testing.assert_allclose(w, o.w, rtol=1e-9, atol=1e-12)


def _comm_worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import bayesiancoresets_b200 as bc
  c = bc.comm.default_comm()
  out = {}
  out['sum'] = c.allreduce_sum(np.arange(5.)*(rank + 1) + 0.1*rank)
  out['max'] = c.allreduce_max(np.array([rank, -rank, 7.]))
  out['bytes'] = c.allgather_bytes(bytes([rank])*3)
  out['obj'] = c.allgather_object({'rank': rank, 'payload': list(range(rank*1000))})     # different sizes per rank
  out['layout'] = bc.comm.shard_layout(c, 10 + rank)
  big = np.full(100000, float(rank + 1))                                                   # 800 KB per rank
  out['big'] = float(c.allreduce_sum(big).sum())
  for _ in range(50):
    c.barrier()
  q.put((rank, out, 'torch' in sys.modules))
  c.barrier()
  c.close()


@pytest.mark.parametrize('world', [2, 4])
def test_native_comm_collectives(world):
  """bcg_comm_* (the C-ABI's torch-free process group) on CPU: all-gather, rank-ordered all-reduce, barrier"""
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  exp_sum = sum(np.arange(5.)*(r + 1) + 0.1*r for r in range(world))
  for rank, out, torch_loaded in res:
    assert not torch_loaded
    assert np.array_equal(out['sum'], res[0][1]['sum'])               # bit-identical on every rank
    np.testing.assert_allclose(out['sum'], exp_sum, rtol=1e-15)
    assert list(out['max']) == [world - 1, 0., 7.]
    assert out['bytes'] == [bytes([r])*3 for r in range(world)]
    assert [o['rank'] for o in out['obj']] == list(range(world)) and len(out['obj'][-1]['payload']) == (world - 1)*1000
    assert out['layout'] == (sum(10 + r for r in range(rank)), sum(10 + r for r in range(world)), [10 + r for r in range(world)])
    assert out['big'] == 100000.*sum(r + 1 for r in range(world))


def test_even_shard_and_layout_serial():
  import bayesiancoresets_b200 as bc
  parts = [bc.comm.even_shard(10, r, 4) for r in range(4)]
  assert parts == [(0, 3), (3, 6), (6, 8), (8, 10)]
  assert bc.comm.shard_layout(bc.comm.SerialComm(), 7) == (0, 7, [7])
  assert bc.comm.even_shard(2, 3, 4) == (2, 2)      # more ranks than rows: empty shard
