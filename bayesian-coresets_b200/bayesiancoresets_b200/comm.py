"""Host-side plumbing for N-sharding over the GPUs of one node: one process per GPU (any launcher that
exports RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT, e.g. torchrun).  A communicator is used ONLY for
bootstrap (exchanging 64-byte mailbox handles, row counts) and for the one-off reductions of b and sum ||a_n|| after
a projection.  The per-iteration exchange of the greedy loop never goes through this module: it is fused into the
kernels over NVLink peer memory (csrc/loop_kernel.cuh, csrc/step_kernels.cuh).

  NativeComm   the library's own torch-free process group (bcg_comm_* of the C-ABI: TCP on the loopback interface);
               `default_comm()` builds it from the launcher's environment.  This is what bench.py and the multi-GPU
               checks use: no `import torch` anywhere in a worker.
  TorchComm    optional adapter over an already initialised torch.distributed group (NCCL or gloo)
  SerialComm   world of one
"""
import ctypes
import os
import pickle
import numpy as np


class SerialComm(object):
  """world of one (the reference's situation)"""
  rank = 0
  world = 1

  def allreduce_sum(self, arr):
    return np.array(arr, dtype=np.float64, copy=True)

  def allreduce_max(self, arr):
    return np.array(arr, dtype=np.float64, copy=True)

  def allgather_object(self, obj):
    return [obj]

  def allgather_bytes(self, blob):
    return [bytes(blob)]

  def barrier(self):
    pass

  def close(self):
    pass


class NativeComm(object):
  """bcg_comm: rank 0 listens on addr:port, the other ranks connect (star); all-reduces are rank-ordered, so every rank
  holds bit-identical results.  Needs no GPU and no torch."""
  def __init__(self, rank, world, addr='127.0.0.1', port=29611, timeout_s=120.):
    from . import _native as nat
    self._nat = nat
    self.handle = ctypes.c_void_p()
    nat.check(nat.lib().bcg_comm_create(str(addr).encode(), int(port), int(rank), int(world), int(timeout_s*1e3),
                                        ctypes.byref(self.handle)))
    self.rank, self.world = int(rank), int(world)

  def allgather_bytes(self, blob):
    """fixed-size all-gather: every rank passes `blob` (same length everywhere) and receives the list of all of them"""
    n = len(blob)
    out = ctypes.create_string_buffer(max(n*self.world, 1))
    self._nat.check(self._nat.lib().bcg_comm_allgather(self.handle, ctypes.c_char_p(bytes(blob)), n, out))
    raw = out.raw
    return [raw[r*n:(r + 1)*n] for r in range(self.world)]

  def _allreduce(self, arr, op):
    a = np.array(arr, dtype=np.float64, copy=True)
    flat = np.ascontiguousarray(a.reshape(-1))
    self._nat.check(self._nat.lib().bcg_comm_allreduce_f64(self.handle, ctypes.c_void_p(flat.ctypes.data), flat.shape[0], op))
    return flat.reshape(a.shape)

  def allreduce_sum(self, arr):
    return self._allreduce(arr, 0)

  def allreduce_max(self, arr):
    return self._allreduce(arr, 1)

  def allgather_object(self, obj):
    """variable-size objects: one fixed-size gather of the pickled lengths, one of the padded payloads"""
    blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sizes = [int.from_bytes(b, 'little') for b in self.allgather_bytes(len(blob).to_bytes(8, 'little'))]
    m = max(sizes)
    parts = self.allgather_bytes(blob + b'\0'*(m - len(blob)))
    return [pickle.loads(p[:n]) for p, n in zip(parts, sizes)]

  def barrier(self):
    self._nat.check(self._nat.lib().bcg_comm_barrier(self.handle))

  def close(self):
    if self.handle:
      self._nat.lib().bcg_comm_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


_DEFAULT = [None]


def default_comm():
  """the process-wide communicator described by the launcher's environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT;
  the library listens on BCG_COMM_PORT, default MASTER_PORT + 211, so it never collides with the launcher's store)"""
  if _DEFAULT[0] is None:
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
      _DEFAULT[0] = SerialComm()
    else:
      port = int(os.environ.get('BCG_COMM_PORT', int(os.environ.get('MASTER_PORT', '29400')) + 211))
      _DEFAULT[0] = NativeComm(int(os.environ['RANK']), world, os.environ.get('MASTER_ADDR', '127.0.0.1'), port)
  return _DEFAULT[0]


class TorchComm(object):
  """torch.distributed-backed communicator; the process group must already be initialised."""
  def __init__(self, group=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
      raise RuntimeError('torch.distributed is not initialised')
    self._torch, self._dist, self.group = torch, dist, group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    self.backend = dist.get_backend(group)
    if self.backend == 'nccl':
      self.device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', self.rank)))
      torch.cuda.set_device(self.device)
    else:
      self.device = torch.device('cpu')

  def allreduce_sum(self, arr):
    t = self._torch.from_numpy(np.array(arr, dtype=np.float64, copy=True)).to(self.device)
    self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
    return t.cpu().numpy()

  def allreduce_max(self, arr):
    t = self._torch.from_numpy(np.array(arr, dtype=np.float64, copy=True)).to(self.device)
    self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
    return t.cpu().numpy()

  def allgather_object(self, obj):
    out = [None]*self.world
    self._dist.all_gather_object(out, obj, group=self.group)
    return out

  def allgather_bytes(self, blob):
    return self.allgather_object(bytes(blob))

  def barrier(self):
    self._dist.barrier(group=self.group)

  def close(self):
    pass


def shard_layout(comm, n_local):
  """(row_offset of this rank, global row count, per-rank counts) for contiguous N-sharding."""
  counts = [int(c) for c in comm.allgather_object(int(n_local))]
  return sum(counts[:comm.rank]), sum(counts), counts


def even_shard(n_total, rank, world):
  """contiguous near-even split of range(n_total): returns (lo, hi) of `rank`"""
  base, rem = divmod(int(n_total), int(world))
  lo = rank*base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def local_part(sub_global, row_offset, n_local):
  """split a GLOBAL index draw (identical on every rank: the ranks share the seeded global RNG stream) into this
  rank's part: (positions within the draw, local row indices), order preserved"""
  sub_global = np.asarray(sub_global, dtype=np.int64)
  pos = np.flatnonzero((sub_global >= row_offset) & (sub_global < row_offset + n_local))
  return pos, sub_global[pos] - row_offset


def gather_rows(comm, data_local, row_offset, idcs_global):
  """rows `idcs_global` of the row-sharded dataset, assembled on every rank"""
  idcs_global = np.asarray(idcs_global, dtype=np.int64)
  mine = (idcs_global >= row_offset) & (idcs_global < row_offset + data_local.shape[0])
  piece = (np.flatnonzero(mine), data_local[idcs_global[mine] - row_offset])
  out = np.empty((idcs_global.shape[0],) + data_local.shape[1:], dtype=data_local.dtype)
  for pos, rows in comm.allgather_object(piece):
    out[pos] = rows
  return out
